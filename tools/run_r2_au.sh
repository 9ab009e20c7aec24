#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "c3" 2>&1 | tail -8 ) > gpurun_out/r2au.log
cat gpurun_out/r2au.log
