#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2ar.log
cat gpurun_out/r2ar.log
