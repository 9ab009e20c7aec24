#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ext.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2av.log
cat gpurun_out/r2av.log
