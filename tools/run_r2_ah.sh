#!/bin/bash
mkdir -p gpurun_out
{ timeout 1500 python -m pytest tests/test_gpu_paths.py tests/test_gpu_rns.py tests/test_gpu_pointwise.py -x -q -m gpu 2>&1 | tail -6; } > gpurun_out/r2ah.log 2>&1
cat gpurun_out/r2ah.log
