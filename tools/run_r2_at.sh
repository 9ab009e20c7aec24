#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_extprod_functional.py tests/test_gpu_bootstrap_functional.py -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2at.log
cat gpurun_out/r2at.log
