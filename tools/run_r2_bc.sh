#!/bin/bash
mkdir -p gpurun_out
{ timeout 600 python -m pytest tests/test_gpu_ext.py -x -q -m gpu 2>&1 | tail -3
  timeout 300 python bench.py --driver capi --gpus 2; } > gpurun_out/r2bc.log 2>&1
cat gpurun_out/r2bc.log
