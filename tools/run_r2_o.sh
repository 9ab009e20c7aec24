#!/bin/bash
mkdir -p gpurun_out
{ timeout 300 python tools/gpu_c3.py; timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_ntt.py -x -q -m gpu -k "c3 or polymul or 13 or 14 or c2" 2>&1 | tail -3; } > gpurun_out/r2o.log 2>&1
cat gpurun_out/r2o.log
