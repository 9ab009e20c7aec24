#!/bin/bash
mkdir -p gpurun_out
{ echo "LOGE=4"; PFHE_BIGN_LOGE=4 timeout 300 python tools/gpu_c3.py; 
  PFHE_BIGN_LOGE=4 timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "c3 or 13 or 14 or c2 or mixed" 2>&1 | tail -4; } > gpurun_out/r2f_c3.log 2>&1
cat gpurun_out/r2f_c3.log
