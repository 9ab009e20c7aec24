#!/bin/bash
mkdir -p gpurun_out
{ timeout 600 python -m pytest tests/test_gpu_rns.py -x -q -m gpu -k "base_conv" 2>&1 | tail -2
  PFHE_DISABLE_F64=1 timeout 600 python -m pytest tests/test_gpu_rns.py tests/test_gpu_paths.py -x -q -m gpu -k "base_conv or variants" 2>&1 | tail -2
  timeout 300 python tools/gpu_rns_stream.py 2>&1 | grep baseconv
  echo "== integer path forced"; PFHE_DISABLE_F64=1 timeout 300 python tools/gpu_rns_stream.py 2>&1 | grep baseconv; } > gpurun_out/r2bd.log 2>&1
cat gpurun_out/r2bd.log
