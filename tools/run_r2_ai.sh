#!/bin/bash
mkdir -p gpurun_out
{ for cfg in "1 4" "1 5" "0 4" "0 5"; do set -- $cfg; echo "== cluster=$1 loge=$2"; PFHE_NTT_CLUSTER=$1 PFHE_BIGN_LOGE=$2 timeout 300 python tools/gpu_c3.py 2>&1 | sed -n 5,8p; done
  PFHE_NTT_CLUSTER=1 PFHE_BIGN_LOGE=4 timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "16384 or 14 or c3 or C3 or dcrt" 2>&1 | tail -3; } > gpurun_out/r2ai.log 2>&1
cat gpurun_out/r2ai.log
