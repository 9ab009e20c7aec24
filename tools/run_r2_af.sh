#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "16384 or 14 or c3 or C3 or dcrt" 2>&1 | tail -6
  echo "== cluster"; timeout 300 python tools/gpu_c3.py 2>&1 | head -8
  echo "== one CTA per polynomial"; PFHE_NTT_CLUSTER=0 timeout 300 python tools/gpu_c3.py 2>&1 | head -8; } > gpurun_out/r2af.log 2>&1
cat gpurun_out/r2af.log
