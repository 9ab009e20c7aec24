#!/bin/bash
mkdir -p gpurun_out
{ timeout 600 python -m pytest tests/test_gpu_paths.py -x -q -m gpu -k "variants" 2>&1 | tail -3
  for r in 1 2; do
    echo "== round $r one CTA per polynomial"; timeout 120 python tools/gpu_n8192.py
    echo "== round $r cluster (transforms + product)"; PFHE_NTT_CLUSTER13=2 timeout 120 python tools/gpu_n8192.py
  done; } > gpurun_out/r2ba.log 2>&1
cat gpurun_out/r2ba.log
