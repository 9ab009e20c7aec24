#!/bin/bash
mkdir -p gpurun_out
{ timeout 600 python -m pytest tests/test_gpu_paths.py -x -q -m gpu -k "variants" 2>&1 | tail -3
  timeout 300 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "16384 or 14 or c3 or C3 or dcrt" 2>&1 | tail -2
  PFHE_NTT_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "16384 or 14 or c3 or C3 or dcrt" 2>&1 | tail -2
  for r in 1 2; do
    echo "== round $r one CTA per polynomial"; PFHE_NTT_CLUSTER=0 timeout 120 python tools/gpu_c3.py 2>&1 | sed -n 5,8p
    echo "== round $r cluster, barrier.cluster (transforms + product)"; PFHE_NTT_CLUSTER=2 PFHE_NTT_CLUSTER_ASYNC=0 timeout 120 python tools/gpu_c3.py 2>&1 | sed -n 5,8p
    echo "== round $r cluster, st.async (transforms + product)"; PFHE_NTT_CLUSTER=2 timeout 120 python tools/gpu_c3.py 2>&1 | sed -n 5,8p
  done
  echo "== thermal (cluster async product)"; PFHE_NTT_CLUSTER=2 timeout 120 python tools/gpu_c3_thermal.py; } > gpurun_out/r2ay.log 2>&1
cat gpurun_out/r2ay.log
