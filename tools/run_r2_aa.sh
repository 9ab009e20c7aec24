#!/bin/bash
mkdir -p gpurun_out
{ timeout 1200 python -m pytest tests/test_gpu_rns.py tests/test_gpu_paths.py tests/test_gpu_lattice.py tests/test_gpu_baseline_shapes.py -x -q -m gpu 2>&1 | tail -6
  echo "== fused"; timeout 300 python tools/gpu_dcrt_ep.py
  echo "== two kernels"; PFHE_DCRT_EP_TWO_KERNEL=1 timeout 300 python tools/gpu_dcrt_ep.py; } > gpurun_out/r2aa.log 2>&1
cat gpurun_out/r2aa.log
