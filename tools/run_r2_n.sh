#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests/test_gpu_rns.py tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -6
  echo fused; timeout 120 python tools/gpu_dcrt_ep.py; echo two-kernel; PFHE_DCRT_EP_TWO_KERNEL=1 timeout 120 python tools/gpu_dcrt_ep.py; } > gpurun_out/r2n.log 2>&1
cat gpurun_out/r2n.log
