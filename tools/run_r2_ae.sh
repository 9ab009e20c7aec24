#!/bin/bash
mkdir -p gpurun_out
export PFHE_DCRT_EP_FUSED_WIDE=1
{ timeout 1500 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py 2>&1 | tail -12
  timeout 1500 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py 2>&1 | tail -6
  timeout 1500 compute-sanitizer --tool synccheck python tools/gpu_sanitize.py 2>&1 | tail -4; } > gpurun_out/r2ae_sanitizer.log 2>&1
cat gpurun_out/r2ae_sanitizer.log
