#!/bin/bash
mkdir -p gpurun_out
{ echo "== fused (128 registers)"; timeout 300 python tools/gpu_dcrt_ep.py; } > gpurun_out/r2ab.log 2>&1
cat gpurun_out/r2ab.log
