#!/bin/bash
mkdir -p gpurun_out
{ echo "== normal"; python tools/gpu_fwd.py 14 4096; python tools/gpu_fwd.py 14 4096
  cp primus_fhe_b200/lib/libpfhe_exp.so primus_fhe_b200/lib/libpfhe_cuda.so
  echo "== EXPERIMENT small twiddle footprint in the last pass (wrong results, timing only)"; python tools/gpu_fwd.py 14 4096; python tools/gpu_fwd.py 14 4096; } > gpurun_out/r2al.log 2>&1
cat gpurun_out/r2al.log
