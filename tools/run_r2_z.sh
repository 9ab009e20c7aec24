#!/bin/bash
mkdir -p gpurun_out
{ for m in 0 2 0 2; do echo "== key prefetch mode $m"; PFHE_EP_KEY_PREFETCH=$m timeout 120 python tools/gpu_br.py ep | grep u64; done; } > gpurun_out/r2z.log 2>&1
cat gpurun_out/r2z.log
