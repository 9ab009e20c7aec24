#!/bin/bash
mkdir -p gpurun_out
{ for st in 0 250 500 1000 2000; do echo "== cluster stagger $st ns"; PFHE_NTT_CLUSTER=1 PFHE_NTT_CLUSTER_STAGGER_NS=$st python tools/gpu_c3.py 2>&1 | sed -n 5,8p; done
  echo "== one CTA per polynomial"; python tools/gpu_c3.py 2>&1 | sed -n 5,8p; } > gpurun_out/r2am.log 2>&1
cat gpurun_out/r2am.log
