#!/bin/bash
mkdir -p gpurun_out
{ for r in 1 2 3; do
    echo "== round $r one CTA per polynomial"; PFHE_NTT_CLUSTER=0 python tools/gpu_c3.py 2>&1 | sed -n 5,8p
    echo "== round $r cluster stagger 0"; PFHE_NTT_CLUSTER=1 PFHE_NTT_CLUSTER_STAGGER_NS=0 python tools/gpu_c3.py 2>&1 | sed -n 5,8p
    echo "== round $r cluster stagger 1000"; PFHE_NTT_CLUSTER=1 PFHE_NTT_CLUSTER_STAGGER_NS=1000 python tools/gpu_c3.py 2>&1 | sed -n 5,8p
  done; } > gpurun_out/r2an.log 2>&1
grep -E "==|DCRT polymul|single fwd" gpurun_out/r2an.log
