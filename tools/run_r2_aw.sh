#!/bin/bash
# the whole GPU parity suite again with the alternative kernel families forced by environment (oracle comparisons, not only digests)
mkdir -p gpurun_out
{ echo "== PFHE_DISABLE_F64=1 (integer pipe everywhere)"; PFHE_DISABLE_F64=1 timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
  echo "== PFHE_NTT_TMA=0 PFHE_BR_FAST=0 PFHE_EP_FAST=0 PFHE_NTT_CLUSTER=0 PFHE_POLYMUL_STASH=0 (round-1 style kernels)"; PFHE_NTT_TMA=0 PFHE_BR_FAST=0 PFHE_EP_FAST=0 PFHE_NTT_CLUSTER=0 PFHE_POLYMUL_STASH=0 timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
  echo "== PFHE_NTT_CLUSTER=2 PFHE_DCRT_EP_FUSED_WIDE=1 PFHE_STAGE=0 (opt-in kernels)"; PFHE_NTT_CLUSTER=2 PFHE_DCRT_EP_FUSED_WIDE=1 PFHE_STAGE=0 timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2; } > gpurun_out/r2aw.log 2>&1
cat gpurun_out/r2aw.log
