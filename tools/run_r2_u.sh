#!/bin/bash
# ncu --set full of the round-2 kernels that had no capture yet (one launch each), summaries only
mkdir -p gpurun_out
cap() { tag=$1; shift; timeout 400 bash tools/ncu_kernel.sh "$@"; python tools/ncu_raw_summary.py gpurun_out/ncu_$tag.raw.csv > gpurun_out/r2u_ncu_$tag.txt 2>&1; python tools/ncu_src_summary.py gpurun_out/ncu_$tag.src.csv 12 >> gpurun_out/r2u_ncu_$tag.txt 2>&1; rm -f gpurun_out/ncu_$tag.src.csv gpurun_out/ncu_$tag.raw.csv; }
cap ep32_fast ep32_fast external_product_u32_kernel 1 -- python tools/gpu_br.py ep
cap ep64 ep64 "external_product_kernel" 1 -- python tools/gpu_br.py ep
TERNARY=1 NLWE=32 cap br_ternary br_ternary blind_rotate_ternary 1 -- python tools/gpu_br.py br 2500
cap dcrt_ep_fused dcrt_ep_fused dcrt_external_product_fused 1 -- python tools/gpu_dcrt_ep.py
cap polymul_stash_n16384 polymul_stash_n16384 polymul_kernel 1 -- python tools/gpu_c3.py
cap ntt_fwd_n4096 ntt_fwd_n4096 ntt_tma_kernel 2 -- python tools/gpu_fwd.py 12 65536
ls -la gpurun_out/r2u_*; head -30 gpurun_out/r2u_ncu_ep32_fast.txt
