#!/bin/bash
mkdir -p gpurun_out
cap() { tag=$1; shift; timeout 400 bash tools/ncu_kernel.sh "$@"; python tools/ncu_raw_summary.py gpurun_out/ncu_$tag.raw.csv > gpurun_out/r2x_ncu_$tag.txt 2>&1; python tools/ncu_src_summary.py gpurun_out/ncu_$tag.src.csv 14 >> gpurun_out/r2x_ncu_$tag.txt 2>&1; rm -f gpurun_out/ncu_$tag.src.csv gpurun_out/ncu_$tag.raw.csv; }
cap baseconv baseconv "baseconv_kernel" 1 -- python tools/gpu_rns_stream.py
cap compose compose "rns_compose_kernel" 1 -- python tools/gpu_rns_stream.py
cap decompose decompose "rns_decompose_kernel" 1 -- python tools/gpu_rns_stream.py
cap gadget gadget "rns_gadget_kernel" 1 -- python tools/gpu_rns_stream.py
head -60 gpurun_out/r2x_ncu_baseconv.txt
