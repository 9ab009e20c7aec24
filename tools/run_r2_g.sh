#!/bin/bash
mkdir -p gpurun_out
PFHE_NTT_PERSIST=0 timeout 600 bash tools/ncu_kernel.sh fwd14 ntt_tma_kernel 2 -- python tools/gpu_fwd.py 14 4096
python tools/ncu_raw_summary.py gpurun_out/ncu_fwd14.raw.csv > gpurun_out/r2g_ncu_fwd14.txt 2>&1
python tools/ncu_src_summary.py gpurun_out/ncu_fwd14.src.csv 14 >> gpurun_out/r2g_ncu_fwd14.txt 2>&1
timeout 600 bash tools/ncu_kernel.sh fwd14p ntt_persist 2 -- python tools/gpu_fwd.py 14 4096
python tools/ncu_raw_summary.py gpurun_out/ncu_fwd14p.raw.csv > gpurun_out/r2g_ncu_fwd14p.txt 2>&1
python tools/ncu_src_summary.py gpurun_out/ncu_fwd14p.src.csv 14 >> gpurun_out/r2g_ncu_fwd14p.txt 2>&1
cat gpurun_out/r2g_ncu_fwd14.txt gpurun_out/r2g_ncu_fwd14p.txt
