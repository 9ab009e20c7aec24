#!/bin/bash
# round-2 GPU session A: parity of the new paths, blind-rotation A/B, ncu of the re-scheduled kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
( timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_paths.py tests/test_gpu_lattice.py -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2a_pytest.log
{
for nb in 1250 2500 10000; do
  for mb in 3 4 5 6; do echo "fast MINB=$mb"; PFHE_BR_MINB=$mb timeout 300 python tools/gpu_br.py br $nb; done
  echo "generic"; PFHE_BR_FAST=0 timeout 300 python tools/gpu_br.py br $nb
done
} > gpurun_out/r2a_br.log 2>&1
NLWE=32 timeout 600 bash tools/ncu_kernel.sh br_fast blind_rotate 1 -- python tools/gpu_br.py br 2500
python tools/ncu_raw_summary.py gpurun_out/ncu_br_fast.raw.csv > gpurun_out/r2a_ncu_br_fast.txt 2>&1
python tools/ncu_src_summary.py gpurun_out/ncu_br_fast.src.csv 16 >> gpurun_out/r2a_ncu_br_fast.txt 2>&1
cat gpurun_out/r2a_pytest.log gpurun_out/r2a_br.log
