#!/bin/bash
mkdir -p gpurun_out
{ for i in 14 13; do for p in 1 0; do echo "PERSIST=$p"; PFHE_NTT_PERSIST=$p timeout 120 python tools/gpu_fwd.py $i $((65536*4096/(1<<i)/2)); done; done
  timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "c3 or 13 or 14 or c2 or mixed" 2>&1 | tail -4
  timeout 600 bash tools/ncu_kernel.sh fwd14p ntt_persist 2 -- python tools/gpu_fwd.py 14 4096
  python tools/ncu_raw_summary.py gpurun_out/ncu_fwd14p.raw.csv; python tools/ncu_src_summary.py gpurun_out/ncu_fwd14p.src.csv 10; } > gpurun_out/r2h.log 2>&1
cat gpurun_out/r2h.log
