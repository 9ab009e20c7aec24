#!/bin/bash
# A/B of the cp.async key prefetch in the generic u64 external product + parity of every lattice test with it on
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_baseline_shapes.py tests/test_gpu_paths.py tests/test_gpu_rns.py tests/test_gpu_ext.py -x -q -m gpu 2>&1 | tail -4
  echo "== prefetch on"; timeout 120 python tools/gpu_br.py ep
  echo "== prefetch off"; PFHE_EP_KEY_PREFETCH=0 timeout 120 python tools/gpu_br.py ep
  echo "== prefetch on"; timeout 120 python tools/gpu_br.py ep
  echo "== prefetch off"; PFHE_EP_KEY_PREFETCH=0 timeout 120 python tools/gpu_br.py ep; } > gpurun_out/r2t.log 2>&1
cat gpurun_out/r2t.log
