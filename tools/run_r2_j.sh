#!/bin/bash
mkdir -p gpurun_out
{ timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_ext.py tests/test_gpu_lattice.py tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -4
  PFHE_STAGE=0 timeout 120 python tools/gpu_e2e_pageable.py
  for th in 2 4 8; do PFHE_STAGE_THREADS=$th timeout 120 python tools/gpu_e2e_pageable.py; done
  PFHE_STAGE_THREADS=8 PFHE_PIPE_CHUNK_MB=16 timeout 120 python tools/gpu_e2e_pageable.py
  PFHE_STAGE_THREADS=8 PFHE_PIPE_CHUNK_MB=64 timeout 120 python tools/gpu_e2e_pageable.py
  nproc; } > gpurun_out/r2j.log 2>&1
cat gpurun_out/r2j.log
