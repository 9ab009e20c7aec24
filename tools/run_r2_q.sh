#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_baseline_shapes.py tests/test_gpu_paths.py tests/test_gpu_rns.py tests/test_gpu_ext.py -x -q -m gpu 2>&1 | tail -4
  PFHE_EP_FAST=0 timeout 120 python tools/gpu_br.py ep
  PFHE_BR_FAST=0 timeout 200 python tools/gpu_br.py br 2500; } > gpurun_out/r2q.log 2>&1
cat gpurun_out/r2q.log
