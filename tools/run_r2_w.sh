#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests/test_gpu_rns.py tests/test_gpu_baseline_shapes.py tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -4
  timeout 600 python tools/gpu_rns_stream.py 2>&1 | grep -v "^Exception\|^Traceback\|File\|TypeError"; } > gpurun_out/r2w.log 2>&1
cat gpurun_out/r2w.log
