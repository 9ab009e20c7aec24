#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_baseline_shapes.py tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2e_pytest.log
{ echo "PERSIST=1"; timeout 300 python tools/gpu_c3.py; echo "PERSIST=0"; PFHE_NTT_PERSIST=0 timeout 300 python tools/gpu_c3.py; } > gpurun_out/r2e_c3.log 2>&1
cat gpurun_out/r2e_pytest.log gpurun_out/r2e_c3.log
