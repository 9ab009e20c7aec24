#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ext.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2d_pytest.log
{ TERNARY=1 timeout 300 python tools/gpu_br.py br 10000; TERNARY=1 timeout 300 python tools/gpu_br.py br 1250; } > gpurun_out/r2d_br.log 2>&1
cat gpurun_out/r2d_pytest.log gpurun_out/r2d_br.log
