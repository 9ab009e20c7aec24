#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2c_pytest.log
{ echo "STASH=1"; timeout 300 python tools/gpu_c3.py; echo "STASH=0"; PFHE_POLYMUL_STASH=0 timeout 300 python tools/gpu_c3.py; } > gpurun_out/r2c_c3.log 2>&1
{ timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/gpu_sanitize.py 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py 2>&1 | tail -3
  timeout 900 compute-sanitizer --tool synccheck python tools/gpu_sanitize.py 2>&1 | tail -3; } > gpurun_out/r2c_sanitizer.log 2>&1
tail -6 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_c3.log gpurun_out/r2c_sanitizer.log
