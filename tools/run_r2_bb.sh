#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2bb_smoke.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_baseline_shapes.py tests/test_gpu_ext.py -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/r2bb_pytest.log
timeout 400 bash tools/ncu_kernel.sh clasync "ntt_cluster_kernel" 2 -- python tools/gpu_fwd.py 14 4096
python tools/ncu_raw_summary.py gpurun_out/ncu_clasync.raw.csv > gpurun_out/r2bb_ncu_cluster_async.txt 2>&1
python tools/ncu_src_summary.py gpurun_out/ncu_clasync.src.csv 14 >> gpurun_out/r2bb_ncu_cluster_async.txt 2>&1
rm -f gpurun_out/ncu_clasync.src.csv gpurun_out/ncu_clasync.raw.csv
cat gpurun_out/r2bb_smoke.log gpurun_out/r2bb_pytest.log; head -48 gpurun_out/r2bb_ncu_cluster_async.txt
