#!/bin/bash
# round-2 GPU session B: whole GPU suite, the new bench line, reference arm, single-process capi driver, launch list
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2b_pytest.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
OMP_NUM_THREADS=1 timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref_omp1.json 2>> gpurun_out/r2b_bench_ref.err
timeout 300 python bench.py --driver capi --gpus 1 > gpurun_out/r2b_capi1.json 2> gpurun_out/r2b_capi1.err
tail -5 gpurun_out/r2b_pytest.log; tail -c 1500 gpurun_out/r2b_bench.err; head -c 3000 gpurun_out/r2b_bench.json; cat gpurun_out/r2b_capi1.json
