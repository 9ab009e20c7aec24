#!/bin/bash
# 8-GPU session: concurrent PCIe probe at 1/2/4/8 ranks, multi-device C-ABI test on real distinct devices, bench at N = 8 (both drivers)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1; nproc >> gpurun_out/r2m_topo.txt; lscpu | grep -E "NUMA|Socket|Model name" >> gpurun_out/r2m_topo.txt
{ for n in 1 2 4 8; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) tools/gpu_pcie_multi.py 2>/dev/null | grep ranks; done; } > gpurun_out/r2m_pcie.log
( timeout 600 python -m pytest tests/test_gpu_ext.py -x -q -m gpu -k "multi_device or bootstrap_slices" 2>&1 | tail -4 ) > gpurun_out/r2m_pytest.log
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 50 --warmup 3 --no-extra > gpurun_out/r2m_bench_${n}gpu.json 2> gpurun_out/r2m_bench_${n}gpu.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref_8gpu.json 2>/dev/null
for n in 8 2; do timeout 600 python bench.py --driver capi --gpus $n > gpurun_out/r2m_capi_${n}gpu.json 2> gpurun_out/r2m_capi_${n}gpu.err; done
cat gpurun_out/r2m_pcie.log gpurun_out/r2m_pytest.log gpurun_out/r2m_capi_8gpu.json gpurun_out/r2m_capi_2gpu.json; head -c 600 gpurun_out/r2m_bench_8gpu.json; tail -c 400 gpurun_out/r2m_bench_8gpu.err
