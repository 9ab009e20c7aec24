#!/bin/bash
# final 8-GPU session: in-library multi-device driver at 1/2/4/8 devices, torchrun bench at 2/4/8 (with extras off), reference arm at 8
mkdir -p gpurun_out
for n in 1 2 4 8; do timeout 600 python bench.py --driver capi --gpus $n > gpurun_out/r2s_capi_${n}gpu.json 2> gpurun_out/r2s_capi_${n}gpu.err; done
( timeout 600 python -m pytest tests/test_gpu_ext.py -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/r2s_pytest.log
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 50 --warmup 3 --no-extra > gpurun_out/r2s_bench_${n}gpu.json 2> gpurun_out/r2s_bench_${n}gpu.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/r2s_bench_ref_8gpu.json 2>/dev/null
cat gpurun_out/r2s_capi_*gpu.json gpurun_out/r2s_pytest.log
python - <<'PY'
import json
for n in (2,4,8):
    d=json.load(open(f'gpurun_out/r2s_bench_{n}gpu.json')); b=d['bootstrap']
    print(n,'ntt',f"{d['value']:.4e}",'e2e',f"{d['e2e']['value']:.4e}",'pageable',f"{d['e2e']['pageable']['value']:.4e}",'pcie_frac',d['e2e'].get('pcie_frac'),'| bs',f"{b['value']:.4e}",'e2e',f"{b['e2e']['value']:.4e}",'roof',round(b['roofline']['frac'],3))
d=json.load(open('gpurun_out/r2s_bench_ref_8gpu.json')); print('ref8',d['value'],d['cpu_baseline'].get('scalar_port_value'),d['cpu_baseline']['cores'],d['bootstrap']['value'])
PY
