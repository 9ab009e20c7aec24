#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests/test_gpu_pointwise.py tests/test_gpu_ntt.py tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -4
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err; tail -3 gpurun_out/r2ak_bench.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ak_bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'checks',all(d['parity_checks'].values()))
for k,v in d['extra'].get('streaming_kernels',{}).items(): print(k, round(v['GB/s']), round(v['frac_of_hbm_peak'],3))
print('c3',d['extra'].get('rns_polymuls_per_s_n16384_l8_u64'), d['extra'].get('error'))
PY
} > gpurun_out/r2ak.log 2>&1
cat gpurun_out/r2ak.log
