#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests/test_gpu_rns.py tests/test_gpu_pointwise.py -x -q -m gpu 2>&1 | tail -4
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -3 gpurun_out/r2y_bench.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/r2y_bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'checks',all(d['parity_checks'].values()) if 'parity_checks' in d else None)
for k,v in d['extra']['streaming_kernels'].items(): print(k, round(v['GB/s']), round(v['frac_of_hbm_peak'],3))
print({k:v for k,v in d.get('parity_checks',{}).items() if not v})
PY
} > gpurun_out/r2y.log 2>&1
cat gpurun_out/r2y.log
