#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2bf_bench.json 2> gpurun_out/r2bf_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2bf_bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'sust',d['roofline'].get('frac_sustained'),'e2e',d['e2e']['value'])
print('bs',d['bootstrap']['value'],d['bootstrap']['e2e']['value'],d['bootstrap']['roofline']['frac'], all(d['parity_checks'].values()))
e=d['extra']; print('c3',e.get('rns_polymuls_per_s_n16384_l8_u64'), e.get('error'))
for k,v in e.get('streaming_kernels',{}).items(): print(k, round(v['frac_of_hbm_peak'],3))
PY
