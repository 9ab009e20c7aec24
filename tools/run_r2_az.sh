#!/bin/bash
mkdir -p gpurun_out
for v in 2 1; do
PFHE_NTT_CLUSTER=$v timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2az_bench_$v.json 2> gpurun_out/r2az_bench_$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2az_bench_$v.json')); e=d['extra']
print('PFHE_NTT_CLUSTER=$v value',d['value'],'c3',e.get('rns_polymuls_per_s_n16384_l8_u64'),e.get('rns_polymul_n16384_l8_roofline',{}).get('frac_of_fp64_pipe'),'dcrt fwd',e.get('dcrt_ntt_fwd_n16384_l8_roofline',{}).get('frac_of_hbm_peak'), all(d['parity_checks'].values()), e.get('error'))
PY
done
