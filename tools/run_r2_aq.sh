#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_paths.py tests/test_gpu_ntt.py tests/test_gpu_baseline_shapes.py -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/r2aq_pytest.log
timeout 900 python bench.py > gpurun_out/r2aq_bench.json 2> gpurun_out/r2aq_bench.err
cat gpurun_out/r2aq_pytest.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2aq_bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'sust',d['roofline'].get('frac_sustained'),'e2e',d['e2e']['value'], d['e2e']['pageable']['value'], d['e2e']['registered']['value'])
print('bs',d['bootstrap']['value'],d['bootstrap']['e2e']['value'],d['bootstrap']['roofline']['frac'])
print(all(d['parity_checks'].values()))
e=d['extra']; print('c3',e.get('rns_polymuls_per_s_n16384_l8_u64'),e.get('rns_polymul_n16384_l8_roofline',{}).get('frac_of_fp64_pipe'),e.get('dcrt_ntt_fwd_n16384_l8_roofline'),e.get('error'))
PY
