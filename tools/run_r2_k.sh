#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2k_pytest.log
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
timeout 300 python bench.py --driver capi --gpus 1 > gpurun_out/r2k_capi1.json 2> gpurun_out/r2k_capi1.err
cat gpurun_out/r2k_pytest.log; tail -c 600 gpurun_out/r2k_bench.err; cat gpurun_out/r2k_capi1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'sust',d['roofline'].get('frac_sustained'),'e2e',d['e2e']['value'],'pageable',d['e2e']['pageable']['value'])
print('bs',d['bootstrap']['value'],d['bootstrap']['e2e']['value'],d['bootstrap']['roofline']['frac'])
print(d['parity_checks'])
print({k:v for k,v in d['extra'].items() if not isinstance(v,dict)})
PY
