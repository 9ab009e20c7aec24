#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2ao_pytest.log
{ timeout 900 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool synccheck python tools/gpu_sanitize.py 2>&1 | tail -4; } > gpurun_out/r2ao_sanitizer.log 2>&1
timeout 900 python bench.py > gpurun_out/r2ao_bench.json 2> gpurun_out/r2ao_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ao_smoke.log 2>&1
cat gpurun_out/r2ao_pytest.log gpurun_out/r2ao_sanitizer.log gpurun_out/r2ao_smoke.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ao_bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'sust',d['roofline'].get('frac_sustained'),'e2e',d['e2e']['value'])
print('bs',d['bootstrap']['value'],d['bootstrap']['e2e']['value'],d['bootstrap']['roofline']['frac'])
print(all(d['parity_checks'].values()))
e=d['extra']; print('c3',e.get('rns_polymuls_per_s_n16384_l8_u64'),e.get('rns_polymul_n16384_l8_roofline'),e.get('dcrt_ntt_fwd_n16384_l8_roofline'),e.get('error'))
PY
