#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/r2f2_pytest.log
{ timeout 600 compute-sanitizer --tool memcheck python tools/gpu_sanitize.py 2>&1 | tail -3
  timeout 600 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py 2>&1 | tail -3
  timeout 600 compute-sanitizer --tool synccheck python tools/gpu_sanitize.py 2>&1 | tail -3; } > gpurun_out/r2f2_sanitizer.log 2>&1
timeout 900 python bench.py > gpurun_out/r2f2_bench.json 2> gpurun_out/r2f2_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f2_bench_ref.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f2_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f2_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --sustain-s 0 > /dev/null 2>&1
cat gpurun_out/r2f2_pytest.log gpurun_out/r2f2_sanitizer.log gpurun_out/r2f2_smoke.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f2_bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'sust',d['roofline'].get('frac_sustained'),'e2e',d['e2e']['value'],'pageable',d['e2e']['pageable']['value'],'registered',d['e2e']['registered']['value'])
print('bs',d['bootstrap']['value'],d['bootstrap']['e2e']['value'],d['bootstrap']['roofline']['frac'])
print(all(d['parity_checks'].values()), d['parity_checks'])
for k,v in d['extra'].items():
    print(k, v if not isinstance(v,dict) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if not isinstance(b,(dict,str))})
PY
