#!/bin/bash
mkdir -p gpurun_out
{ timeout 900 python -m pytest tests/test_gpu_ext.py -x -q -m gpu 2>&1 | tail -4
  timeout 900 python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; tail -3 gpurun_out/r2ad_bench.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ad_bench.json'))
e=d['e2e']; print('pinned',e['value'],'pageable',e['pageable']['value'],'registered',e['registered'])
PY
} > gpurun_out/r2ad.log 2>&1
cat gpurun_out/r2ad.log
