#!/bin/bash
mkdir -p gpurun_out
{ for i in 1 2; do timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2aj_bench$i.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2aj_bench$i.json'));e=d['extra'];print('bench',e.get('rns_polymuls_per_s_n16384_l8_u64'),e.get('polymul_per_s_n8192_u64'),e.get('error'))"; done
  timeout 300 python tools/gpu_c3.py 2>&1 | head -8; } > gpurun_out/r2aj.log 2>&1
cat gpurun_out/r2aj.log
