#!/bin/bash
mkdir -p gpurun_out
timeout 400 bash tools/ncu_kernel.sh clfwd "ntt_cluster_kernel" 2 -- python tools/gpu_fwd.py 14 4096
python tools/ncu_raw_summary.py gpurun_out/ncu_clfwd.raw.csv > gpurun_out/r2ag_ncu_cluster_fwd.txt 2>&1
python tools/ncu_src_summary.py gpurun_out/ncu_clfwd.src.csv 16 >> gpurun_out/r2ag_ncu_cluster_fwd.txt 2>&1
grep -E "launch__|occupancy|cluster" gpurun_out/ncu_clfwd.raw.csv | head -5
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/ncu_clfwd.raw.csv')))
h,u,v=rows[0],rows[1],rows[2]
for i,x in enumerate(h):
    if any(k in x for k in ('launch__occupancy','launch__cluster','sm__ctas_launched','launch__waves','sm__maximum_warps','launch__block','launch__grid','achieved_occupancy','sm__warps_active')):
        print(x,v[i],u[i])
PY
rm -f gpurun_out/ncu_clfwd.src.csv gpurun_out/ncu_clfwd.raw.csv
cat gpurun_out/r2ag_ncu_cluster_fwd.txt
