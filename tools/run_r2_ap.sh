#!/bin/bash
mkdir -p gpurun_out
{ python tools/gpu_c3_thermal.py; PFHE_NTT_CLUSTER=0 python tools/gpu_c3_thermal.py; } > gpurun_out/r2ap.log 2>&1
cat gpurun_out/r2ap.log
