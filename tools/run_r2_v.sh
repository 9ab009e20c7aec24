#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_rns_stream.py > gpurun_out/r2v_rns_stream.log 2>&1
cat gpurun_out/r2v_rns_stream.log | tail -30
