#!/bin/bash
mkdir -p gpurun_out
{ timeout 120 python tools/gpu_fwd.py 12 32768 1152921504606830593
  timeout 120 python tools/gpu_fwd.py 13 16384 1152921504606830593
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "--- launch list: multi-limb external product"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv python tools/gpu_dcrt_ep.py 2>/dev/null | grep -E "gadget|dcrt_external" | awk -F'","' '{print $5, $NF}' | tail -8
  echo "--- ncu streaming kernels"
  timeout 600 bash tools/ncu_kernel.sh sliceop slice_op_kernel 3 -- python tools/gpu_stream.py
  python tools/ncu_raw_summary.py gpurun_out/ncu_sliceop.raw.csv
  timeout 600 bash tools/ncu_kernel.sh gadget rns_gadget_kernel 1 -- python tools/gpu_stream.py
  python tools/ncu_raw_summary.py gpurun_out/ncu_gadget.raw.csv
  timeout 600 bash tools/ncu_kernel.sh compose rns_compose_kernel 1 -- python tools/gpu_stream.py
  python tools/ncu_raw_summary.py gpurun_out/ncu_compose.raw.csv
  timeout 600 bash tools/ncu_kernel.sh baseconv baseconv_kernel 1 -- python tools/gpu_stream.py
  python tools/ncu_raw_summary.py gpurun_out/ncu_baseconv.raw.csv
} > gpurun_out/r2i.log 2>&1
cat gpurun_out/r2i.log
