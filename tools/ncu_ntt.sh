#!/bin/bash
# ncu capture of the headline forward NTT kernel (one launch) for a given variant; usage: tools/ncu_ntt.sh <tag> [env...]
# writes gpurun_out/ncu_<tag>.ncu-rep + raw csv + source csv
tag=$1; shift
mkdir -p gpurun_out
env "$@" ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 3 -c 1 -f -o gpurun_out/ncu_$tag \
    python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_$tag.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_$tag.src.csv 2>/dev/null
rm -f gpurun_out/ncu_$tag.ncu-rep
