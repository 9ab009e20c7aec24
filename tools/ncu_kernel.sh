#!/bin/bash
# generic ncu capture: tools/ncu_kernel.sh <tag> <kernel-regex> <skip> -- <command...>; keeps only the raw/source CSV pages
tag=$1; regex=$2; skip=$3; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/ncu_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_$tag.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_$tag.src.csv 2>/dev/null
rm -f gpurun_out/ncu_$tag.ncu-rep
