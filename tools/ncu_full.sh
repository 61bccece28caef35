#!/bin/bash
# ncu --set full capture of one launch of the named kernel(s) from a short single-slot bench run.
# Usage: tools/ncu_full.sh <tag> <kernel-regex>      -> gpurun_out/full_<tag>.ncu-rep + gpurun_out/full_<tag>.csv
tag=${1:-tmp}; kern=${2:-k_matching}
ncu --set full --import-source on --clock-control none -k regex:"$kern" -s 4 -c 1 -f -o gpurun_out/full_${tag} \
    python bench.py --steps 1 --warmup 3 --batch 8 --slots 1 --workers 1 --no-cpu-baseline --no-4k > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/full_${tag}.ncu-rep --page raw --csv > gpurun_out/full_${tag}.csv 2>/dev/null
