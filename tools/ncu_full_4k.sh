#!/bin/bash
# ncu --set full capture of one k_matching launch at 4096x2160 (BASELINE.json configs[4] geometry).
ncu --set full --import-source on --clock-control none -k regex:k_matching -s 2 -c 1 -f -o gpurun_out/full_k7_4k \
    python tools/stage_times.py 4096 2160 256 3 > gpurun_out/ncu_full_4k.log 2>&1
ncu -i gpurun_out/full_k7_4k.ncu-rep --page raw --csv > gpurun_out/full_k7_4k.csv 2>/dev/null
