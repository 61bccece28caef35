#!/bin/bash
# ncu launch list of one short single-slot bench run; prints the per-kernel table.  Usage: tools/ncu_launches.sh <tag>
tag=${1:-tmp}
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --batch 8 --slots 1 --no-cpu-baseline --no-4k > gpurun_out/ncu_b.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${tag}.csv
