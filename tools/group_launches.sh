#!/bin/bash
# ncu launch list (gpu__time_duration + instructions) of a short bench run with frame groups; per-kernel table.
# Usage: tools/group_launches.sh <tag> [bench args]
tag=${1:-tmp}; shift
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 260 --csv \
    --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --batch 16 --slots 1 --workers 1 --no-cpu-baseline --no-4k "$@" > gpurun_out/ncu_b.log 2>&1
python tools/ncu_budget.py gpurun_out/launches_${tag}.csv
