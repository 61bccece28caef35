#!/bin/bash
# Per-kernel instruction / pipe budget of one frame.  Usage: tools/ncu_budget.sh <tag>
tag=${1:-tmp}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__cycles_active.avg,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
    --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/budget_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --batch 8 --slots 1 --no-cpu-baseline --no-4k > gpurun_out/ncu_b.log 2>&1
python tools/ncu_budget.py gpurun_out/budget_${tag}.csv
