"""Prints the interesting metrics of an `ncu --page raw --csv` export.  Usage: python tools/ncu_read.py file.csv [regex]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2:]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
    r"gpu__time_duration.sum|sm__cycles_active.avg|sm__cycles_elapsed.avg$|smsp__inst_executed.sum$|dram__bytes_(read|write).sum$|"
    r"launch__(grid_size|block_size|registers_per_thread|waves_per_multiprocessor|occupancy_limit.*)|sm__warps_active.avg.pct_of_peak|"
    r"smsp__issue_active.avg.pct|sm__inst_executed_pipe_(xu|alu|fma|lsu|fmaheavy|uniform)\.|smsp__pcsamp_warps_issue_stalled_.*(?<!not_issued)$|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|lts__t_sector_hit_rate.pct|sm__throughput.avg.pct|l1tex__t_sector_hit_rate.pct|smsp__warp_issue_stalled.*_per_warp_active.pct")
for v in vals:
    print("==", v[hdr.index("Kernel Name")][:60])
    out = []
    for h, u, x in zip(hdr, units, v):
        if pat.search(h):
            out.append((h, x, u))
    for h, x, u in out:
        print(f"  {h:75s} {x:>14s} {u}")
