"""Summarises an `ncu --page source --csv` dump by runs of equally-often executed instructions.
Usage: python tools/ncu_source_blocks.py <source.csv> [min_share_percent] [-v]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, it, isamp = hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
data = rows[2:]
tot = sum(int(r[ia]) for r in data)
warps = int(data[0][ia])
share = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 0.4
print(f"total warp instructions {tot}, warps {warps}, per warp {tot / warps:.0f}, samples {sum(int(r[isamp]) for r in data)}")
if "-v" in sys.argv:
    for i, r in enumerate(data):
        print(f"{i:4d} {int(r[ia]):8d} {float(r[it]):5.1f} {int(r[isamp]):4d}  {r[1].strip()[:90]}")
    sys.exit()
start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(int(data[i][ia]) - int(data[start][ia])) > 0.02 * max(int(data[start][ia]), 1) + 5:
        n = sum(int(r[ia]) for r in data[start:i]); s = sum(int(r[isamp]) for r in data[start:i])
        thr = sum(float(r[it]) for r in data[start:i]) / (i - start)
        if n > share / 100 * tot:
            print(f"[{start:4d}-{i - 1:4d}] len {i - start:4d} x {int(data[start][ia]) / warps:5.2f}/warp = {100 * n / tot:5.1f}%  samples {s:4d} thr {thr:4.1f}")
        start = i
