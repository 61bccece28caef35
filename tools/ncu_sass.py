"""Hot SASS of an ncu report: python tools/ncu_sass.py file.ncu-rep [min_share_pct]  (runs ncu --page source --csv)"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(float(r["Instructions Executed"] or 0) for r in rows)
samp = sum(float(r["# Samples"] or 0) for r in rows)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
print(f"total warp instructions {tot:.0f}, samples {samp:.0f}, {len(rows)} SASS lines")
for i, r in enumerate(rows):
    ex = float(r["Instructions Executed"] or 0); s = float(r["# Samples"] or 0)
    if 100 * ex / tot >= thr or 100 * s / max(samp, 1) >= thr:
        print(f"{i:5d} {100*ex/tot:5.2f}% ex {100*s/max(samp,1):5.2f}% smp  thr {r['Avg. Threads Executed']:>5s}  {r['Source'][:110]}")
