"""Per-kernel table (mean per launch) from tools/ncu_budget.sh output."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("elasb::<unnamed>::", "").replace("void ", "")
    m, unit = row["Metric Name"], row["Metric Unit"]
    if m == "gpu__time_duration.sum":
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    if "bytes" in m:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    agg[name][m].append(v)
cols = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum"]
print(f"{'kernel':28s} {'n':>3s} {'us':>7s} {'Minst':>7s} {'Mxu':>6s} {'Mlsu':>6s} {'kcyc_act':>8s} {'dramR MB':>8s} {'dramW MB':>8s} {'L2 MB':>7s}")
tot = collections.Counter()
for k, d in sorted(agg.items(), key=lambda kv: -sum(kv[1][cols[0]])):
    mean = [sum(d[c]) / max(len(d[c]), 1) for c in cols]
    if "elementwise" in k: continue
    print(f"{k[:28]:28s} {len(d[cols[0]]):3d} {mean[0]:7.2f} {mean[1]/1e6:7.2f} {mean[2]/1e6:6.2f} {mean[3]/1e6:6.2f} {mean[4]/1e3:8.1f} {mean[5]/1e6:8.2f} {mean[6]/1e6:8.2f} {mean[7]/1e6:7.1f}")
    for c, m in zip(cols, mean): tot[c] += m
print(f"{'per frame':28s}     {tot[cols[0]]:7.2f} {tot[cols[1]]/1e6:7.2f} {tot[cols[2]]/1e6:6.2f} {tot[cols[3]]/1e6:6.2f} {tot[cols[4]]/1e3:8.1f} {tot[cols[5]]/1e6:8.2f} {tot[cols[6]]/1e6:8.2f} {tot[cols[7]]/1e6:7.1f}")
