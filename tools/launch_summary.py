"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, mean us, share."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("elasb::<unnamed>::", "").replace("void ", "")
    agg[name].append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':42s} {'launches':>8s} {'mean us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:42]:42s} {len(v):8d} {sum(v)/len(v):9.2f} {100*sum(v)/tot:6.1f}%")
frames = max(len(v) for v in agg.values()) if agg else 1
print(f"sum of kernel time per frame (serialised, cold cache): {tot/ max(len(agg.get('k_support', [1])),1):.1f} us")
