"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and grid) -> markdown table."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = r["Kernel Name"].split("(")[0]
    key = (name, r.get("Grid Size", ""))
    agg[key][0] += 1
    agg[key][1] += ns
    total += ns
byname = defaultdict(lambda: [0, 0.0])
for (n, g), (c, t) in agg.items():
    byname[n][0] += c
    byname[n][1] += t
print(f"total kernel time {total/1e6:.2f} ms over {sum(c for c, _ in agg.values())} launches\n")
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for n, (c, t) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {c} | {t/1e6:.3f} | {100*t/total:.1f}% |")
print("\n| kernel | grid | launches | ms | share |\n|---|---|---:|---:|---:|")
for (n, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {n} | {g} | {c} | {t/1e6:.3f} | {100*t/total:.1f}% |")
