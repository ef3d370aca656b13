"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/launch_summary.py file.csv"""
import collections
import csv
import sys

for path in sys.argv[1:]:
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    print(path)
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"  {k[:56]:56s} n={n:4d} total={t:9.1f} us avg={t / n:8.1f} us")
