"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"<.*", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("us") else v * 1e3 if unit.startswith("ms") else v)
    agg[name][0] += 1
    agg[name][1] += v_us
    total += v_us
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>9s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {n:8d} {t:12.1f} {100 * t / total:6.1f}% {t / n:9.2f}")
print(f"{'TOTAL':60s} {sum(v[0] for v in agg.values()):8d} {total:12.1f}")
