#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X): name, launches, mean, total us.
usage: tools/launch_summary.py launches.csv [skip_first_n]"""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "us")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        rows.append((r["Kernel Name"], v, r.get("Grid Size", "")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
acc = defaultdict(list)
for name, v, grid in rows:
    acc[(name.split("(")[0][:70], grid)].append(v)
tot = sum(v for _, v, _ in rows)
print(f"{len(rows)} launches, {tot:.1f} us in total")
for (name, grid), vs in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(vs):10.1f} us {100 * sum(vs) / tot:5.1f} %  n={len(vs):4d}  mean {sum(vs) / len(vs):8.2f}  max {max(vs):8.2f}  grid {grid:>14s}  {name}")
