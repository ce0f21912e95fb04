"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel markdown table.
    python tools/ncu_launch_table.py gpurun_out/launches.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else (v * 1e3 if r["Metric Unit"].startswith("ms") else v)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("ctc::", "").strip()
    rows.append((name, r["Grid Size"], r["Block Size"], us))
agg = OrderedDict()
for name, grid, block, us in rows:
    a = agg.setdefault((name, block), [0, 0.0, set()])
    a[0] += 1; a[1] += us; a[2].add(grid)
total = sum(us for *_, us in rows)
print(f"{len(rows)} launches, {total / 1e3:.2f} ms\n")
print("| kernel | grid | block | launches | total ms | share | avg us |\n|---|---|---|---|---|---|---|")
for (name, block), (n, us, grids) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    g = next(iter(grids)) if len(grids) == 1 else f"{len(grids)} shapes"
    print(f"| `{name}` | {g} | {block} | {n} | {us / 1e3:.3f} | {100 * us / total:.1f}% | {us / n:.1f} |")
