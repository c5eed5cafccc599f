"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid, block)."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
lines = [l for l in open(path) if not l.startswith("==")]
agg = OrderedDict()
total = 0.0
for i, x in enumerate(csv.DictReader(lines)):
    if i < lo or i >= hi:
        continue
    name = re.sub(r"^void ", "", x["Kernel Name"])
    name = re.sub(r"\(.*", "", name)
    key = (name, x["Grid Size"], x["Block Size"])
    us = float(x["Metric Value"].replace(",", "")) / 1000.0
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
print(f"{'kernel':60s} {'grid':>16s} {'n':>5s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
for (name, grid, block), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {grid:>16s} {n:5d} {us:10.1f} {us / n:8.1f} {100 * us / total:5.1f}%")
print(f"total {total:.1f} us")
