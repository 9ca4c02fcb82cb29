"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total us, share."""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(rows)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    tot[name][0] += 1
    tot[name][1] += us
s = sum(v[1] for v in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'us':>10s} {'share':>6s}")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {us:10.1f} {us / s * 100:5.1f}%")
print(f"{'total':60s} {sum(v[0] for v in tot.values()):8d} {s:10.1f}")
