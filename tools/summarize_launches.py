"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void\s+", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit.startswith("ns") else (v if unit.startswith("us") else v * 1000.0)
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print("%-70s %6s %10s %7s %8s" % ("kernel", "n", "total_us", "share", "avg_us"))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %6d %10.1f %6.1f%% %8.2f" % (k[:70], n, us, 100 * us / tot, us / n))
print("%-70s %6d %10.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))
