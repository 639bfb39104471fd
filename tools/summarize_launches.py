"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
per kernel name: launches, total / average duration, share of the step and (when captured) DRAM bytes."""
import csv
import re
import sys
from collections import defaultdict

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
agg = defaultdict(lambda: {"ids": set(), "us": 0.0, "r": 0.0, "w": 0.0})
scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
for r in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void\s+", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    a = agg[name]
    m = r.get("Metric Name")
    if m == "gpu__time_duration.sum":
        a["ids"].add(r["ID"])
        a["us"] += v / 1000.0 if unit.startswith("ns") else (v if unit.startswith("us") else v * 1000.0)
    elif m == "dram__bytes_read.sum":
        a["r"] += v * scale.get(unit, 1e-6)
    elif m == "dram__bytes_write.sum":
        a["w"] += v * scale.get(unit, 1e-6)
tot = sum(v["us"] for v in agg.values())
print("%-66s %5s %10s %6s %8s %10s %10s" % ("kernel", "n", "total_us", "share", "avg_us", "dramR_MB", "dramW_MB"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = len(v["ids"])
    print("%-66s %5d %10.1f %5.1f%% %8.2f %10.1f %10.1f" % (k[:66], n, v["us"], 100 * v["us"] / tot, v["us"] / max(n, 1),
                                                          v["r"], v["w"]))
print("%-66s %5d %10.1f" % ("TOTAL", sum(len(v["ids"]) for v in agg.values()), tot))
