"""Turns an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch, --csv)
into profiles/traffic_rNN.json: DRAM bytes per step and per launch for every kernel, plus the same figures for the
kernel groups bench.py reports (`roofline.traffic` is read from the "groups" table).

  python tools/make_traffic.py gpurun_out/launches_r01_final.csv profiles/traffic_r01.json
"""
import csv
import json
import re
import sys
from collections import defaultdict

GROUPS = [  # bench.py kernel-group name, regex on the demangled kernel name
    ("conv_tc_kernel(fprop+dgrad)", r"^tc::conv_tc_kernel"),
    ("wgrad_tc_kernel", r"^tc::wgrad_tc_kernel"),
    ("basi_bn_bwd_reduce", r"^bn_bwd_reduce"),
    ("basi_bn_bwd_apply", r"^bn_bwd_apply"),
    ("basi_bn_bwd_fused", r"^bn_bwd_resident"),
    ("basi_bn_bwd_coop", r"^bn_bwd_coop"),
    ("basi_bn_apply", r"^bn_apply"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(src, dst):
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    per = defaultdict(lambda: {"ids": set(), "us": 0.0, "bytes": 0.0})
    for r in csv.DictReader(lines):
        name = re.sub(r"^void\s+", "", re.sub(r"\(.*", "", r["Kernel Name"]))
        v = float(r["Metric Value"].replace(",", ""))
        m, unit = r["Metric Name"], r["Metric Unit"]
        a = per[name]
        if m == "gpu__time_duration.sum":
            a["ids"].add(r["ID"])
            a["us"] += v / 1000.0 if unit.startswith("ns") else (v if unit.startswith("us") else v * 1000.0)
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            a["bytes"] += v * SCALE.get(unit, 1.0)
    out = {"source": src, "kernels": {}, "groups": {}}
    for k, a in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        n = len(a["ids"])
        out["kernels"][k] = {"launches": n, "dram_bytes_per_step": a["bytes"], "dram_bytes_per_launch": a["bytes"] / max(n, 1),
                             "us_per_step_ncu": round(a["us"], 2)}
    for g, rx in GROUPS:
        ks = [a for k, a in per.items() if re.search(rx, k)]
        n = sum(len(a["ids"]) for a in ks)
        if n:
            b = sum(a["bytes"] for a in ks)
            out["groups"][g] = {"launches": n, "dram_bytes_per_step": b, "dram_bytes_per_launch": b / n,
                                "us_per_step_ncu": round(sum(a["us"] for a in ks), 2)}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote %s (%d kernels, %d groups)" % (dst, len(out["kernels"]), len(out["groups"])))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
