#!/usr/bin/env python
"""Summarises an ncu --csv launch list (gpu__time_duration.sum, optionally dram__bytes_*):
per kernel name: launches, total time, share; writes JSON with the GEMM kernel's average DRAM traffic per launch.
  python tools/launch_summary.py launches.csv [out.json]"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
per = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rd:
    name = re.sub(r"\(.*", "", r.get("Kernel Name", "")).replace("void ", "").replace("t2v::", "")
    m, v, u = r.get("Metric Name"), r.get("Metric Value", "0").replace(",", ""), r.get("Metric Unit", "")
    try:
        v = float(v)
    except ValueError:
        continue
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)      # -> us
        per[name]["n"] += 1
    elif m and m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per[name][m] += v
tot = sum(d["gpu__time_duration.sum"] for d in per.values())
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>8s} {'dram MB/launch':>14s}")
out = {}
for name, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n = max(d["n"], 1)
    t = d["gpu__time_duration.sum"]
    traffic = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    print(f"{name[:60]:60s} {int(n):8d} {t:12.1f} {100 * t / tot:6.1f}% {t / n:8.2f} {traffic / n / 1e6:14.3f}")
    out[name] = {"launches": int(n), "total_us": t, "share": t / tot, "avg_us": t / n, "dram_bytes_per_launch": traffic / n}
print(f"{'TOTAL':60s} {int(sum(d['n'] for d in per.values())):8d} {tot:12.1f}")
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
