#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total, share, average.
    python tools/ncu_launch_summary.py gpurun_out/launches_r01.csv > profiles/r01_ncu_launches_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.DictReader(lines)
tot = OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6
    c, t = tot.get(name, (0, 0.0))
    tot[name] = (c + 1, t + us)
total = sum(t for _, t in tot.values()) or 1.0
print("%-62s %5s %14s %8s %12s" % ("kernel", "count", "total_us", "share", "avg_us"))
for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %5d %14.1f %7.1f%% %12.1f" % (name[:62], c, t, 100 * t / total, t / c))
