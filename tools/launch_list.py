#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch log into the table committed under profiles/.
    python tools/launch_list.py gpurun_out/launches.csv "<command line that was profiled>" > profiles/rN_launches_bench.txt"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].replace("bf::", "").split("(")[0]
    key = (name, r[ix["Grid Size"]], r[ix["Block Size"]])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    c, t = agg.get(key, (0, 0.0))
    agg[key] = (c + 1, t + ms)
print("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<library kernels> : %s" % (sys.argv[2] if len(sys.argv) > 2 else ""))
print("# per-launch times are cold-cache and serialised (compare shares, not absolutes)")
print("%-52s %-16s %-13s %5s %12s %12s" % ("kernel", "grid", "block", "count", "total ms", "ms/launch"))
for (name, grid, block), (c, t) in agg.items():
    print("%-52s %-16s %-13s %5d %12.3f %12.4f" % (name[:52], grid.replace(" ", ""), block.replace(" ", ""), c, t, t / c))
