#!/usr/bin/env python3
"""Record a per-launch ncu figure in profiles/traffic.json, keyed to the kernel sources it was captured from.

    python tools/traffic_entry.py KEY REPORT.ncu-rep KERNEL_REGEX [--metric dram|tensor_pct] SOURCE.cu [SOURCE ...]

dram:       dram__bytes_read.sum + dram__bytes_write.sum of the first matching launch (bytes)
tensor_pct: sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
bench.py reports an entry only while the hash of the listed sources is unchanged (bench.py:_traffic)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    metric = "dram"
    for a in sys.argv[1:]:
        if a.startswith("--metric"):
            metric = a.split("=", 1)[1] if "=" in a else sys.argv[sys.argv.index(a) + 1]
    args = [a for a in args if a != metric]
    key, rep, rx = args[0], args[1], args[2]
    sources = args[3:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    row = next(r for r in rows[2:] if re.search(rx, r[ix["Kernel Name"]]))
    num = lambda name: float(row[ix[name]].replace(",", ""))
    if metric == "dram":
        def to_bytes(name):
            unit = rows[1][ix[name]].lower()
            scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
            return num(name) * scale
        value = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    else:
        value = num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    import bench
    path = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = {"value": value, "kernel": row[ix["Kernel Name"]][:80], "capture": os.path.basename(rep),
                 "sources": sources, "source_sha": bench._source_sha(sources)}
    json.dump(data, open(path, "w"), indent=1)
    print(key, data[key])


if __name__ == "__main__":
    main()
