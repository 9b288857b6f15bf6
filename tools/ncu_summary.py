#!/usr/bin/env python3
"""Turn an .ncu-rep capture into the short text summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r1_<kernel>.txt
"""
import csv
import subprocess
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none  (%s)" % rep.split("/")[-1])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("\n== %s  (launch id %s)" % (d.get("Kernel Name"), d.get("ID")))
        for k in KEYS:
            if k in d and d[k] != "":
                print("  %-82s %s %s" % (k, d[k], units[hdr.index(k)]))
        st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[k]))
              for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k]]
        st = sorted(st, key=lambda kv: -kv[1])[:8]
        print("  warp stall reasons per issued instruction: " + ", ".join("%s %.2f" % kv for kv in st))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    for n, start in enumerate(hi[:1]):
        h = rows[start]
        end = hi[n + 1] - 1 if n + 1 < len(hi) else len(rows)
        body = rows[start + 1:end]
        ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        c, ce = Counter(), Counter()
        for r in body:
            t = r[ia].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            c[op] += int(r[isamp] or 0)
            ce[op] += int(r[ie] or 0)
        tot = sum(c.values()) or 1
        print("\n-- SASS opcode mix of the first captured launch (warp-level instructions executed, pc samples)")
        for op, n_ in ce.most_common(12):
            print("  %-10s executed %12d   samples %5.1f%%" % (op, n_, 100.0 * c[op] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
