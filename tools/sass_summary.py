#!/usr/bin/env python3
"""Per-kernel SASS mnemonic counts of the built library -> profiles/sass_summary.txt.

    python tools/sass_summary.py [path/to/libbf_b200.so]

Evidence that the hot kernels use what DESIGN.md says they use: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D bulk TMA), UTMALDG = tensor-map TMA, FADD2 / FFMA2 =
packed fp32, BRX = indexed branch of the interpreter loop, SYNCS = mbarrier operations.  Static counts
(instructions in the binary, not executed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "zybo-rt-sampler-image-detection_b200", "lib", "libbf_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "FADD2", "FFMA2", "FMUL2", "HMMA", "DFMA",
         "BRX", "BSSY", "LDS", "LDGSTS", "SYNCS", "STL", "LDL"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else SO
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    names, counts, total = [], {}, {}
    cur = None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            names.append(cur)
            counts[cur] = collections.Counter()
            total[cur] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            if op in WATCH:
                counts[cur][op] += 1
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    out = ["# static SASS mnemonic counts per kernel of %s (tools/sass_summary.py)" % os.path.basename(so),
           "# " + " ".join(WATCH), ""]
    rows = []
    for n, d in zip(names, dem):
        d = re.sub(r"\(.*\)$", "", d).replace("void ", "").replace("bf::", "")
        rows.append((d, total[n], counts[n]))
    for d, t, c in sorted(rows):
        hits = "  ".join("%s %d" % (k, c[k]) for k in WATCH if c[k])
        out.append("%-72s %6d instr  %s" % (d[:72], t, hits))
    agg = collections.Counter()
    for _, _, c in rows:
        agg.update(c)
    out += ["", "library total: " + "  ".join("%s %d" % (k, agg[k]) for k in WATCH if agg[k])]
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)
    print(out[-1])


if __name__ == "__main__":
    main()
