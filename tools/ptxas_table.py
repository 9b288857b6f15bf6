#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` output (stdin) as one line per kernel: registers, spills, smem."""
import re, subprocess, sys
txt = sys.stdin.read()
names = {}
cur = None
rows = []
for ln in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        cur = {"name": m.group(1)}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m:
        cur["stack"], cur["st"], cur["ld"] = map(int, m.groups())
    m = re.search(r"Used (\d+) registers", ln)
    if m:
        cur["regs"] = int(m.group(1))
dem = subprocess.run(["c++filt"] + [r["name"] for r in rows], capture_output=True, text=True).stdout.splitlines()
for r, d in zip(rows, dem):
    d = re.sub(r"\(bf::MimoParams\)|void bf::", "", d)
    print(f"{d:75s} regs {r.get('regs', '?'):>3} stack {r.get('stack', 0):>4} spill st/ld {r.get('st', 0):>4}/{r.get('ld', 0):<4}")
