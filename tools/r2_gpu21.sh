#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --no-cpu --no-extras --steps 3 --warmup 3 --algo pad"
BF_MIMO_DUAL=1 $B > $O/r2_g21_plain.log 2>&1 && \
BF_MIMO_DUAL=1 ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 4 -c 1 -f -o $O/r2x_mimo_pad_dual $B > $O/r2_g21_ncu.log 2>&1
ls -la $O/r2x_mimo_pad_dual.ncu-rep
