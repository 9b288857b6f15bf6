#!/bin/bash
# round-2 GPU call 38: ncu of the float64 Cholesky / inverse kernels (C4 size)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python tools/mvdr_c4.py --reps 1 > $O/r2_g38_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mvdr_chol_blocked\|mvdr_trinv_blocked -c 2 -f -o $O/r2h_mvdr_factor python tools/mvdr_c4.py --reps 1 > $O/r2_g38_ncu.log 2>&1
ls -la $O/r2h_mvdr_factor.ncu-rep; tail -2 $O/r2_g38_ncu.log
