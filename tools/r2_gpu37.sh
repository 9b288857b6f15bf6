#!/bin/bash
# round-2 GPU call 37: float64 factor kernels with the cp.async per-thread prefetch ring
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python tools/mvdr_c4.py > $O/r2_g37_mvdr.log 2>&1; tail -1 $O/r2_g37_mvdr.log
timeout 900 python -m pytest tests/test_gpu_mvdr.py tests/test_gpu_c4_size.py -x -q 2>&1 | tail -3
