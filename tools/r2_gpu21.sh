#!/bin/bash
# round-2 GPU call 21: float64 Cholesky / triangular inverse at 4 resident CTAs per SM (one wave for 512 bins) against the old build
# (BF_MVDR_MINB was a one-call hook: the 128-register build won and is now the only one, see profiles/r2_kernel_variants.md)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for m in 1 4; do
  BF_MVDR_MINB=$m python tools/mvdr_c4.py > $O/r2_g21_mvdr_minb$m.log 2>&1
  tail -3 $O/r2_g21_mvdr_minb$m.log
done
python -m pytest tests/test_gpu_mvdr.py tests/test_gpu_c4_size.py -x -q 2>&1 | tail -3
