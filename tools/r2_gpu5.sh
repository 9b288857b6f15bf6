#!/bin/bash
# round-2 GPU call 5: MVDR v3 (kind::f16) correctness + timing, then the whole GPU suite
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_mvdr.py -x -q -s 2>&1 | grep -v "Will use" | tail -12 > $O/r2_g5_mvdr.log
timeout 300 python tools/mvdr_c4.py --tc 3 > $O/r2_g5_c4_tc3.log 2>&1
timeout 300 python tools/mvdr_c4.py --tc 2 > $O/r2_g5_c4_tc2.log 2>&1
timeout 600 python -m pytest tests/test_gpu_c4_size.py -x -q -s 2>&1 | grep -v "Will use" | tail -12 > $O/r2_g5_c4size.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/r2_g5_all.log
tail -5 $O/r2_g5_mvdr.log $O/r2_g5_c4_tc3.log $O/r2_g5_c4_tc2.log $O/r2_g5_c4size.log $O/r2_g5_all.log
