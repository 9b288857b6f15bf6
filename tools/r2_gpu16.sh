#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for tc in 3 4; do for dbg in 0 4; do
  BF_MVDR_DBG=$dbg timeout 300 python tools/mvdr_c4.py --tc $tc > $O/r2_g16_c4_tc${tc}_dbg$dbg.log 2>&1
  echo "tc $tc dbg $dbg: $(grep -o '"steering": [0-9.]*' $O/r2_g16_c4_tc${tc}_dbg$dbg.log)"
done; done
BF_MVDR_TC=4 BF_MVDR_DBG=4 timeout 600 python -m pytest tests/test_gpu_c4_size.py -x -q -s -k mvdr 2>&1 | grep -v "Will use" | tail -4
BF_MVDR_TC=4 BF_MVDR_DBG=4 timeout 600 python -m pytest tests/test_gpu_mvdr.py -x -q -s -k "against and 4" 2>&1 | grep -v "Will use" | tail -4
