#!/bin/bash
# round-2 GPU call 15: quarter-tile MVDR variant (BF_MVDR_TC=4) parity + timing
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_mvdr.py -x -q -s -k "against" 2>&1 | grep -v "Will use" | tail -8 > $O/r2_g15_mvdr.log
for tc in 3 4; do
  timeout 300 python tools/mvdr_c4.py --tc $tc > $O/r2_g15_c4_tc$tc.log 2>&1
  echo "tc $tc: $(grep -o '"steering": [0-9.]*' $O/r2_g15_c4_tc$tc.log) $(grep -o '"power_max": [0-9.]*' $O/r2_g15_c4_tc$tc.log)"
done
BF_MVDR_TC=4 timeout 600 python -m pytest tests/test_gpu_c4_size.py -x -q -s -k mvdr 2>&1 | grep -v "Will use" | tail -4 > $O/r2_g15_c4size.log
cat $O/r2_g15_mvdr.log $O/r2_g15_c4size.log
