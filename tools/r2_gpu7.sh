#!/bin/bash
# round-2 GPU call 7: MVDR v3 with converged MMA warp (uniform descriptors) + per-bin scale
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_mvdr.py -x -q -s -k "3 or singular" 2>&1 | grep -v "Will use" | tail -6 > $O/r2_g7_mvdr.log
timeout 300 python tools/mvdr_c4.py --tc 3 > $O/r2_g7_c4_tc3.log 2>&1
timeout 600 python -m pytest tests/test_gpu_c4_size.py -x -q -s 2>&1 | grep -v "Will use" | tail -8 > $O/r2_g7_c4size.log
for dbg in 1 2; do
  BF_MVDR_DBG=$dbg timeout 300 python tools/mvdr_c4.py --tc 3 > $O/r2_g7_dbg$dbg.log 2>&1
  echo "dbg $dbg: $(grep -o '"steering": [0-9.]*' $O/r2_g7_dbg$dbg.log)"
done
tail -n 6 $O/r2_g7_mvdr.log $O/r2_g7_c4_tc3.log $O/r2_g7_c4size.log
timeout 300 python tools/mvdr_c4.py --tc 3 --bins 32 --reps 1 > $O/r2_g7_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mvdr_tc_steer_kernel3 -s 1 -c 1 -f -o $O/r2_mvdr_tc3b_32bins \
    python tools/mvdr_c4.py --tc 3 --bins 32 --reps 1 > $O/r2_g7_ncu.log 2>&1
echo done
