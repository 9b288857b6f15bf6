#!/bin/bash
# round-2 GPU call 6: what bounds the f16 MVDR steering kernel? (timing knobs + one ncu capture)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for dbg in 0 1 2 3; do
  BF_MVDR_DBG=$dbg timeout 300 python tools/mvdr_c4.py --tc 3 > $O/r2_g6_dbg$dbg.log 2>&1
  echo "dbg $dbg: $(grep -o '"steering": [0-9.]*' $O/r2_g6_dbg$dbg.log)"
done
timeout 300 python tools/mvdr_c4.py --tc 3 --bins 32 --reps 1 > $O/r2_g6_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mvdr_tc_steer_kernel3 -s 1 -c 1 -f -o $O/r2_mvdr_tc3_32bins \
    python tools/mvdr_c4.py --tc 3 --bins 32 --reps 1 > $O/r2_g6_ncu.log 2>&1
echo done
