#!/bin/bash
# round-2 A/B #3: carry + fastbr as the default schedule for pad / lerp / shared; parity + ncu
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -s 2>&1 | grep -v "Will use" | tail -15 > $O/r2_ab3_pytest.log
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
$B --algo pad > $O/r2_ab3_pad.log 2>&1
$B --algo pad --exact-sum 0 > $O/r2_ab3_pad_tree.log 2>&1
$B --algo pad --exact-sum 2 > $O/r2_ab3_pad_shared.log 2>&1
$B --algo lerp > $O/r2_ab3_lerp.log 2>&1
$B --algo lerp --exact-sum 2 > $O/r2_ab3_lerp_shared.log 2>&1
BF_MIMO_WARPS_LERP=19 $B --algo lerp > $O/r2_ab3_lerp19.log 2>&1
for f in $O/r2_ab3_*.log; do echo "== $f"; tail -1 $f | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t)
    print('value %.0f  kernel_ms %.3f  fp32 %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['fp32_frac_of_148x128_lanes'] or 0))
except Exception as e: print(t[-600:])
"; done
python bench.py --no-cpu --no-extras --steps 3 --warmup 3 --algo pad > $O/r2_ab3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 4 -c 1 -f -o $O/r2_mimo_pad_vm2_F128 \
    python bench.py --no-cpu --no-extras --steps 3 --warmup 3 --algo pad > $O/r2_ab3_ncu.log 2>&1
echo done
