#!/bin/bash
# round-2 A/B #4: weave schedule (pad), lerp schedules x warps
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
BF_MIMO_VM=3 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > $O/r2_ab4_pytest_x1.log
BF_MIMO_VM=5 BF_MIMO_WARPS_LERP=19 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > $O/r2_ab4_pytest_x3_19.log
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
BF_MIMO_VM=3 $B --algo pad > $O/r2_ab4_pad_x1.log 2>&1
for vm in 1 3 4 5; do for w in 15 19; do
  BF_MIMO_VM=$vm BF_MIMO_WARPS_LERP=$w $B --algo lerp > $O/r2_ab4_lerp_vm${vm}_w${w}.log 2>&1
done; done
for f in $O/r2_ab4_*.log; do echo "== $f"; tail -1 $f | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t)
    print('value %.0f  kernel_ms %.3f  fp32 %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['fp32_frac_of_148x128_lanes'] or 0))
except Exception as e: print(t[-600:])
"; done
