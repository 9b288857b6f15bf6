#!/bin/bash
# round-2 A/B #2: schedules of the interpreter loop (carry / fastbr) x producer warp or not
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
BF_MIMO_VM=4 BF_MIMO_NOPROD=1 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > $O/r2_ab2_pytest_x2np.log
BF_MIMO_VM=3 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > $O/r2_ab2_pytest_x1.log
BF_MIMO_VM=5 BF_MIMO_NOPROD=1 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or random" 2>&1 | tail -3 > $O/r2_ab2_pytest_x3np.log
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3 --algo pad"
for vm in 1 3 4 5; do for np in 0 1; do
  BF_MIMO_VM=$vm BF_MIMO_NOPROD=$np $B > $O/r2_ab2_vm${vm}_np${np}.log 2>&1
done; done
for f in $O/r2_ab2_*.log; do echo "== $f"; tail -1 $f | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t)
    print('value %.0f  kernel_ms %.3f  fp32 %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['fp32_frac_of_148x128_lanes'] or 0))
except Exception as e: print(t[-300:])
"; done
