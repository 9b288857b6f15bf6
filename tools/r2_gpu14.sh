#!/bin/bash
# round-2 GPU call 14: dual-role feeder experiment (pad, lerp) + parity under it; lerp ncu capture; FD test tolerance; full bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
BF_MIMO_DUAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4 > $O/r2_g14_pytest_dual.log
timeout 300 python -m pytest tests/test_gpu_fd.py -x -q 2>&1 | tail -6 > $O/r2_g14_pytest_fd.log
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
for dual in 0 1; do for algo in pad lerp; do
  BF_MIMO_DUAL=$dual $B --algo $algo > $O/r2_g14_${algo}_dual$dual.log 2>&1
done; done
BF_MIMO_DUAL=1 $B --algo pad --frames 1 --steps 200 > $O/r2_g14_pad_dual1_f1.log 2>&1
for f in $O/r2_g14_pad_*.log $O/r2_g14_lerp_*.log; do echo "== $f $(tail -1 $f | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t); print('value %.0f  kernel_ms %.4f  fp32 %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['fp32_frac_of_148x128_lanes'] or 0))
except Exception as e: print(t[-300:])
")"; done
cat $O/r2_g14_pytest_dual.log $O/r2_g14_pytest_fd.log
python bench.py --no-cpu --no-extras --steps 3 --warmup 3 --algo lerp > $O/r2_g14_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 4 -c 1 -f -o $O/r2_mimo_lerp_F128 \
    python bench.py --no-cpu --no-extras --steps 3 --warmup 3 --algo lerp > $O/r2_g14_ncu.log 2>&1
echo done
