#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_device_api.py -x -q -k "producer_loop" 2>&1 | tail -40 > $O/r2_g11_loop.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py -q 2>&1 | tail -15 > $O/r2_g11_pytest.log
B="python bench.py --no-cpu --no-extras --warmup 3 --algo pad"
$B --steps 200 --frames 1 > $O/r2_g11_f1.log 2>&1
$B --steps 10 --frames 128 > $O/r2_g11_f128.log 2>&1
$B --steps 50 --frames 16 > $O/r2_g11_f16.log 2>&1
for f in $O/r2_g11_f*.log; do echo "== $f $(tail -1 $f | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t); print('value %.0f  kernel_ms %.4f  fp32 %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['fp32_frac_of_148x128_lanes'] or 0))
except Exception as e: print(t[-300:])
")"; done
cat $O/r2_g11_loop.log; cat $O/r2_g11_pytest.log
