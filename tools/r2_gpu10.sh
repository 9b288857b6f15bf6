#!/bin/bash
# round-2 GPU call 10: new GPU tests (stream replay, slices), consumer-warp sweep at F=128 and F=1, short bench with the stream leg
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
free -g | head -2 > $O/r2_g10_mem.log; nproc >> $O/r2_g10_mem.log
timeout 600 python -m pytest tests/test_gpu_device_api.py tests/test_gpu_mvdr.py -x -q 2>&1 | tail -6 > $O/r2_g10_pytest.log
B="python bench.py --no-cpu --no-extras --warmup 3 --algo pad"
for w in 12 14 15 16 19; do
  BF_MIMO_WARPS=$w $B --steps 10 --frames 128 > $O/r2_g10_w${w}_f128.log 2>&1
  BF_MIMO_WARPS=$w $B --steps 200 --frames 1 > $O/r2_g10_w${w}_f1.log 2>&1
  BF_MIMO_WARPS=$w $B --steps 50 --frames 16 > $O/r2_g10_w${w}_f16.log 2>&1
done
for f in $O/r2_g10_w*.log; do echo "== $f $(tail -1 $f | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t); print('value %.0f  kernel_ms %.4f  fp32 %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['fp32_frac_of_148x128_lanes'] or 0))
except Exception as e: print(t[-300:])
")"; done
BF_C5_MINUTES=2 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r2_g10_bench.log 2> $O/r2_g10_bench.err
tail -6 $O/r2_g10_pytest.log; cat $O/r2_g10_mem.log
tail -1 $O/r2_g10_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps(d['replay'])[:3000]); print('e2e', d['e2e']['value'], 'value', d['value'])"
tail -3 $O/r2_g10_bench.err
