#!/bin/bash
# round-2 GPU call 32: producer warp parked on its ring waits (try_wait with a suspend-time hint) instead of spinning
# (no change measured; BF_MIMO_PARK_NS no longer exists in the library)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
show() { tail -1 $1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline'].get('fp32_frac_of_148x128_lanes'))"; }
for rep in 1 2; do
for ns in 0 500 2000 8000; do
  for a in pad lerp; do BF_MIMO_PARK_NS=$ns $B --algo $a > $O/r2_g32_${a}_ns$ns.log 2>&1; show $O/r2_g32_${a}_ns$ns.log; done
done
done
BF_MIMO_PARK_NS=2000 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
