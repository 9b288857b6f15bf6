#!/bin/bash
# round-2 GPU call 41: second half of the consumer warps one chunk behind the first (epilogues no longer coincide)
# (no change measured; BF_MIMO_STAGGER no longer exists in the library)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
show() { tail -1 $1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['roofline'].get('kernel_ms'), d['roofline'].get('fp32_frac_of_148x128_lanes'))"; }
for st in 0 1 0 1; do BF_MIMO_STAGGER=$st $B --algo pad > $O/r2_g41_pad_s$st.log 2>&1; show $O/r2_g41_pad_s$st.log; done
for st in 0 1; do BF_MIMO_STAGGER=$st $B --algo lerp > $O/r2_g41_lerp_s$st.log 2>&1; show $O/r2_g41_lerp_s$st.log; done
BF_MIMO_STAGGER=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
