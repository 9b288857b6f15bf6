#!/bin/bash
# round-2 GPU call 29: uniform-pair handler (code 10: two uniform entries per dispatch) -- parity, then pairs on / off
# (the pair handler lost 3.6 % and is a generator option only: BF_MIMO_PAIRS no longer exists in the library)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py tests/test_gpu_peer_gather.py -x -q > $O/r2_g29_pytest.log 2>&1; tail -3 $O/r2_g29_pytest.log
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
show() { tail -1 $1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline'].get('fp32_frac_of_148x128_lanes'))"; }
for rep in 1 2; do
for pr in 0 1; do
  BF_MIMO_PAIRS=$pr $B --algo pad > $O/r2_g29_pad_p$pr.log 2>&1; show $O/r2_g29_pad_p$pr.log
done
done
for pr in 0 1; do
  BF_MIMO_PAIRS=$pr $B --algo pad --exact-sum 2 > $O/r2_g29_shared_p$pr.log 2>&1; show $O/r2_g29_shared_p$pr.log
  BF_MIMO_PAIRS=$pr $B --algo pad --frames 16 --steps 40 > $O/r2_g29_pad_F16_p$pr.log 2>&1; show $O/r2_g29_pad_F16_p$pr.log
  BF_MIMO_PAIRS=$pr $B --algo pad --frames 1 --steps 40 > $O/r2_g29_pad_F1_p$pr.log 2>&1; show $O/r2_g29_pad_F1_p$pr.log
done
$B --algo lerp > $O/r2_g29_lerp.log 2>&1; show $O/r2_g29_lerp.log
