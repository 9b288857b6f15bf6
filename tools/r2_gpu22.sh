#!/bin/bash
# round-2 GPU call 22: balanced unit ranges per CTA (TileWalk) -- parity, then throughput at F = 128 / 16 / 1, a 1/8 slice, W sweep
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py tests/test_gpu_peer_gather.py -x -q > $O/r2_g22b_pytest.log 2>&1; tail -3 $O/r2_g22b_pytest.log
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
show() { tail -1 $1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline'].get('fp32_frac_of_148x128_lanes'))"; }
for a in pad lerp; do $B --algo $a > $O/r2_g22b_$a.log 2>&1; show $O/r2_g22b_$a.log; done
for w in 16 15 12; do BF_MIMO_WARPS=$w $B --algo pad > $O/r2_g22b_pad_w$w.log 2>&1; show $O/r2_g22b_pad_w$w.log; done
for f in 16 1; do
  for w in 19 16 15 12; do BF_MIMO_WARPS=$w $B --algo pad --frames $f --steps 40 > $O/r2_g22b_pad_F${f}_w$w.log 2>&1; show $O/r2_g22b_pad_F${f}_w$w.log; done
done
for w in 19 16; do BF_MIMO_WARPS=$w python tools/gather_single.py > $O/r2_g22b_gather_w$w.log 2>&1; tail -1 $O/r2_g22b_gather_w$w.log; done
