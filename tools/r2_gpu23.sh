#!/bin/bash
# round-2 GPU call 23: tile hand-out A/B in one binary: BF_MIMO_SPLIT = 0 whole tiles (round 1), 1 frame-major unit ranges,
# 2 full rounds + last round split into equal pieces
# (mode 2 gained nothing and was removed after this call: the shipped library knows BF_MIMO_SPLIT = 0 / 1 only;
# results in profiles/r2_kernel_variants.md)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py tests/test_gpu_peer_gather.py -x -q > $O/r2_g23b_pytest.log 2>&1; tail -3 $O/r2_g23b_pytest.log
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
show() { tail -1 $1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline'].get('fp32_frac_of_148x128_lanes'))"; }
for rep in 1 2; do
for m in 0 1 2; do
  export BF_MIMO_SPLIT=$m
  for a in pad lerp; do $B --algo $a > $O/r2_g23b_${a}_s$m.log 2>&1; show $O/r2_g23b_${a}_s$m.log; done
done
done
for m in 0 1 2; do
  export BF_MIMO_SPLIT=$m
  for f in 16 1; do
    for w in 19 16 15; do BF_MIMO_WARPS=$w $B --algo pad --frames $f --steps 40 > $O/r2_g23b_pad_F${f}_w${w}_s$m.log 2>&1; show $O/r2_g23b_pad_F${f}_w${w}_s$m.log; done
  done
  $B --algo lerp --frames 16 --steps 40 > $O/r2_g23b_lerp_F16_s$m.log 2>&1; show $O/r2_g23b_lerp_F16_s$m.log
  for w in 19 16; do BF_MIMO_WARPS=$w python tools/gather_single.py > $O/r2_g23b_gather_w${w}_s$m.log 2>&1; echo "slice8 w$w s$m $(tail -1 $O/r2_g23b_gather_w${w}_s$m.log)"; done
done
