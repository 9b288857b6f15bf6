#!/bin/bash
# round-2 GPU call 24: overlapping gather steps (programmatic stream serialisation) on one GPU + the two-mode tile hand-out
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py tests/test_gpu_peer_gather.py -x -q > $O/r2_g24b_pytest.log 2>&1; tail -3 $O/r2_g24b_pytest.log
for sl in 8 4 1; do
  for f in 128 16; do
    timeout 300 python tools/gather_single.py --slice $sl --frames $f --run 40 --overlap 1 > $O/r2_g24b_run_s${sl}_F$f.log 2>&1; tail -1 $O/r2_g24b_run_s${sl}_F$f.log
  done
done
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
show() { tail -1 $1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline'].get('fp32_frac_of_148x128_lanes'))"; }
for a in pad lerp; do $B --algo $a > $O/r2_g24b_$a.log 2>&1; show $O/r2_g24b_$a.log; done
BF_MIMO_SPLIT=0 $B --algo pad > $O/r2_g24b_pad_s0.log 2>&1; show $O/r2_g24b_pad_s0.log
for f in 16 1; do $B --algo pad --frames $f --steps 40 > $O/r2_g24b_pad_F$f.log 2>&1; show $O/r2_g24b_pad_F$f.log; done
$B --exact-sum 2 > $O/r2_g24b_shared.log 2>&1; show $O/r2_g24b_shared.log
