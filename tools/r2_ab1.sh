#!/bin/bash
# round-2 A/B of the das_mimo microphone loop (interpreter vs legacy), one GPU
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -s 2>&1 | tail -25 > $O/r2_ab1_pytest.log
B="python bench.py --no-cpu --no-extras --steps 10 --warmup 3"
$B --algo pad > $O/r2_ab1_pad_vm.log 2>&1
BF_MIMO_VM=0 $B --algo pad > $O/r2_ab1_pad_legacy.log 2>&1
$B --algo pad --exact-sum 0 > $O/r2_ab1_pad_vm_tree.log 2>&1
$B --algo pad --exact-sum 2 > $O/r2_ab1_pad_shared.log 2>&1
$B --algo lerp > $O/r2_ab1_lerp_vm15.log 2>&1
BF_MIMO_WARPS_LERP=19 $B --algo lerp > $O/r2_ab1_lerp_vm19.log 2>&1
BF_MIMO_VM=0 $B --algo lerp > $O/r2_ab1_lerp_legacy.log 2>&1
$B --algo lerp --exact-sum 2 > $O/r2_ab1_lerp_shared.log 2>&1
BF_MIMO_WARPS_LERP=19 $B --algo lerp --exact-sum 2 > $O/r2_ab1_lerp_shared19.log 2>&1
for f in $O/r2_ab1_*.log; do echo "== $f"; tail -c 1500 $f | grep -o '"value": [0-9.]*\|fp32_frac_of_148x128_lanes": [0-9.]*\|kernel_ms": [0-9.]*\|passed\|failed\|rel err [^(]*' | tr '\n' ' '; echo; done
python bench.py --no-cpu --no-extras --steps 3 --warmup 3 --algo pad > $O/r2_ab1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 4 -c 1 -f -o $O/r2_mimo_pad_vm_F128 \
    python bench.py --no-cpu --no-extras --steps 3 --warmup 3 --algo pad > $O/r2_ab1_ncu.log 2>&1
echo done
