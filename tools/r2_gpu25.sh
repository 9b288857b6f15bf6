#!/bin/bash
# round-2 GPU call 25 (2 GPUs): overlapping gather steps across ranks -- two-rank tests, N=2 bench with and without overlap
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -x -q 2>&1 | tail -6 > $O/r2_g25_pytest.log; cat $O/r2_g25_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
show() { tail -1 $1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['roofline'].get('per_rank_kernel_ms'), d['gather_check'], d['config']['parallelism'][:60])"; }
BF_GATHER_OVERLAP=0 timeout 600 $T --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > $O/r2_g25_n2_plain.log 2> $O/r2_g25_n2_plain.err; show $O/r2_g25_n2_plain.log
timeout 600 $T --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > $O/r2_g25_n2_overlap.log 2> $O/r2_g25_n2_overlap.err; show $O/r2_g25_n2_overlap.log
timeout 600 $T --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras --frames 16 > $O/r2_g25_n2_overlap_F16.log 2> $O/r2_g25_n2_overlap_F16.err; show $O/r2_g25_n2_overlap_F16.log
tail -3 $O/r2_g25_n2_overlap.err
