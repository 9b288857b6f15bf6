#!/bin/bash
# round-2 GPU call 40 (2 GPUs): two-rank tests + a short N=2 bench after the gather_layout refactor
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > $O/r2_g40_n2.log 2> $O/r2_g40_n2.err
tail -1 $O/r2_g40_n2.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(round(d['value']), d['gather_check'], d.get('shard_weights'))"
