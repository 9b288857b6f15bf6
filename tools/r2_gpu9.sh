#!/bin/bash
# round-2 GPU call 9 (2 GPUs): fused gather + FD sharded parity, bench at N=2 with the sharded e2e and MVDR legs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_peer_gather.py tests/test_gpu_mvdr.py -x -q 2>&1 | tail -8 > $O/r2_g9_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_g9_bench_n2.log 2> $O/r2_g9_bench_n2.err
tail -8 $O/r2_g9_pytest.log
tail -c 2500 $O/r2_g9_bench_n2.log
tail -5 $O/r2_g9_bench_n2.err
