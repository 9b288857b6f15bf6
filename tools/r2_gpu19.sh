#!/bin/bash
# round-2 GPU call 19 (2 GPUs): bin-sharded MVDR parity + N=2 bench mvdr leg
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_peer_gather.py tests/test_gpu_mvdr.py tests/test_gpu_c4_size.py -x -q 2>&1 | tail -12 > $O/r2_g19_pytest.log
BF_C5_MINUTES=0.5 BF_C5_STREAM_MINUTES=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_g19_bench_n2.log 2> $O/r2_g19_bench_n2.err
cat $O/r2_g19_pytest.log
tail -1 $O/r2_g19_bench_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'], d['gather_check'])
print('mvdr', d['mvdr']['ms_per_map'], d['mvdr'].get('sharded'))
"
tail -4 $O/r2_g19_bench_n2.err
