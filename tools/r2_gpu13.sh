#!/bin/bash
# round-2 GPU call 13 (2 GPUs): weighted sharding + e2e with input all-gather, quick
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
BF_C5_MINUTES=0.5 BF_C5_STREAM_MINUTES=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_g13_bench_n2.log 2> $O/r2_g13_bench_n2.err
tail -1 $O/r2_g13_bench_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms/step',d['ms_per_step'], d['gather_check'], d['shard_weights'])
print('e2e',d['e2e']['value'], 'sharded', d['e2e'].get('sharded'))
print('latency', d.get('latency'))
"
tail -5 $O/r2_g13_bench_n2.err
