#!/bin/bash
# round-2 GPU call 31 (2 GPUs): padded row stride of the gather buffers (ragged slices), two-rank tests + N=2 bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -x -q 2>&1 | tail -6 > $O/r2_g31_pytest.log; cat $O/r2_g31_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
BF_C5_MINUTES=0.5 BF_C5_STREAM_MINUTES=2 timeout 900 $T --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_g31_n2.log 2> $O/r2_g31_n2.err
tail -1 $O/r2_g31_n2.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline'].get('per_rank_kernel_ms'), d['gather_check'], d.get('shard_weights'))
print('e2e', d['e2e']['value'], 'sharded', (d['e2e'].get('sharded') or {}).get('value'), (d['e2e'].get('sharded') or {}).get('host_maps_bit_exact_vs_one_gpu'))
print('mvdr sharded', d['mvdr'].get('sharded'))"
tail -3 $O/r2_g31_n2.err
