#!/bin/bash
# round-2 GPU call 18 (8 GPUs): weighted slices vs equal slices; full line once
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
BF_SHARD_WEIGHTED=0 timeout 600 $T --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > $O/r2_g18_n8_equal.log 2> $O/r2_g18_n8_equal.err
timeout 600 $T --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > $O/r2_g18_n8_weighted.log 2> $O/r2_g18_n8_weighted.err
BF_C5_MINUTES=2 timeout 900 $T --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2_g18_n8_full.log 2> $O/r2_g18_n8_full.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 3 --no-extras > $O/r2_g18_n4.log 2> $O/r2_g18_n4.err
for f in $O/r2_g18_n8_equal.log $O/r2_g18_n8_weighted.log $O/r2_g18_n8_full.log $O/r2_g18_n4.log; do echo "== $f"; tail -1 $f | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t)
    print('value %.0f ms/step %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']), d['roofline']['per_rank_kernel_ms'], d['gather_check'])
    print('weights', d.get('shard_weights'))
    if d.get('e2e'): print('e2e', d['e2e']['value'], 'sharded', (d['e2e'].get('sharded') or {}).get('value'), (d['e2e'].get('sharded') or {}).get('host_maps_bit_exact_vs_one_gpu'))
    if d.get('mvdr'): print('mvdr sharded', d['mvdr'].get('sharded'))
    if d.get('replay'): print('replay', d['replay'].get('frames_per_s'))
except Exception as e: print('ERR', e, t[-500:])
"; done
tail -3 $O/r2_g18_n8_full.err
