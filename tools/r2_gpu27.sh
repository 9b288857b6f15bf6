#!/bin/bash
# round-2 GPU call 27 (8 GPUs): overlapping steps + rendezvous at N=8 against the plain launch order; full line once; N=4
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
show() { tail -1 $1 | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t)
    print('$1', 'value %.0f ms/step %.4f' % (d['value'], d['ms_per_step']), d['roofline']['per_rank_kernel_ms'], d['gather_check'])
    if d.get('e2e'): print('  e2e', d['e2e']['value'], 'sharded', (d['e2e'].get('sharded') or {}).get('value'), (d['e2e'].get('sharded') or {}).get('host_maps_bit_exact_vs_one_gpu'))
    if d.get('mvdr'): print('  mvdr sharded', d['mvdr'].get('sharded'))
    if d.get('replay'): print('  replay', d['replay'].get('frames_per_s'))
except Exception as e: print('ERR', e, t[-500:])
"; }
BF_GATHER_OVERLAP=0 BF_RENDEZVOUS=0 timeout 600 $T --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > $O/r2_g27_n8_plain.log 2> $O/r2_g27_n8_plain.err; show $O/r2_g27_n8_plain.log
timeout 600 $T --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > $O/r2_g27_n8_overlap.log 2> $O/r2_g27_n8_overlap.err; show $O/r2_g27_n8_overlap.log
BF_C5_MINUTES=2 timeout 900 $T --master-port 29553 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_g27_n8_full.log 2> $O/r2_g27_n8_full.err; show $O/r2_g27_n8_full.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > $O/r2_g27_n4.log 2> $O/r2_g27_n4.err; show $O/r2_g27_n4.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu > $O/r2_g27_n1.log 2> $O/r2_g27_n1.err; tail -1 $O/r2_g27_n1.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('n1', round(d['value']), d['ms_per_step'])"
tail -3 $O/r2_g27_n8_full.err
