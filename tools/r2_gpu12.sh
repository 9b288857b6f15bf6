#!/bin/bash
# round-2 GPU call 12 (8 GPUs): scaling run at N=8 with two flag depths; extras once
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2_g12_gpus.log; free -g | head -2 >> $O/r2_g12_gpus.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for depth in 2 4; do
  BF_GATHER_DEPTH=$depth timeout 600 $T --master-port 2951$depth bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > $O/r2_g12_n8_d$depth.log 2> $O/r2_g12_n8_d$depth.err
done
BF_C5_MINUTES=2 timeout 900 $T --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2_g12_n8_full.log 2> $O/r2_g12_n8_full.err
for f in $O/r2_g12_n8_d2.log $O/r2_g12_n8_d4.log $O/r2_g12_n8_full.log; do echo "== $f"; tail -1 $f | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t)
    print('value %.0f ms/step %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']), d['roofline']['per_rank_kernel_ms'], d['gather_check'])
    if d.get('e2e'): print('e2e', d['e2e']['value'], 'sharded', d['e2e'].get('sharded'))
    if d.get('mvdr'): print('mvdr sharded', d['mvdr'].get('sharded'))
    if d.get('replay'): print('replay', json.dumps(d['replay'])[:1500])
except Exception as e: print('ERR', e, t[-500:])
"; done
tail -3 $O/r2_g12_n8_full.err
