#!/bin/bash
# round-2 GPU call 30 (8 GPUs): weighted slices from the rounding-free speed probe vs equal slices, overlapping steps on
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
show() { tail -1 $1 | python -c "
import sys, json
t=sys.stdin.read()
try:
    d=json.loads(t)
    print('$1', 'value %.0f ms/step %.4f' % (d['value'], d['ms_per_step']), d['roofline']['per_rank_kernel_ms'], d['gather_check'], d.get('shard_weights'))
except Exception as e: print('ERR', e, t[-500:])
"; }
timeout 600 $T --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > $O/r2_g30_n8_weighted.log 2> $O/r2_g30_n8_weighted.err; show $O/r2_g30_n8_weighted.log
BF_SHARD_WEIGHTED=0 timeout 600 $T --master-port 29562 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > $O/r2_g30_n8_equal.log 2> $O/r2_g30_n8_equal.err; show $O/r2_g30_n8_equal.log
timeout 600 $T --master-port 29563 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > $O/r2_g30_n8_weighted2.log 2> $O/r2_g30_n8_weighted2.err; show $O/r2_g30_n8_weighted2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu > $O/r2_g30_n1.log 2> $O/r2_g30_n1.err; tail -1 $O/r2_g30_n1.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('n1', round(d['value']), d['ms_per_step'])"
tail -3 $O/r2_g30_n8_weighted.err
