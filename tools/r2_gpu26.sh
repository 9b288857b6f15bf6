#!/bin/bash
# round-2 GPU call 26 (2 GPUs): device-side rendezvous in front of the timed region (rank skew of the host barrier)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
show() { tail -1 $1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], d['roofline'].get('per_rank_kernel_ms'), d['gather_check'])"; }
p=29540
for f in 16 128; do
  for r in 0 1; do
    for st in 20 200; do
      p=$((p+1))
      BF_RENDEZVOUS=$r timeout 600 $T --master-port $p bench.py --gpus 2 --steps $st --warmup 5 --no-extras --frames $f > $O/r2_g26_F${f}_r${r}_s$st.log 2> $O/r2_g26_F${f}_r${r}_s$st.err; show $O/r2_g26_F${f}_r${r}_s$st.log
    done
  done
done
