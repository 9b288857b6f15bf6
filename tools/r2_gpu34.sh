#!/bin/bash
# round-2 GPU call 34 (2 GPUs): PeerInput (copy-engine input all-gather) -- two-rank test, then e2e.sharded p2p vs nccl
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_peer_gather.py -x -q 2>&1 | tail -8 > $O/r2_g34_pytest.log; cat $O/r2_g34_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in p2p nccl; do
BF_E2E_INPUT=$mode BF_C5_MINUTES=0.2 BF_C5_STREAM_MINUTES=0.5 timeout 900 $T --master-port 2958$((RANDOM % 10)) bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_g34_n2_$mode.log 2> $O/r2_g34_n2_$mode.err
tail -1 $O/r2_g34_n2_$mode.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$mode', round(d['value']), d['gather_check'])
print('  e2e', d['e2e']['value'], 'sharded', d['e2e'].get('sharded'))"
done
tail -3 $O/r2_g34_n2_p2p.err
