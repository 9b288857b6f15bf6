#!/bin/bash
# round-2 GPU call 35 (8 GPUs): e2e.sharded with the copy-engine input all-gather (PeerInput) against the NCCL one
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
p=29590
for mode in p2p nccl p2p; do
p=$((p+1))
BF_EXTRAS=e2e BF_E2E_INPUT=$mode timeout 600 $T --master-port $p bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_g35_n8_$mode.log 2> $O/r2_g35_n8_$mode.err
tail -1 $O/r2_g35_n8_$mode.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$mode', round(d['value']), d['gather_check'])
sh=d['e2e'].get('sharded') or {}
print('  e2e', d['e2e']['value'], 'sharded', sh.get('value'), sh.get('host_maps_bit_exact_vs_one_gpu'), sh.get('input_gather'), sh.get('error'))"
done
tail -3 $O/r2_g35_n8_p2p.err
