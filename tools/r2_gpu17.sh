#!/bin/bash
# round-2 GPU call 17: what the driver runs at round end -- GPU suite, smoke, reference arm, default bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2_g17_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > $O/r2_g17_smoke.log
timeout 900 python bench.py --impl reference > $O/r2_g17_ref.log 2> $O/r2_g17_ref.err
timeout 1200 python bench.py > $O/r2_g17_bench.log 2> $O/r2_g17_bench.err
cat $O/r2_g17_pytest.log $O/r2_g17_smoke.log
tail -1 $O/r2_g17_ref.log | cut -c1-900
tail -1 $O/r2_g17_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
keep={k:d[k] for k in ('metric','value','unit','n_gpus','steps','ms_per_step','gpu_launches','clocks','roofline','cpu_baseline','shared_sum_mode','latency')}
print(json.dumps(keep)[:3500])
print('e2e', json.dumps(d['e2e'])[:800])
print('mvdr', json.dumps(d['mvdr'])[:1200])
print('miso', json.dumps(d['miso'])[:600])
print('replay', json.dumps(d['replay'])[:1500])
"
tail -3 $O/r2_g17_bench.err
