#!/bin/bash
# round-2 GPU call 36 (1 GPU): last check of the committed tree -- smoke, the whole GPU suite, default bench, reference arm
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/r2_g36_smoke.log 2>&1; tail -2 $O/r2_g36_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2_g36_pytest.log 2>&1; tail -3 $O/r2_g36_pytest.log
python bench.py > $O/r2_g36_bench.log 2> $O/r2_g36_bench.err; tail -1 $O/r2_g36_bench.log | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], 'traffic', d['roofline']['traffic'], 'clocks', d['clocks'])
print('keys', sorted(d.keys()))"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_g36_ref.log 2> $O/r2_g36_ref.err; tail -1 $O/r2_g36_ref.log | cut -c1-400
