#!/bin/bash
# round-2 GPU calls 28 and 33 (1 GPU): whole GPU suite, default bench line, ncu re-captures of das_mimo (sources changed: tile
# hand-out + overlapping steps) and the launch list of the bench command
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2_g28_pytest.log 2>&1; tail -4 $O/r2_g28_pytest.log
python bench.py --steps 20 --warmup 5 > $O/r2_g28_bench.log 2> $O/r2_g28_bench.err; tail -1 $O/r2_g28_bench.log | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
print('roofline', d['roofline'])
print('mvdr', d['mvdr']['ms_per_map'], d['mvdr']['stage_ms'], d['mvdr']['roofline']['frac'])
print('latency', d.get('latency'))
print('replay', d['replay'].get('frames_per_s'), d['replay'].get('one_hour_stream'))
"
B="python bench.py --no-cpu --no-extras --steps 3 --warmup 3"
$B --algo pad > $O/r2_g28_plain_pad.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 4 -c 1 -f -o $O/r2g_mimo_pad_F128 $B --algo pad > $O/r2_g28_ncu1.log 2>&1
$B --algo lerp > $O/r2_g28_plain_lerp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 4 -c 1 -f -o $O/r2g_mimo_lerp_F128 $B --algo lerp > $O/r2_g28_ncu2.log 2>&1
python tools/gather_single.py > $O/r2_g28_plain_gather.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 2 -c 1 -f -o $O/r2g_mimo_gather_slice8 python tools/gather_single.py > $O/r2_g28_ncu3.log 2>&1
BF_C5_MINUTES=0.5 BF_C5_STREAM_MINUTES=1 python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_g28_plain_bench.log 2>&1 && \
BF_C5_MINUTES=0.5 BF_C5_STREAM_MINUTES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_g28_ncu5.log 2>&1
ls -la $O/r2g_*
