#!/bin/bash
# round-2 GPU call 20: final ncu captures of the shipped kernels + launch list of the bench command
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu --no-extras --steps 3 --warmup 3"
$B --algo pad > $O/r2_g20_plain_pad.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 4 -c 1 -f -o $O/r2f_mimo_pad_F128 $B --algo pad > $O/r2_g20_ncu1.log 2>&1
$B --algo lerp > $O/r2_g20_plain_lerp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 4 -c 1 -f -o $O/r2f_mimo_lerp_F128 $B --algo lerp > $O/r2_g20_ncu2.log 2>&1
python tools/gather_single.py > $O/r2_g20_plain_gather.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:das_mimo -s 2 -c 1 -f -o $O/r2f_mimo_gather_slice8 python tools/gather_single.py > $O/r2_g20_ncu3.log 2>&1
python tools/mvdr_c4.py --bins 32 --reps 1 > $O/r2_g20_plain_mvdr.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mvdr_tc_steer_kernel3 -s 1 -c 1 -f -o $O/r2f_mvdr_tc4_32bins python tools/mvdr_c4.py --bins 32 --reps 1 > $O/r2_g20_ncu4.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_g20_plain_bench.log 2>&1 && \
BF_C5_MINUTES=0.5 BF_C5_STREAM_MINUTES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_g20_ncu5.log 2>&1
BF_C5_MINUTES=0.5 BF_C5_STREAM_MINUTES=1 ncu --set full --clock-control none -k regex:miso_stream_kernel -c 16 -f -o $O/r2f_miso_stream python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_g20_ncu6.log 2>&1
cat $O/r2_g20_plain_gather.log | tail -1
ls -la $O/r2f_*
