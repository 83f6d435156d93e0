#!/bin/bash
# Runs on the GPU box (under gpurun): every ncu pass is preceded by the same command without ncu.
# usage: tools/profile_round.sh <tag>      -> gpurun_out/<tag>_*
tag=${1:-rX}; O=gpurun_out; mkdir -p $O
H="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu"
run() { name=$1; shift; "$@" > $O/${tag}_plain_$name.log 2>&1 || { echo "plain $name failed"; tail -3 $O/${tag}_plain_$name.log; return 1; }; }
# 1. headline launch list
run f16 $H && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${tag}_launches_float16k.csv $H > $O/${tag}_ncu_l16.log 2>&1
# 2. full captures: float F=100, float F=1, fixed F=100
run f16 $H && ncu --set full --clock-control none --import-source on -k regex:nsf_process -s 4 -c 1 -f -o $O/${tag}_prof_float $H > $O/${tag}_ncu_f.log 2>&1
run f1 $H --frames-per-step 1 --steps 50 && ncu --set full --clock-control none --import-source on -k regex:nsf_process -s 20 -c 1 -f -o $O/${tag}_prof_float_f1 $H --frames-per-step 1 --steps 50 > $O/${tag}_ncu_f1.log 2>&1
run x16 $H --fixed && ncu --set full --clock-control none --import-source on -k regex:nsx_process -s 4 -c 1 -f -o $O/${tag}_prof_fixed $H --fixed > $O/${tag}_ncu_x.log 2>&1
# 3. 48 kHz: launch list + band kernels
B="python bench.py --fs 48000 --streams 2048 --frames-per-step 50 --steps 3 --warmup 3 --no-e2e --no-cpu"
run f48 $B && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${tag}_launches_48k.csv $B > $O/${tag}_ncu_l48.log 2>&1
run f48 $B && NSB200_BAND_CHUNK=50 ncu --set full --clock-control none --import-source on -k regex:"qmf_|resample_" -s 14 -c 6 -f -o $O/${tag}_prof_bands $B > $O/${tag}_ncu_b.log 2>&1
ls -la $O | grep ${tag}_
