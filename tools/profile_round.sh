#!/bin/bash
# Runs on the GPU box (under gpurun): every ncu pass is preceded by the same command without ncu.
# usage: tools/profile_round.sh <tag>      -> gpurun_out/<tag>_*
tag=${1:-rX}; O=gpurun_out; mkdir -p $O
H="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-tick --no-extra"
run() { name=$1; shift; "$@" > $O/${tag}_plain_$name.log 2>&1 || { echo "plain $name failed"; tail -3 $O/${tag}_plain_$name.log; return 1; }; }
# 1. headline launch list
run f16 $H && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${tag}_launches_float16k.csv $H > $O/${tag}_ncu_l16.log 2>&1
# 2. full captures: float F=100, float F=1, fixed F=100
run f16 $H && ncu --set full --clock-control none --import-source on -k regex:nsf_process -s 4 -c 1 -f -o $O/${tag}_prof_float $H > $O/${tag}_ncu_f.log 2>&1
# F = 1 (the 10 ms tick) in steady state: launch 290 is past the start-up regimes (50 / 200 frames)
T="python bench.py --warmup 3 --no-e2e --no-cpu --no-tick --no-extra --frames-per-step 1 --steps 300"
run f1 $T --streams 32768 && ncu --set full --clock-control none --import-source on -k regex:nsf_process -s 290 -c 1 -f -o $O/${tag}_prof_float_f1_32768 $T --streams 32768 > $O/${tag}_ncu_f1.log 2>&1
run f1s $T && ncu --set full --clock-control none --import-source on -k regex:nsf_process -s 290 -c 1 -f -o $O/${tag}_prof_float_f1 $T > $O/${tag}_ncu_f1s.log 2>&1
run x1 $T --streams 32768 --fixed && ncu --set full --clock-control none --import-source on -k regex:nsx_process -s 290 -c 1 -f -o $O/${tag}_prof_fixed_f1_32768 $T --streams 32768 --fixed > $O/${tag}_ncu_x1.log 2>&1
run x16 $H --fixed && ncu --set full --clock-control none --import-source on -k regex:nsx_process -s 4 -c 1 -f -o $O/${tag}_prof_fixed $H --fixed > $O/${tag}_ncu_x.log 2>&1
# Summaries are made here, on the box: the reports together exceed what gpurun copies back (64 MiB).
summ() { [ -f $O/${tag}_prof_$1.ncu-rep ] && python tools/ncu_summary.py kernel $O/${tag}_prof_$1.ncu-rep $O/${tag}_$2.md $O/${tag}_$2.json $3 > /dev/null; }
summ float nsf_kernel_F100 409600; summ float_f1_32768 nsf_kernel_F1_32768 32768; summ float_f1 nsf_kernel_F1 4096
summ fixed nsx_kernel_F100 409600; summ fixed_f1_32768 nsx_kernel_F1_32768 32768
python tools/ncu_lines.py $O/${tag}_prof_float.ncu-rep 409600 > $O/${tag}_nsf_kernel_F100_lines.txt 2>&1
python tools/ncu_lines.py $O/${tag}_prof_float_f1_32768.ncu-rep 32768 > $O/${tag}_nsf_kernel_F1_32768_lines.txt 2>&1
python tools/ncu_summary.py launches $O/${tag}_launches_float16k.csv $O/${tag}_launches_float16k.md > /dev/null
[ -z "$KEEP_REPS" ] && rm -f $O/${tag}_prof_fixed*.ncu-rep $O/${tag}_prof_float_f1.ncu-rep $O/${tag}_prof_float_f1_32768.ncu-rep
[ -n "$SKIP_48K" ] && { ls -la $O | grep ${tag}_; exit 0; }
# 3. 48 kHz: launch list + band kernels
B="python bench.py --fs 48000 --streams 2048 --frames-per-step 50 --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra"
run f48 $B && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${tag}_launches_48k.csv $B > $O/${tag}_ncu_l48.log 2>&1
run f48 $B && NSB200_BAND_CHUNK=50 ncu --set full --clock-control none --import-source on -k regex:"qmf_|resample_" -s 14 -c 6 -f -o $O/${tag}_prof_bands $B > $O/${tag}_ncu_b.log 2>&1
ls -la $O | grep ${tag}_
