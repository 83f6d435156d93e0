#!/bin/bash
# GPU box: ncu capture of the float kernel at F=100 (headline) with the per-line table.  usage: tools/profile_float.sh <tag>
tag=${1:-rX}; O=gpurun_out; mkdir -p $O
H="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-tick"
$H > $O/${tag}_plain_f16.log 2>&1 || { echo "plain failed"; tail -5 $O/${tag}_plain_f16.log; exit 1; }
tail -c 600 $O/${tag}_plain_f16.log
ncu --set full --clock-control none --import-source on -k regex:nsf_process -s 4 -c 1 -f -o $O/${tag}_prof_float $H > $O/${tag}_ncu_f.log 2>&1
python tools/ncu_summary.py kernel $O/${tag}_prof_float.ncu-rep $O/${tag}_nsf_kernel_F100.md $O/${tag}_nsf_kernel_F100.json 409600 > /dev/null
python tools/ncu_lines.py $O/${tag}_prof_float.ncu-rep 409600 4 > $O/${tag}_nsf_kernel_F100_lines.txt 2>&1
cat $O/${tag}_nsf_kernel_F100.md | head -60
python tools/ncu_opcodes.py $O/${tag}_prof_float.ncu-rep 409600 > $O/${tag}_nsf_kernel_F100_opcodes.txt 2>&1
[ -z "$KEEP_REPS" ] && rm -f $O/${tag}_prof_float.ncu-rep
ls -la $O | grep ${tag}_
