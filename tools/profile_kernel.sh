#!/bin/bash
# GPU box: ncu --set full capture of one kernel of a bench.py configuration, summarised on the box.
# usage: tools/profile_kernel.sh <tag> <kernel regex> <units per launch> <name> [bench.py args...]
#   -> gpurun_out/<tag>_<name>.md / .json / _lines.txt
tag=$1; rx=$2; units=$3; name=$4; shift 4
O=gpurun_out; mkdir -p $O
H="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-tick --no-extra $*"
$H > $O/${tag}_plain_$name.log 2>&1 || { echo "plain failed"; tail -5 $O/${tag}_plain_$name.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 -f -o $O/${tag}_prof_$name $H > $O/${tag}_ncu_$name.log 2>&1
python tools/ncu_summary.py kernel $O/${tag}_prof_$name.ncu-rep $O/${tag}_$name.md $O/${tag}_$name.json $units > /dev/null
python tools/ncu_lines.py $O/${tag}_prof_$name.ncu-rep $units 4 > $O/${tag}_${name}_lines.txt 2>&1
cat $O/${tag}_$name.md | head -40
[ -z "$KEEP_REPS" ] && rm -f $O/${tag}_prof_$name.ncu-rep
