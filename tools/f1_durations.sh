#!/bin/bash
# Tuning helper (GPU box): GPU-side kernel durations (ncu launch list) of short launches, per library variant.
# usage: tools/f1_durations.sh "<variants>" "<frames list>" "<streams list>"
for v in $1; do for F in $2; do for n in $3; do
  lib=$PWD/audiosignalprocess_b200/variants/lib$v.so; [ $v = main ] && lib=$PWD/audiosignalprocess_b200/libwebrtc_ns_b200.so
  NSB200_LIB=$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nsf_process -s 100 -c 8 --csv --log-file /tmp/d.csv python bench.py --steps 120 --warmup 3 --no-e2e --no-cpu --frames-per-step $F --streams $n > /dev/null 2>&1
  python - <<PY
import csv
v=[float(r[-1]) for r in csv.reader(open('/tmp/d.csv')) if r and r[-1].replace('.','').isdigit()]
v.sort(); m=v[len(v)//2]/1e3
print("$v F=$F n=$n: median %.1f us per launch = %.2f ns per stream-frame" % (m, m*1e3/($F*$n)))
PY
done; done; done
