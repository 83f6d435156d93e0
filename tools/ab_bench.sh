#!/bin/bash
# Tuning helper (GPU box): A/B of library variants on bench.py configurations (plain CUDA-event timings).
# usage: tools/ab_bench.sh "<variants: main head ...>" "<F:streams:steps[:--fixed] ...>"
for cfg in $2; do
  IFS=: read F n steps extra <<< "$cfg"
  for v in $1; do
    lib=$PWD/audiosignalprocess_b200/variants/lib$v.so; [ $v = main ] && lib=$PWD/audiosignalprocess_b200/libwebrtc_ns_b200.so
    NSB200_LIB=$lib timeout 300 python bench.py --steps $steps --warmup 5 --no-e2e --no-cpu --no-extra --no-tick --frames-per-step $F --streams $n $extra 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$v F=$F n=$n $extra: %.1f us/step  %.3e audio-s/s  hbm frac %.3f  (median launch %.1f us, clocks %s MHz)' % (
            d['ms_per_step'] * 1e3, d['value'], d['roofline']['frac'], d['roofline']['launch_ms_median'] * 1e3, d['clocks'].get('sm_mhz')))
"
  done
done
