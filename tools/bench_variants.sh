#!/bin/bash
# Tuning helper: bench.py over alternative builds of the library (NSB200_LIB) -- kernel-only numbers.
# usage: tools/bench_variants.sh "<bench args>" name1 name2 ...   (audiosignalprocess_b200/variants/lib<name>.so)
args="$1"; shift
for v in "$@"; do
  r=$(NSB200_LIB=$PWD/audiosignalprocess_b200/variants/lib$v.so python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu $args 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f ms/step %.4g audio-s/s'%(d['ms_per_step'],d['value']))" 2>&1 | tail -1)
  echo "$v [$args]: $r"
done
