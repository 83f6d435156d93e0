#!/bin/bash
# Tuning helper: bench.py over alternative builds of the library (NSB200_LIB) -- kernel-only numbers.
for v in "$@"; do
  for bal in 1 0; do
    for kind in "" "--fixed"; do
      if [ $bal = 0 ]; then export NSB200_NO_BALANCE=1; else unset NSB200_NO_BALANCE; fi
      r=$(NSB200_LIB=$PWD/audiosignalprocess_b200/variants/lib$v.so python bench.py --steps 15 --warmup 3 --no-e2e --no-cpu $kind 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f ms/step %.4g audio-s/s'%(d['ms_per_step'],d['value']))")
      echo "$v balance=$bal ${kind:-float}: $r"
    done
  done
done
