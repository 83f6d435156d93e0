#!/bin/bash
# All bench configurations of the README table (run on the GPU box): tools/bench_matrix.sh <tag>
tag=${1:-rX}; O=gpurun_out/${tag}_matrix.log; : > $O
run() { echo "== $*" >> $O; python bench.py "$@" 2>&1 | tail -1 >> $O; }
run --steps 30 --warmup 3
run() { echo "== $*" >> $O; python bench.py --no-tick "$@" 2>&1 | tail -1 >> $O; }
run --steps 30 --warmup 3 --fixed
run --steps 30 --warmup 3 --fixed --streams 8192
run --steps 30 --warmup 3 --fixed --streams 8192 --fs 8000
run --steps 30 --warmup 3 --fs 8000
run --steps 20 --warmup 3 --fs 32000 --streams 2048 --frames-per-step 50
run --steps 20 --warmup 3 --fs 48000 --streams 2048 --frames-per-step 50
run --steps 20 --warmup 3 --fs 48000 --streams 2048 --frames-per-step 50 --fixed
run --steps 200 --warmup 10 --frames-per-step 1 --no-cpu --no-e2e
run --steps 200 --warmup 10 --frames-per-step 1 --no-cpu --no-e2e --streams 32768
run --steps 100 --warmup 5 --frames-per-step 10 --no-cpu --no-e2e
run --config5 --total-streams 8192 --seconds 600
python - "$O" <<'PY'
import json,sys
lines=open(sys.argv[1]).read().splitlines()
for i in range(0,len(lines)-1,2):
    if not lines[i].startswith("=="): continue
    try: d=json.loads(lines[i+1])
    except Exception: print(lines[i],"FAILED"); continue
    cb=d.get("cpu_baseline") or {}
    e=d.get("e2e") or {}
    print("%-72s %8.4f ms/step  value %10.0f  e2e %10.0f (blocking %10.0f)  cpu %8.0f (%s cores)  frac %.4f" % (
        lines[i][3:], d["ms_per_step"], d["value"], e.get("value") or 0, e.get("blocking_value") or 0,
        cb.get("value") or 0, cb.get("cores"), d["roofline"]["frac"]))
PY
