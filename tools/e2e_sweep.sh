#!/bin/bash
# Tuning helper (GPU box): end-to-end streaming / blocking throughput over staging-buffer counts and chunk sizes.
for bufs in ${BUFS:-3}; do for ch in ${CHUNKS:-0 6 12 20}; do
  if [ $ch = 0 ]; then unset NSB200_CHUNK_FRAMES; else export NSB200_CHUNK_FRAMES=$ch; fi
  NSB200_STAGE_BUFS=$bufs python bench.py --steps 30 --warmup 3 --no-cpu "$@" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']
print('bufs $bufs chunk $ch: streaming %.4g (%.2f ms/step)  blocking %.4g (%.2f ms/step)' % (e['value'], 40960/e['value']*1e3*d['config']['streams_per_gpu']/4096, e['blocking_value'], 40960/e['blocking_value']*1e3*d['config']['streams_per_gpu']/4096))"
done; done
