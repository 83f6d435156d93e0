#!/bin/bash
# GPU box with N GPUs: the scaling sweep the driver runs at round end (both arms), kept for profiles/.
# usage: tools/scale_run.sh <tag> "<N list>"      -> gpurun_out/<tag>_scale.log (one JSON line per run)
tag=${1:-rX}; O=gpurun_out; mkdir -p $O; : > $O/${tag}_scale.log
for n in $2; do
  if [ "$n" = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517"; fi
  $L bench.py --impl reference --gpus $n --steps 20 --warmup 3 2>/dev/null | grep '^{' | tee -a $O/${tag}_scale.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('N=$n reference: %.0f audio-s/s (%d threads)' % (d['value'], d['cpu_baseline']['cores']))"
  $L bench.py --gpus $n --steps 30 --warmup 3 --no-cpu 2>/dev/null | grep '^{' | tee -a $O/${tag}_scale.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); e=d['e2e']
    print('N=$n ours: value %.4g  e2e %.4g (blocking %.4g)  PCIe %.1f of %.1f GB/s per direction (probe), frac %.2f' % (d['value'], e['value'], e['blocking_value'], e['pcie_gbs'], e['pcie_probe_gbs'], e['frac_of_pcie_probe']))"
done
