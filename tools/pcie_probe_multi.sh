#!/bin/bash
# GPU box: what each GPU's host<->device copies reach when 1, 2, 4, 8 GPUs copy at once (one process per GPU).
# usage: tools/pcie_probe_multi.sh "<N list>"   -> stdout table (tee it into gpurun_out/)
for n in $1; do
  echo "== $n GPU(s) copying at once"
  for ((i = 0; i < n; ++i)); do CUDA_VISIBLE_DEVICES=$i tools/bin/pcie_probe dma 2 > /tmp/pp_$i.log 2>&1 & done
  wait
  for k in "duplex 1-D" "duplex 2-D" "H2D 1-D" "D2H 1-D"; do
    vals=$(for ((i = 0; i < n; ++i)); do grep "$k" /tmp/pp_$i.log | awk '{print $(NF-1)}'; done | tr '\n' ' ')
    sum=$(echo $vals | tr ' ' '\n' | awk '{s+=$1} END {printf "%.1f", s}')
    echo "  $k: per GPU [$vals] total $sum GB/s"
  done
done
