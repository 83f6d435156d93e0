#!/bin/bash
# Tuning helper: alternative build of the library with extra -D flags, for tools/bench_variants.sh.
# usage: tools/build_variant.sh <name> [-DNSF_CTAS_PER_SM=10 ...]   -> audiosignalprocess_b200/variants/lib<name>.so
name=$1; shift
d=$(dirname "$0")/../audiosignalprocess_b200
mkdir -p $d/variants
nvcc -std=c++17 -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo -shared \
  -Xcompiler -fPIC -Xcompiler -Wno-enum-compare "$@" -Xptxas -v -o $d/variants/lib$name.so $d/csrc/ns_capi.cu 2>&1 \
  | grep -A2 "nsf_process_kernelILi256ELi1ELb1ELb0" | grep -E "registers|spill" | tr '\n' ' '
echo " <- $name"
