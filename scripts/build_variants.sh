#!/bin/bash
# Build differently configured copies of the library for one-box comparisons (scripts/gpu_variants.sh).
# usage: scripts/build_variants.sh name:"-DFLAG=.. -DFLAG2=.." ...
set -eu
cd "$(dirname "$0")/.."
mkdir -p multiregionfoam_b200/lib/variants
pids=()
for V in "$@"; do
  name=${V%%:*}; flags=""
  case "$V" in *:*) flags="${V#*:}";; esac
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-O3,-Wall -shared $flags \
      -Xptxas=-v -o multiregionfoam_b200/lib/variants/$name.so multiregionfoam_b200/csrc/b200_ldu.cu -ldl \
      > multiregionfoam_b200/lib/variants/$name.log 2>&1 && echo "built $name" || echo "FAILED $name" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
