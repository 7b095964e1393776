#!/bin/bash
# Compare differently compiled builds of the sweep kernel on one box (outputs under gpurun_out/).
set -u
mkdir -p gpurun_out
OUT=gpurun_out/variants.log
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $OUT 2>&1
( cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/consumer consumer.cu && /tmp/consumer ) >> $OUT 2>&1
for V in "$@"; do
  lib=multiregionfoam_b200/lib/variants/${V%%:*}.so
  envs=""
  case "$V" in *:*) envs="${V#*:}";; esac
  echo "=== $V" >> $OUT
  env B200_LDU_LIB=$lib $envs timeout 300 python scripts/sweep_variants.py "$V" >> $OUT 2>&1
done
cat $OUT
