#!/bin/bash
# Bench lines of the BASELINE configurations that are not the default bench (profiles/): the 2-D-faithful variants of
# C2 / C3 (SURVEY 8d asks for both), the block-coupled C5, and (when present) the partitioned C4.
set -u
mkdir -p gpurun_out
for W in "$@"; do
  echo "=== $W" >&2
  timeout ${CFG_TIMEOUT:-600} python bench.py --workload $W --steps ${CFG_STEPS:-3} --warmup 3 > gpurun_out/r02_bench_${W}.json 2> gpurun_out/r02_bench_${W}.err
  echo "exit $?" >> gpurun_out/r02_bench_${W}.err
  tail -c 400 gpurun_out/r02_bench_${W}.json >&2
done
