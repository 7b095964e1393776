#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + one full capture of the dominant kernels.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --iters 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
# skip the warm-up solves: capture in the steady part (sweeps of the 3rd solve onwards)
ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_amul|k_bicg_xr' -s 60 -c 6 -o gpurun_out/top_kernels \
    python bench.py --steps 1 --warmup 3 --iters 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
