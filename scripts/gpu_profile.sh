#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + one full capture of the dominant kernels.
#   scripts/gpu_profile.sh [tag] [extra bench.py arguments]      (default workload of bench.py: C3, 64 M cells)
set -u
TAG=${1:-r02}
shift || true
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --iters 2 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
# skip the warm-up solves: capture in the steady part (sweeps of the 3rd solve onwards)
ncu --set full --clock-control none --import-source on -k regex:'k_sweep|k_amul|k_bicg_xr' -s 60 -c 6 -o gpurun_out/${TAG}_top_kernels \
    python bench.py --steps 1 --warmup 3 --iters 2 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${TAG}_bench_under_ncu2.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md 2>&1
python scripts/ncu_summary.py full gpurun_out/${TAG}_top_kernels.ncu-rep > gpurun_out/${TAG}_full.md 2>&1
ls -la gpurun_out | tail -12
