#!/usr/bin/env python
"""bench.py -- coupled-solve throughput of the B200 LDU solver on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2] [--iters 50]

A *step* is one monolithic coupled solve (BiCGStab + DILU, fixed ``--iters`` Krylov iterations:
minIter = maxIter, tolerance 0) of the synthetic two-region CHT system, from the same initial
guess every step.  metric = cell-iterations/s = cells x iterations executed / time.

* ``value``: device-resident (matrix, x0, b already in HBM), timed with CUDA events inside the library
  on its stream (b200_perf.deviceMs summed over the K steps; max over ranks).
* ``e2e``: the reference-facing call ``b200_sys_set_coeffs`` + ``b200_solve`` with HOST buffers
  (page-locked): coefficients, x and b go host->device and x comes back inside the timed region.
* ``roofline``: the dominant kernel class (by device time, measured live with per-launch CUDA events
  in a separate profiled pass over the same steps) against MEASURED_PEAKS.json.
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle port (oracle/ldu_oracle.c, serial, same
  algorithm) on a bounded sample of the same workload; the reference's own solver lives in
  foam-extend 4.1, which is not in the reference tree.

N > 1 (torchrun): every rank owns one z-slab of the same size (weak scaling, foam-extend processor
decomposition ``simple (1 1 N)``), halo exchange + all-reduce over NCCL inside the library.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "coupled_solve_cell_iterations_per_s"
UNIT = "cell-iterations/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.proc, self.path = None, f"/tmp/b200_clocks_{os.getpid()}.csv"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# workloads whose TOTAL size is fixed: N ranks share the case (strong scaling); every other workload gives each rank a
# slab of the listed size (weak scaling)
STRONG_WORKLOADS = {"C3"}


def build_rank_system(workload: str, rank: int, nranks: int):
    from multiregionfoam_b200.assembly import WORKLOADS, cht_rank_slab
    r, L = WORKLOADS[workload]
    if workload in STRONG_WORKLOADS:
        if L % nranks:
            raise SystemExit(f"{workload} has {L} z-layers: not divisible by {nranks} ranks")
        L //= nranks
    return cht_rank_slab(r, L, rank, nranks), (r, L)


def oracle_sample(workload: str, iters: int, nsub: int = 1, threads: int = 1):
    """The CPU port on the N=1 workload for `iters` BiCGStab+DILU iterations.  nsub > 1: the case is decomposed into
    nsub z-slab sub-domains (foam-extend's processor decomposition, block-Jacobi preconditioner) worked on by
    `threads` threads standing in for the MPI ranks of a decomposed foam-extend run on the same host."""
    from multiregionfoam_b200.assembly import WORKLOADS, cht_rank_slab
    from multiregionfoam_b200.case import Case
    from oracle import pyoracle
    r, L = WORKLOADS[workload]
    if nsub > 1:
        assert L % nsub == 0
        case = Case(workload, [cht_rank_slab(r, L // nsub, g, nsub) for g in range(nsub)])
    else:
        case = Case(workload, [cht_rank_slab(r, L, 0, 1)])
    O = pyoracle.OracleSystem(case)
    pyoracle.set_threads(threads)
    x0, b = case.concat("psi"), case.concat("source")
    t = time.perf_counter()
    _, info = O.solve(x0, b, "BiCGStab", "DILU", tolerance=0.0, minIter=iters, maxIter=iters)
    dt = time.perf_counter() - t
    pyoracle.set_threads(1)
    return case.nCells * info["nIterations"] / dt, dt, case.nCells, case.nFaces, info["nIterations"]


# decomposition of the CPU arm: a property of the WORKLOAD, not of the box (the iteration count a decomposed
# block-Jacobi solve needs, and its arithmetic, depend on it); the sub-domains are work items of a thread pool, so the
# thread count only changes the speed
REF_SUBDOMAINS = {"C2": 11, "C2-2D": 1, "C3-slab8": 23, "C3": 23, "C3-2D": 1, "C1": 1}


def host_decomposition(workload: str):
    """(sub-domains, threads) of the decomposed CPU arm: a fixed z-slab decomposition per workload (a divisor of its
    layer count), worked on by min(sub-domains, host cores) threads."""
    from multiregionfoam_b200.assembly import WORKLOADS
    _, L = WORKLOADS[workload]
    nsub = REF_SUBDOMAINS.get(workload, 1)
    assert L % nsub == 0
    return nsub, max(1, min(nsub, os.cpu_count() or 1))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, times = [], []
    # the same step as the GPU arm: args.iters Krylov iterations (--ref-iters only to shorten a manual run); the CPU needs
    # no five warm-up solves, and the sample is bounded by running at most ref_max_steps of them
    it = max(1, args.iters if args.ref_iters is None else min(args.iters, args.ref_iters))
    nCells = nFaces = 0
    nsub, threads = host_decomposition(args.workload)
    args.warmup = min(args.warmup, 1)
    args.steps = min(args.steps, args.ref_max_steps)
    t_begin = time.perf_counter()
    for s in range(args.warmup + args.steps):
        v, dt, nCells, nFaces, _ = oracle_sample(args.workload, it, nsub, threads)
        if s >= args.warmup:
            vals.append(v)
            times.append(dt)
        if times and time.perf_counter() - t_begin > args.ref_max_seconds:
            break   # bounded sample: a step of the 64 M-cell case is ~40 s of CPU
    total = nCells * it * len(times) / sum(times)
    sample = (f"{len(times)} steps of {it} BiCGStab+DILU iterations on {'one GPU-share (z-slab) of ' if args.gpus > 1 else 'the full '}"
              f"{args.workload} ({nCells} cells) decomposed into {nsub} z-slab sub-domains (block-Jacobi DILU, as foam-extend's MPI run), "
              f"{threads} threads of the CPU oracle port")
    out = {
        "impl": "reference", "metric": METRIC, "value": total, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong" if args.workload in STRONG_WORKLOADS else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "cells": nCells, "faces": nFaces, "solver": "BiCGStab", "preconditioner": "DILU",
                   "iterations_per_step": it, "decomposition": f"simple (1 1 {nsub})", "cells_per_gpu": nCells},
        "cpu_baseline": {"value": total, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": total, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU oracle port of foam-extend's coupled BiCGStab/DILU (foam-extend 4.1 is not in the reference tree; oracle/_ref unbuildable)",
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------ block-coupled workload (C5)
# BASELINE configs[4]: "block-coupled vector equation" at 16 M cells: the fvBlockMatrix<vector4> p-U system of
# pUCoupledIcoFluid (src/regions/pUCoupledIcoFluid/pUCoupledIcoFluid.C:584-621) on a structured box, solved by
# BlockBiCGStab + BlockCholesky through include/b200_blk.h (fvBlockMatrix<vector4>::solve, filesToReplace/fvBlockMatrix.C:1360-1388).
BLOCK_WORKLOADS = {"C5": (256, 256, 244), "C5-2M": (160, 128, 100)}  # box nx ny nz: 15 990 784 / 2 048 000 cells
BLOCK_ITERS = 30


def block_system(workload: str):
    from multiregionfoam_b200.assembly import pu_block_matrix
    from multiregionfoam_b200.mesh import Block, StructuredRegion
    nx, ny, nz = BLOCK_WORKLOADS[workload]
    m = StructuredRegion("box", [Block(nx, 0.0, 1.0, 1.0)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
    M = pu_block_matrix(m.nCells, m.lowerAddr, m.upperAddr)
    return M, (nx, ny, nz)


def block_oracle_sample(M, iters: int):
    from oracle import pyblk
    O = pyblk.BlockOracle(M.l, M.u, M.nCells, M.diag, M.upper, M.lower)
    t = time.perf_counter()
    O.solve(M.psi, M.source, "BiCGStab", "Cholesky", tolerance=0.0, minIter=iters, maxIter=iters)
    dt = time.perf_counter() - t
    O.close()
    return M.nCells * iters / dt, dt


def run_block_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    M, box = block_system(args.workload)
    # bounded sample: the scalar port manages ~1.4 M block cell-iterations/s per core, so a step is `it` iterations with
    # cells x it / 1.4e6 of the order of 20 s
    it = max(1, min(BLOCK_ITERS, int(20 * 1.4e6 / M.nCells) or 1))
    steps = max(1, min(args.steps, 3))
    times = []
    for _ in range(steps):
        _, dt = block_oracle_sample(M, it)
        times.append(dt)
    total = M.nCells * it * len(times) / sum(times)
    sample = (f"{len(times)} steps of {it} BlockBiCGStab+BlockCholesky iterations on the full {args.workload} system ({M.nCells} cells), "
              f"1 thread of the CPU block oracle port (oracle/blk_oracle.c)")
    out = {"impl": "reference", "metric": METRIC, "value": total, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 0,
           "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": args.workload, "case": f"p-U block system (vector4, SQUARE coefficients), box {box[0]}x{box[1]}x{box[2]}",
                      "cells": M.nCells, "faces": int(M.l.size), "solver": "BlockBiCGStab", "preconditioner": "BlockCholesky",
                      "iterations_per_step": it},
           "cpu_baseline": {"value": total, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "host_cores_available": os.cpu_count()},
           "e2e": {"value": total, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
           "note": "reference = CPU oracle port of foam-extend's BlockBiCGStab / BlockCholesky (foam-extend 4.1 is not in the reference tree)"}
    print(json.dumps(out))


def run_block(args):
    """bench.py --workload C5: the same contract line for the block-coupled path (one GPU: the block library has no
    processor patches yet)."""
    import torch  # noqa: F401  (device bring-up as in the scalar arm)
    from multiregionfoam_b200 import blockldu, ldu
    if args.gpus != 1:
        raise SystemExit("the block-coupled workload runs on one GPU")
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    t0 = time.perf_counter()
    M, box = block_system(args.workload)
    n, F = M.nCells, int(M.l.size)
    assemble_s = time.perf_counter() - t0
    ctx = ldu.Context(0)
    t0 = time.perf_counter()
    S = blockldu.BlockSystem(ctx, M.l, M.u, n)
    S.set_coeffs(M.diag, M.upper, M.lower)
    finalize_s = time.perf_counter() - t0
    S.upload(M.psi, M.source)
    S.x_save()
    iters = BLOCK_ITERS
    opts = dict(solver=blockldu.SOLVER_BICGSTAB, precond=ldu.PRECOND_CHOLESKY, tolerance=0.0, minIter=iters, maxIter=iters)

    def step():
        S.x_restore()
        return S.solve_resident(**opts)

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(0)
    launches0 = ctx.launches
    dev_ms, its = 0.0, 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        p = step()
        dev_ms += p["deviceMs"]
        its += p["nIterations"]
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    launches = ctx.launches - launches0
    clk = clocks.stop()
    value = n * its / (dev_ms * 1e-3)
    # per kernel class
    S.set_profiling(True)
    S.kernel_times(reset=True)
    for _ in range(min(args.steps, 2)):
        step()
    kt = S.kernel_times(reset=True)
    S.set_profiling(False)
    peak, peak_src = measured_peak_gbs()
    dk = 16 if np.asarray(M.diag).ndim == 3 else (4 if np.asarray(M.diag).ndim == 2 else 1)
    uk = 16 if np.asarray(M.upper).ndim == 3 else (4 if np.asarray(M.upper).ndim == 2 else 1)
    # algorithmic bytes per launch (block rows): Amul reads diag, x, writes y, reads upper + lower + addressing;
    # a sweep of BlockCholesky reads the inverted diagonal and one coefficient array, reads and writes the vector
    alg = {"amul": (8 * dk + 64) * n + (16 * uk + 8) * F, "sweep_fwd": (8 * dk + 64) * n + (8 * uk + 8) * F,
           "sweep_bwd": (8 * dk + 64) * n + (8 * uk + 8) * F}
    kernels = {}
    for k, (ms, cnt) in kt.items():
        if cnt == 0:
            continue
        kernels[k] = {"ms_total": ms, "launches": cnt, "ms_per_launch": ms / cnt}
        if k in alg and ms > 0:
            kernels[k]["gbs"] = alg[k] * cnt / (ms * 1e-3) / 1e9
            kernels[k]["frac"] = kernels[k]["gbs"] / peak
    tot = sum(v["ms_total"] for v in kernels.values())
    for v in kernels.values():
        v["share"] = v["ms_total"] / tot if tot else 0.0
    dom = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["ms_total"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom]}
    # per BlockBiCGStab iteration: 2 Amul + 2 x (fwd + bwd) + 10 vector passes of 32 N
    bytes_it = 2 * alg["amul"] + 4 * alg["sweep_fwd"] + 320 * n
    solve_gbs = bytes_it * its / (dev_ms * 1e-3) / 1e9
    # ---- end to end: coefficients, x and b from (page-locked) host arrays, x back
    e2e = None
    if not args.no_e2e:
        keep = [np.ascontiguousarray(a) for a in (M.diag, M.upper, M.lower, M.source)]
        xs = [np.ascontiguousarray(M.psi.copy()) for _ in range(4)]
        for a in keep + xs:
            ctx.host_register(a)
        n_e = 2

        def e2e_step(k):
            S.set_coeffs(keep[0], keep[1], keep[2])
            x, pp = S.solve(xs[k], keep[3], history=False, **opts)
            return pp["nIterations"]

        e2e_step(0)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        its_e = sum(e2e_step(1 + k) for k in range(n_e))
        e_s = time.perf_counter() - t1
        h2d = sum(a.nbytes for a in keep) + xs[0].nbytes
        e2e = {"value": n * its_e / e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(xs[0].nbytes),
               "ms_per_step": 1e3 * e_s / n_e, "steps": n_e}
        for a in keep + xs:
            ctx.host_unregister(a)
    cpu = None
    if not args.no_cpu_baseline:
        it = max(1, min(iters, int(20 * 1.4e6 / n) or 1))
        v, dt = block_oracle_sample(M, it)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{it} BlockBiCGStab+BlockCholesky iterations on the full {args.workload} system ({n} cells), 1 thread of the CPU block "
                         f"oracle port, {dt:.1f} s", "host_cores_available": os.cpu_count()}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.workload, "case": f"p-U block system (vector4, SQUARE coefficients), box {box[0]}x{box[1]}x{box[2]}",
                      "cells": n, "faces": F, "cells_per_gpu": n, "solver": "BlockBiCGStab", "preconditioner": "BlockCholesky",
                      "iterations_per_step": iters, "l2": "inputs larger than L2 (matrix + vectors >> 126 MB)"},
           "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
           "solve_hbm_gbs_per_gpu": solve_gbs, "solve_roofline_frac": solve_gbs / peak, "algorithmic_bytes_per_iteration_per_gpu": bytes_it,
           "kernels": kernels, "wall_ms_per_step": wall_ms / args.steps, "assemble_s": assemble_s, "finalize_s": finalize_s,
           "final_residual": [float(v) for v in np.atleast_1d(p["finalResidual"])]}
    real_stdout.write(json.dumps(out) + "\n")
    real_stdout.flush()
    S.close()
    ctx.close()


# ------------------------------------------------------------------------------------------ partitioned FSI workload (C4)
def fsi_oracle_ops(case):
    """The callbacks of fsi.coupling_iteration on the CPU oracle (cpu_baseline / reference arm only)."""
    from multiregionfoam_b200.case import Case, RankSystem
    from oracle import pyoracle
    O = {k: pyoracle.OracleSystem(Case(k, [RankSystem(0, 1, [reg])])) for k, reg in (("U", case.fluidU), ("p", case.fluidP), ("D", case.solidD))}

    def solve(key, x0, b, solver, precond, iters):
        return O[key].solve(x0, b, solver, precond, tolerance=0.0, minIter=iters, maxIter=iters)[0]

    def transfer(tab, f):
        return pyoracle.ggi_interpolate(tab[0], tab[1], tab[2], f, 3)

    return solve, transfer


def fsi_config(args, case):
    from multiregionfoam_b200 import fsi
    return {"workload": args.workload,
            "case": f"HronTurekFsi3 topology (fluid 24 blocks + solid (105 6 1)) r={case.r}, {case.layers} z-layers, partitioned coupling over a GGI interface",
            "cells": case.nFluid + case.nSolid, "cells_fluid": case.nFluid, "cells_solid": case.nSolid,
            "faces": case.fluidU.nFaces + case.solidD.nFaces, "interface_faces": [int(case.fluidFaceCells.size), int(case.solidFaceCells.size)],
            "solver": f"per coupling iteration: U 3 x PBiCG+DILU ({fsi.N_U} its), p PCG+DIC ({fsi.N_P}), D 3 x PCG+DIC ({fsi.N_D}), 2 GGI transfers",
            "cell_iterations_per_step": case.cell_iterations(), "l2": "inputs larger than L2"}


def run_fsi_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from multiregionfoam_b200 import fsi
    from oracle import pyoracle
    case = fsi.fsi_case(*fsi.FSI_WORKLOADS[args.workload])
    solve, transfer = fsi_oracle_ops(case)
    pyoracle.set_threads(min(8, os.cpu_count() or 1))
    state = fsi.initial_state(case)
    steps = max(1, min(args.steps, 3))
    t = time.perf_counter()
    for _ in range(steps):
        state = fsi.coupling_iteration(case, solve, transfer, state)
    dt = time.perf_counter() - t
    pyoracle.set_threads(1)
    total = case.cell_iterations() * steps / dt
    threads = min(8, os.cpu_count() or 1)
    sample = f"{steps} coupling iterations of the full {args.workload} case on the CPU oracle port ({threads} threads for the vector updates and Amul rows, sweeps serial)"
    out = {"impl": "reference", "metric": METRIC, "value": total, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 0,
           "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": fsi_config(args, case),
           "cpu_baseline": {"value": total, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "host_cores_available": os.cpu_count()},
           "e2e": {"value": total, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
           "note": "reference = CPU oracle port (foam-extend 4.1 is not in the reference tree)"}
    print(json.dumps(out))


def run_fsi(args):
    """bench.py --workload C4: one step = one Dirichlet-Neumann coupling iteration of the partitioned FSI case through the
    C ABI (every solve: x0 and b host -> device, x back; every interface transfer: b200_ggi_interpolate).  `value` counts the
    device time of the solves (b200_perf.deviceMs), `e2e` the wall clock of the whole iteration including the transfers."""
    import torch  # noqa: F401
    from multiregionfoam_b200 import fsi, ldu
    if args.gpus != 1:
        raise SystemExit("the partitioned FSI workload runs on one GPU")
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    case = fsi.fsi_case(*fsi.FSI_WORKLOADS[args.workload])
    ctx = ldu.Context(0)
    t0 = time.perf_counter()
    D = fsi.DeviceFsi(ctx, case)
    finalize_s = time.perf_counter() - t0
    state = fsi.initial_state(case)
    for _ in range(args.warmup):
        state = fsi.coupling_iteration(case, D.solve, D.transfer, state)
    D.device_ms, D.iterations = 0.0, 0
    clocks = ClockSampler(0)
    launches0 = ctx.launches
    t1 = time.perf_counter()
    for _ in range(args.steps):
        state = fsi.coupling_iteration(case, D.solve, D.transfer, state)
    wall_s = time.perf_counter() - t1
    launches = ctx.launches - launches0
    clk = clocks.stop()
    work = case.cell_iterations() * args.steps
    value = work / (D.device_ms * 1e-3)
    # per kernel class over the three systems
    for S in D.sys.values():
        S.set_profiling(True)
        S.kernel_times(reset=True)
    fsi.coupling_iteration(case, D.solve, D.transfer, state)
    peak, peak_src = measured_peak_gbs()
    kernels, alg_tot = {}, {}
    for key, S in D.sys.items():
        N, F = S.nCells, S.nFaces
        sym = key != "U"
        alg = {"amul": 24 * N + (16 if sym else 24) * F, "sweep_fwd": 24 * N + 16 * F, "sweep_bwd": 24 * N + 16 * F}
        for k, (ms, cnt) in S.kernel_times(reset=True).items():
            if cnt == 0:
                continue
            e = kernels.setdefault(k, {"ms_total": 0.0, "launches": 0})
            e["ms_total"] += ms
            e["launches"] += cnt
            if k in alg:
                alg_tot[k] = alg_tot.get(k, 0) + alg[k] * cnt
        S.set_profiling(False)
    tot = sum(v["ms_total"] for v in kernels.values())
    for k, v in kernels.items():
        v["ms_per_launch"] = v["ms_total"] / v["launches"]
        v["share"] = v["ms_total"] / tot if tot else 0.0
        if k in alg_tot and v["ms_total"] > 0:
            v["gbs"] = alg_tot[k] / (v["ms_total"] * 1e-3) / 1e9
            v["frac"] = v["gbs"] / peak
    dom = max((k for k in kernels if k in alg_tot), key=lambda k: kernels[k]["ms_total"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_tot[dom] / kernels[dom]["launches"]}
    xbytes = 8 * (case.nFluid * 4 + case.nSolid * 3)
    e2e = {"value": work / wall_s, "unit": UNIT, "h2d_bytes_per_step": int(2 * xbytes), "d2h_bytes_per_step": int(xbytes),
           "ms_per_step": 1e3 * wall_s / args.steps, "steps": args.steps,
           "note": "whole coupling iteration through the C ABI: x0 and b of the 7 solves in, x out, interface fields through b200_ggi_interpolate"}
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import pyoracle
        solve, transfer = fsi_oracle_ops(case)
        threads = min(8, os.cpu_count() or 1)
        pyoracle.set_threads(threads)
        t2 = time.perf_counter()
        fsi.coupling_iteration(case, solve, transfer, fsi.initial_state(case))
        dt = time.perf_counter() - t2
        pyoracle.set_threads(1)
        cpu = {"value": case.cell_iterations() / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"one coupling iteration of the full {args.workload} case on the CPU oracle port, {dt:.1f} s", "host_cores_available": os.cpu_count()}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": D.device_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": fsi_config(args, case), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
           "kernels": kernels, "finalize_s": finalize_s}
    real_stdout.write(json.dumps(out) + "\n")
    real_stdout.flush()
    D.close()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="default: C2 on one GPU (BASELINE configs[1]); C3-slab8 - one eighth of the "
                    "64 M-cell C3 per GPU, configs[2] - on several")
    ap.add_argument("--iters", type=int, default=50, help="Krylov iterations per step (minIter = maxIter)")
    ap.add_argument("--ref-iters", type=int, default=None, help="CPU arm: fewer iterations per step than the GPU arm (manual runs only)")
    ap.add_argument("--ref-max-steps", type=int, default=12, help="CPU arm: at most this many timed steps (bounded sample)")
    ap.add_argument("--ref-max-seconds", type=float, default=150.0, help="CPU arm: no further step is started after this many seconds")
    ap.add_argument("--cpu-baseline-iters", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload is None:
        # the configuration north_star quotes its targets on: the 64 M-cell two-region CHT case (BASELINE configs[2]).  It fits
        # one B200 (~10 GB), so N = 1 runs the whole case and N = 2 / 4 / 8 split the SAME case into z-slabs (strong scaling);
        # other rank counts fall back to one C3 slab per rank.  configs[1] (C2, 4.3 M cells) is `--workload C2`.
        args.workload = "C3" if args.gpus in (1, 2, 4, 8) else "C3-slab8"

    from multiregionfoam_b200.fsi import FSI_WORKLOADS
    if args.workload in FSI_WORKLOADS:
        return run_fsi_reference(args) if args.impl == "reference" else run_fsi(args)
    if args.workload in BLOCK_WORKLOADS:
        return run_block_reference(args) if args.impl == "reference" else run_block(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from multiregionfoam_b200 import ldu
    from multiregionfoam_b200.case import algorithmic_bytes_per_iteration

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
            os.execv(sys.executable, cmd)
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    # Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner there) write to fd 1 too
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.frombuffer(bytearray(ldu.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = ldu.Context(device=local, rank=rank, nranks=world, unique_id=uid)

    # ---- N > 1: before anything is timed, the halo exchange and the global sums of THIS run's transport are checked
    # against the CPU oracle on the same decomposition (small case: r = 1, 3 layers per rank): Amul bit-exact, 20
    # iterations of the residual history within 1e-10, the field within 1e-8.  A run that fails this prints no number.
    parity = None
    if world > 1:
        from multiregionfoam_b200.assembly import cht_rank_slab
        from multiregionfoam_b200.case import Case
        prs = cht_rank_slab(1, 3, rank, world)
        PS = ldu.LduSystem(ctx, prs)
        px0 = np.concatenate([g.psi for g in prs.regions])
        pb = np.concatenate([g.source for g in prs.regions])
        pxr = np.random.default_rng(100 + rank).standard_normal(px0.size)
        py = PS.amul(pxr)
        pxs, pinfo = PS.solve(px0, pb, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=0.0, minIter=20, maxIter=20)
        PS.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, dict(xr=pxr, y=py, xs=pxs, hist=pinfo["history"]))
        if rank == 0:
            from oracle import pyoracle
            pcase = Case("parity", [cht_rank_slab(1, 3, g, world) for g in range(world)])
            PO = pyoracle.OracleSystem(pcase)
            yo = PO.amul(np.concatenate([g["xr"] for g in gathered]))
            xo, io = PO.solve(pcase.concat("psi"), pcase.concat("source"), "BiCGStab", "DILU", tolerance=0.0, minIter=20, maxIter=20)
            hg, ho = gathered[0]["hist"][:21], io["history"][:21]
            parity = {"ranks": world, "cells": int(pcase.nCells),
                      "amul_bit_exact": bool(np.array_equal(np.concatenate([g["y"] for g in gathered]), yo)),
                      "history_max_rel_err_20": float(np.max(np.abs(hg - ho) / np.abs(ho))),
                      "field_rel_l2": float(np.linalg.norm(np.concatenate([g["xs"] for g in gathered]) - xo) / np.linalg.norm(xo)),
                      "history_identical_on_all_ranks": bool(all(np.array_equal(g["hist"], gathered[0]["hist"]) for g in gathered)),
                      "transport": os.environ.get("B200_TRANSPORT", "auto (peer-to-peer, NCCL fall-back)")}
            okp = (parity["amul_bit_exact"] and parity["history_max_rel_err_20"] < 1e-10 and parity["field_rel_l2"] < 1e-8
                   and parity["history_identical_on_all_ranks"])
            print(f"[bench] multi-rank parity: {parity}", file=sys.stderr, flush=True)
        flag = torch.tensor([1 if (rank != 0 or okp) else 0], device="cuda")
        dist.broadcast(flag, 0)
        if int(flag.item()) == 0:
            raise SystemExit("multi-rank parity against the oracle FAILED: no bench line")

    rs, (r, L) = build_rank_system(args.workload, rank, world)
    t0 = time.perf_counter()
    S = ldu.LduSystem(ctx, rs)
    finalize_s = time.perf_counter() - t0
    nLocal, fLocal = S.nCells, S.nFaces
    nGlobal, fGlobal = nLocal * world, fLocal * world
    x0 = np.concatenate([reg.psi for reg in rs.regions])
    b = np.concatenate([reg.source for reg in rs.regions])
    S.upload(x0, b)
    S.x_save()
    kw = dict(solver=ldu.SOLVER_BICGSTAB, precond=ldu.PRECOND_DILU, tolerance=0.0, relTol=0.0, minIter=args.iters, maxIter=args.iters)

    def resident_step():
        S.x_restore()
        return S.solve_resident(**kw)

    # ---- device-resident timing
    for _ in range(args.warmup):
        resident_step()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launches
    dev_ms, its = 0.0, 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        info = resident_step()
        dev_ms += info["deviceMs"]
        its += info["nIterations"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    launches = ctx.launches - launches0
    clk = clocks.stop() if clocks else None
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    value = nGlobal * its / (dev_ms * 1e-3)
    final_res = info["finalResidual"]

    # ---- profiled pass: per-launch CUDA events per kernel class (separate from the timed run)
    S.set_profiling(True)
    S.kernel_times(reset=True)
    prof_steps = min(args.steps, 2)
    for _ in range(prof_steps):
        resident_step()
    kt = S.kernel_times(reset=True)
    S.set_profiling(False)
    peak, peak_src = measured_peak_gbs()
    N, F = nLocal, fLocal
    alg_bytes = {  # per launch, SURVEY 8(d) / DESIGN.md section 5
        "amul": 24 * N + 24 * F, "sweep_fwd": 24 * N + 16 * F, "sweep_bwd": 24 * N + 16 * F,
    }
    kernels = {}
    for k, (ms, n) in kt.items():
        if n == 0:
            continue
        kernels[k] = {"ms_total": ms, "launches": n, "ms_per_launch": ms / n}
        if k in alg_bytes and ms > 0:
            kernels[k]["gbs"] = alg_bytes[k] * n / (ms * 1e-3) / 1e9
            kernels[k]["frac"] = kernels[k]["gbs"] / peak
    tot_ms = sum(v["ms_total"] for v in kernels.values())
    for v in kernels.values():
        v["share"] = v["ms_total"] / tot_ms if tot_ms else 0.0
    dom = max((k for k in kernels if k in alg_bytes), key=lambda k: kernels[k]["ms_total"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes[dom]}
    bytes_it = algorithmic_bytes_per_iteration(N, F, "BiCGStab")
    solve_gbs = bytes_it * its / (dev_ms * 1e-3) / 1e9  # per GPU (N, F local; dev_ms = max over ranks)

    # ---- end to end through the reference-facing call with host buffers
    e2e = None
    if not args.no_e2e:
        n_e = max(2, min(args.steps, 5))
        # one set of x buffers (initial guess in, solution out) per step, made before the timed region: a caller's x is
        # its own field, not something the solver resets
        hostx_steps = [[np.ascontiguousarray(reg.psi.copy()) for reg in rs.regions] for _ in range(n_e + 2)]
        keep = []
        for reg in rs.regions:
            for a in (reg.diag, reg.upper, reg.lower, reg.source):
                if a is not None:
                    keep.append(a)
            for itf in reg.interfaces:
                keep += [itf.bouCoeffs, itf.intCoeffs]
        for hx in hostx_steps:
            keep += hx
        for a in keep:
            ctx.host_register(a)
        # a symmetric region's upper coefficients cross the bus once (the library fills the lower half on the device)
        h2d = sum(reg.diag.nbytes + reg.upper.nbytes * (2 if reg.lower is not None else 1) + reg.source.nbytes + reg.psi.nbytes +
                  sum(2 * i.bouCoeffs.nbytes for i in reg.interfaces) for reg in rs.regions)
        d2h = sum(reg.psi.nbytes for reg in rs.regions)
        import ctypes as C
        Lib = ldu.load()
        opts, perf = ldu.SolverOpts(kw["solver"], kw["precond"], 0.0, 0.0, args.iters, args.iters), ldu.Perf()
        bs = ldu._dpp([reg.source for reg in rs.regions])

        def e2e_step(k):
            S.set_all_coeffs()                       # this solve's matrix: host -> device
            xs = ldu._dpp(hostx_steps[k])
            ctx.check(Lib.b200_solve(S.h, C.byref(opts), xs, bs, C.byref(perf), None, 0))  # x, b H2D; solve; x D2H
            return perf.nIterations

        for k in range(2):
            e2e_step(k)
        barrier()
        t1 = time.perf_counter()
        its_e = 0
        for k in range(n_e):
            its_e += e2e_step(2 + k)
        barrier()
        e_s = max_over_ranks(time.perf_counter() - t1)
        e2e = {"value": nGlobal * its_e / e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * e_s / n_e, "steps": n_e}
        for a in keep:
            ctx.host_unregister(a)

        # ---- informational (N = 1): the same step with the T equations assembled on the device (b200_sys_assemble_T,
        # SURVEY 8(f) rank 3): per step only the field crosses the bus (x in, solution out).  Reported beside, never instead
        # of, the host-assembled e2e above (what an unmodified foam-extend caller does); a failure here cannot touch the
        # contract line.
        if world == 1:
            try:
                from multiregionfoam_b200.assembly import cht_fv_tables_slab
                tables, _ = cht_fv_tables_slab(r, L, 0, 1)
                S2 = ldu.LduSystem(ctx, rs, set_coeffs=False)
                try:
                    for ri, t in enumerate(tables):
                        S2.set_fv_geometry(ri, t["V"], t["magSf"], t["deltaCoeffs"], t["bCells"], t["bInt"], t["bSrc"])
                        for i, itf in enumerate(rs.regions[ri].interfaces):
                            S2.set_interface_coeffs(ri, i, itf.bouCoeffs, itf.intCoeffs)
                    hx0 = [np.ascontiguousarray(reg.psi.copy()) for reg in rs.regions]
                    hx = [np.ascontiguousarray(reg.psi.copy()) for reg in rs.regions]
                    for a in hx0 + hx:
                        ctx.host_register(a)
                    first = [True]

                    def dev_step():
                        ctx.check(Lib.b200_upload(S2.h, ldu._dpp(hx0), None))                      # T.oldTime(): H2D
                        for ri, t in enumerate(tables):
                            S2.assemble_T(ri, t["form"], t["rhoC"], t["rDeltaT"], t["kappa"], phi=t["phi"] if first[0] else None)
                        first[0] = False
                        ctx.check(Lib.b200_solve_resident(S2.h, C.byref(opts), C.byref(perf), None, 0))
                        ctx.check(Lib.b200_download(S2.h, ldu._dpp(hx)))                            # solution: D2H
                        return perf.nIterations

                    for _ in range(2):
                        dev_step()
                    barrier()
                    t1 = time.perf_counter()
                    its_d = 0
                    for _ in range(n_e):
                        its_d += dev_step()
                    barrier()
                    d_s = time.perf_counter() - t1
                    xbytes = sum(a.nbytes for a in hx0)
                    e2e["device_assembled"] = {"value": nGlobal * its_d / d_s, "unit": UNIT, "ms_per_step": 1e3 * d_s / n_e,
                                               "h2d_bytes_per_step": int(xbytes), "d2h_bytes_per_step": int(xbytes), "steps": n_e,
                                               "note": "T equations assembled on the device from the uploaded field (b200_sys_assemble_T)"}
                    for a in hx0 + hx:
                        ctx.host_unregister(a)
                finally:
                    S2.close()
            except Exception as ex:  # informational leg only
                e2e["device_assembled"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample (~20 s decomposed + ~15 s serial): iterations scaled to the case with nominal port rates of
        # 70 M / 18 M cell-iterations/s
        it = max(2, min(args.iters, args.cpu_baseline_iters, int(20 * 70e6 / nGlobal)))
        nsub, threads = host_decomposition(args.workload)
        v, dt, nc, nf, _ = oracle_sample(args.workload, it, nsub, threads)
        it1 = max(1, min(it // 3, int(15 * 18e6 / nGlobal)))
        v1, dt1, _, _, _ = oracle_sample(args.workload, it1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{it} BiCGStab+DILU iterations on the full {args.workload} system ({nc} cells) decomposed into {nsub} z-slab "
                         f"sub-domains (block-Jacobi DILU, as foam-extend's MPI run), {threads} threads of the CPU oracle port, {dt:.1f} s",
               "serial": {"value": v1, "cores": 1, "sample": f"{it1} iterations, undecomposed, {dt1:.1f} s"},
               "host_cores_available": os.cpu_count()}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.workload in STRONG_WORKLOADS else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "case": f"two-region CHT (flowOverHeatedPlate topology) r={r}, {L} z-layers per GPU",
                       "cells": nGlobal, "faces": fGlobal, "cells_per_gpu": nLocal, "solver": "BiCGStab", "preconditioner": "DILU",
                       "iterations_per_step": args.iters, "l2": "inputs larger than L2 (matrix + vectors >> 126 MB)",
                       "decomposition": f"simple (1 1 {world})"},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clk, "multirank_parity": parity,
            "solve_hbm_gbs_per_gpu": solve_gbs, "solve_roofline_frac": solve_gbs / peak,
            "algorithmic_bytes_per_iteration_per_gpu": bytes_it,
            "kernels": kernels, "wall_ms_per_step": wall_ms / args.steps, "finalize_s": finalize_s,
            "final_residual": final_res,
        }
        real_stdout.write(json.dumps(out) + "\n")
        real_stdout.flush()
    S.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
