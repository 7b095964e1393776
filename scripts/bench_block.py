"""Measurement of the block-coupled (vector4) path (BASELINE config 5) next to bench.py, which
measures the scalar coupled path: synthetic p-U block system on a structured box, BlockBiCGStab + BlockCholesky, fixed
iteration count, device-resident; per kernel class the algorithmic bytes (block Amul 192 N + 264 F for SQUARE
coefficients with the 10-of-16 pattern stored dense) against the measured HBM peak; the CPU oracle port timed beside it
on a bounded number of iterations.  usage: python scripts/bench_block.py [nx ny nz] [--iters K]
Prints one JSON line (not the driver's bench contract - a profile for profiles/)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from block_helpers import box_addr, pu_matrix
from multiregionfoam_b200 import blockldu, ldu

args = [a for a in sys.argv[1:] if not a.startswith("--")]
nx, ny, nz = (int(a) for a in args[:3]) if len(args) >= 3 else (160, 128, 100)
iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 30
peak = 6540.5
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

t0 = time.time()
n, l, u = box_addr(nx, ny, nz)
M = pu_matrix(n, l, u)
F = int(M.l.size)
print(f"assembled p-U block system: {n} cells, {F} faces in {time.time() - t0:.1f} s", file=sys.stderr)
ctx = ldu.Context(0)
S = blockldu.BlockSystem(ctx, M.l, M.u, n)
S.set_coeffs(M.diag, M.upper, M.lower)
S.upload(M.psi, M.source)
S.x_save()
opts = dict(solver=blockldu.SOLVER_BICGSTAB, precond=ldu.PRECOND_CHOLESKY, tolerance=0.0, minIter=iters, maxIter=iters)
for _ in range(2):
    S.x_restore()
    S.solve_resident(**opts)
S.set_profiling(True)
S.kernel_times(reset=True)
reps, ms = 3, 0.0
for _ in range(reps):
    S.x_restore()
    p = S.solve_resident(**opts)
    ms += p["deviceMs"]
kt = S.kernel_times()
S.set_profiling(False)
ms /= reps


def kind_doubles(a):
    return 1 if a.ndim == 1 else (4 if a.ndim == 2 else 16)


dk, uk = kind_doubles(np.asarray(M.diag)), kind_doubles(np.asarray(M.upper))
# algorithmic bytes: Amul reads diag (8 dk N) + x (32 N), writes y (32 N), reads upper + lower (2 * 8 uk F) + addressing (8 F)
amul_bytes = (8 * dk + 64) * n + (16 * uk + 8) * F
out = {"workload": f"p-U block system (vector4), box {nx}x{ny}x{nz}", "cells": n, "faces": F, "solver": "BlockBiCGStab",
       "preconditioner": "BlockCholesky", "iterations": iters, "diag_kind": dk, "offdiag_kind": uk,
       "cell_iterations_per_s": n * iters / (ms * 1e-3), "ms_per_solve": ms, "final_residual": [float(v) for v in p["finalResidual"]],
       "kernels": {}}
for k, (tms, cnt) in kt.items():
    if cnt:
        e = {"ms_per_launch": tms / cnt, "launches_per_solve": cnt / reps}
        if k == "amul":
            e["gbs"] = amul_bytes / (tms / cnt * 1e-3) / 1e9
            e["frac_of_hbm_peak"] = e["gbs"] / peak
            e["algorithmic_bytes_per_launch"] = amul_bytes
        out["kernels"][k] = e
# CPU oracle (port) on a bounded number of iterations of the same system
if "--no-cpu" not in sys.argv:
    from oracle import pyblk
    O = pyblk.BlockOracle(M.l, M.u, n, M.diag, M.upper, M.lower)
    ci = max(2, min(iters, 5))
    t0 = time.time()
    O.solve(M.psi, M.source, "BiCGStab", "Cholesky", tolerance=0.0, minIter=ci, maxIter=ci)
    dt = time.time() - t0
    out["cpu_baseline"] = {"value": n * ci / dt, "unit": "cell-iterations/s", "cores": 1, "kind": "port",
                           "sample": f"{ci} BlockBiCGStab + BlockCholesky iterations of the same system, {dt:.1f} s"}
out["hbm_peak_gbs"] = peak
print(json.dumps(out))
