"""Measurement of the device-side T-equation assembly next to bench.py's e2e leg: one time step of
the two-region CHT case (default C2) driven (a) the way an unmodified caller does - host matrix in, b200_sys_set_coeffs +
b200_solve with host buffers - and (b) with b200_sys_assemble_T from the resident field + b200_solve_resident + the D2H
of the solution.  Also times the two assembly kernels alone (CUDA events, 'pack' class) against their algorithmic bytes.
usage: python scripts/bench_assemble.py [--workload C2] [--iters 50] [--steps 5]
Prints one JSON line (not the driver's bench contract - a profile for profiles/)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import WORKLOADS, assemble_cht, cht_fv_tables
from multiregionfoam_b200.mesh import flow_over_heated_plate

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
peak = 6540.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

r, L = WORKLOADS[a.workload]
fluid, solid = flow_over_heated_plate(r, L)
case = assemble_cht(fluid, solid)
rs = case.ranks[0]
tables = cht_fv_tables(fluid, solid)
N = sum(reg.nCells for reg in rs.regions)
F = sum(reg.nFaces for reg in rs.regions)
ctx = ldu.Context(0)
kw = dict(solver=ldu.SOLVER_BICGSTAB, precond=ldu.PRECOND_DILU, tolerance=0.0, minIter=a.iters, maxIter=a.iters)
x0, b = case.concat("psi"), case.concat("source")

# page-locked host buffers on both legs (as bench.py's e2e leg)
import ctypes as C
Lib = ldu.load()
keep = []
for reg in rs.regions:
    keep += [arr for arr in (reg.diag, reg.upper, reg.lower, reg.source) if arr is not None]
    for itf in reg.interfaces:
        keep += [itf.bouCoeffs, itf.intCoeffs]
hx = [np.ascontiguousarray(reg.psi.copy()) for reg in rs.regions]      # the caller's field: guess in, solution out
hx0 = [np.ascontiguousarray(reg.psi.copy()) for reg in rs.regions]
keep += hx + hx0
for arr in keep:
    ctx.host_register(arr)
opts, perf = ldu.SolverOpts(kw["solver"], kw["precond"], 0.0, 0.0, a.iters, a.iters), ldu.Perf()
bs = ldu._dpp([reg.source for reg in rs.regions])

# (a) host-assembled: matrix, x, b H2D; solve; x D2H
H = ldu.LduSystem(ctx, rs)
def host_step():
    for h, h0 in zip(hx, hx0):
        h[...] = h0
    H.set_all_coeffs()
    ctx.check(Lib.b200_solve(H.h, C.byref(opts), ldu._dpp(hx), bs, C.byref(perf), None, 0))
    return perf.nIterations
# (b) device-assembled from a host field: x H2D, assemble, solve, x D2H
S = ldu.LduSystem(ctx, rs, set_coeffs=False)
for ri, t in enumerate(tables):
    S.set_fv_geometry(ri, t["V"], t["magSf"], t["deltaCoeffs"], t["bCells"], t["bInt"], t["bSrc"])
    for i, itf in enumerate(rs.regions[ri].interfaces):
        S.set_interface_coeffs(ri, i, itf.bouCoeffs, itf.intCoeffs)
S.upload(x0, None)
S.x_save()
for ri, t in enumerate(tables):
    S.assemble_T(ri, t["form"], t["rhoC"], t["rDeltaT"], t["kappa"], phi=t["phi"])
def dev_step(resident=False):
    if resident:
        S.x_restore()                      # device-side copy: the field of the previous step is already there
    else:
        ctx.check(Lib.b200_upload(S.h, ldu._dpp(hx0), None))
    for ri, t in enumerate(tables):
        S.assemble_T(ri, t["form"], t["rhoC"], t["rDeltaT"], t["kappa"])
    ctx.check(Lib.b200_solve_resident(S.h, C.byref(opts), C.byref(perf), None, 0))
    ctx.check(Lib.b200_download(S.h, ldu._dpp(hx)))
    return perf.nIterations

out = {"workload": a.workload, "cells": N, "faces": F, "iterations_per_step": a.iters}
for name, step in (("host_assembled", host_step), ("device_assembled", dev_step),
                   ("device_assembled_resident_field", lambda: dev_step(True))):
    for _ in range(2):
        step()
    t0, its = time.perf_counter(), 0
    for _ in range(a.steps):
        its += step()
    dt = time.perf_counter() - t0
    out[name] = {"ms_per_step": 1e3 * dt / a.steps, "cell_iterations_per_s": N * its / dt}
# the assembly kernels alone
S.set_profiling(True)
S.kernel_times(reset=True)
reps = 10
for _ in range(reps):
    for ri, t in enumerate(tables):
        S.assemble_T(ri, t["form"], t["rhoC"], t["rDeltaT"], t["kappa"])
S.download()
pack_ms, pack_launches = S.kernel_times(reset=True)["pack"]
# algorithmic bytes per step: faces 16 (geometry) + 16 (upper, lower) [+ 8 phi in the fluid]; cells 8 V + 12 row pointers + 4 slot + 8 x + 16 out
alg = sum(reg.nFaces * (32 + (8 if t["phi"] is not None else 0)) + reg.nCells * 48 for reg, t in zip(rs.regions, tables))
asm_ms = pack_ms / reps
out["assemble_kernels"] = {"ms_per_step": asm_ms, "launches_per_step": (pack_launches - 1) / reps, "algorithmic_bytes": alg,
                           "gbs": alg / (asm_ms * 1e-3) / 1e9 if asm_ms > 0 else None, "peak_gbs": peak,
                           "note": "the 'pack' class also holds the D2H permutation of the final download (1 of 41 launches)"}
out["host_memory"] = "page-locked (b200_host_register) on all legs"
print(json.dumps(out))
