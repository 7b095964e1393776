"""Gauss-Seidel smoother / smoothSolver for one lduMatrix on the device (include/b200_smooth.h; SURVEY 8(f) rank 4, first piece).

Mirrors foam-extend's run-time selection:  ``solver smoothSolver; smoother GaussSeidel; nSweeps n;`` (the `smoothSolver`
entries of the tutorials' fvSolution files).  No CPU fallback: everything goes through libb200ldu.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ldu

# every symbol include/b200_smooth.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = ["b200_gs_create", "b200_gs_destroy", "b200_gs_set_coeffs", "b200_gs_sweep", "b200_gs_smooth", "b200_gs_solve"]

# lduMatrix::smoother names served by the library: Gauss-Seidel (this module, b200_gs_*), the DIC / DILU family
# (ldu.LduSystem.smooth = b200_smooth: the preconditioner's sweeps applied to the residual) and their combination
SMOOTHER_TABLE = {
    "GaussSeidel": "gs", "cudaGaussSeidel": "gs",
    "DIC": "dic", "cudaDIC": "dic", "DILU": "dilu", "cudaDILU": "dilu",
    "DICGaussSeidel": "dic+gs", "cudaDICGaussSeidel": "dic+gs",
}

_bound = False


def _lib():
    global _bound
    L = ldu.load()
    if not _bound:
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
        L.b200_gs_create.argtypes = [vp, C.c_int32, C.c_int32, ip, ip, C.POINTER(vp)]
        L.b200_gs_destroy.argtypes = [vp]
        L.b200_gs_set_coeffs.argtypes = [vp, dp, dp, dp]
        L.b200_gs_sweep.argtypes = [vp, dp, dp]
        L.b200_gs_smooth.argtypes = [vp, dp, dp, C.c_int]
        L.b200_gs_solve.argtypes = [vp, C.POINTER(ldu.SolverOpts), C.c_int, dp, dp, C.POINTER(ldu.Perf), dp, C.c_int]
        _bound = True
    return L


class GaussSeidel:
    """GaussSeidelSmoother of one lduMatrix (lowerAddr / upperAddr in upper-triangular order)."""

    def __init__(self, ctx: ldu.Context, lowerAddr, upperAddr, nCells: int):
        L = _lib()
        self.ctx, self.nCells = ctx, int(nCells)
        l, u = ldu._i32(lowerAddr), ldu._i32(upperAddr)
        self.h = C.c_void_p()
        ctx.check(L.b200_gs_create(ctx.h, self.nCells, int(l.size), ldu._ip(l), ldu._ip(u), C.byref(self.h)))

    def close(self):
        if self.h:
            _lib().b200_gs_destroy(self.h)
            self.h = C.c_void_p()

    def set_coeffs(self, diag, upper, lower=None):
        d, u = ldu._f64(diag), ldu._f64(upper)
        lo = None if lower is None else ldu._f64(lower)
        self.ctx.check(_lib().b200_gs_set_coeffs(self.h, ldu._dp(d), ldu._dp(u), ldu._dp(lo)))

    def sweep(self, psi, bPrime) -> np.ndarray:
        """One sweep with the caller's bPrime (source + coupled-patch contributions)."""
        x = np.array(ldu._f64(psi), copy=True)
        self.ctx.check(_lib().b200_gs_sweep(self.h, ldu._dp(x), ldu._dp(ldu._f64(bPrime))))
        return x

    def smooth(self, psi, source, nSweeps: int = 1) -> np.ndarray:
        x = np.array(ldu._f64(psi), copy=True)
        self.ctx.check(_lib().b200_gs_smooth(self.h, ldu._dp(x), ldu._dp(ldu._f64(source)), int(nSweeps)))
        return x

    def solve(self, psi, source, nSweeps: int = 1, tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000):
        """smoothSolver::solve -> (psi, dict(initialResidual, finalResidual, nIterations, converged, normFactor, history))."""
        x = np.array(ldu._f64(psi), copy=True)
        o = ldu.SolverOpts(0, 0, tolerance, relTol, minIter, maxIter)
        p = ldu.Perf()
        cap = maxIter // max(1, nSweeps) + 3
        hist = np.full(cap, np.nan)
        self.ctx.check(_lib().b200_gs_solve(self.h, C.byref(o), int(nSweeps), ldu._dp(x), ldu._dp(ldu._f64(source)), C.byref(p), ldu._dp(hist), cap))
        k = p.nIterations // max(1, nSweeps)
        return x, dict(initialResidual=p.initialResidual, finalResidual=p.finalResidual, nIterations=p.nIterations,
                       converged=bool(p.converged), normFactor=p.normFactor, deviceMs=p.deviceMs, history=hist[: k + 1].copy())
