"""ctypes binding of the block-coupled (vector4) CPU oracle, oracle/blk_oracle.c (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libblk_oracle.so")

SOLVERS = {"CG": 0, "BiCGStab": 1}
PRECONDS = {"none": 0, "diagonal": 1, "Cholesky": 4}


class BlkOpts(C.Structure):
    _fields_ = [("solver", C.c_int), ("precond", C.c_int), ("tolerance", C.c_double), ("relTol", C.c_double),
                ("minIter", C.c_int), ("maxIter", C.c_int)]


class BlkPerf(C.Structure):
    _fields_ = [("initialResidual", C.c_double * 4), ("finalResidual", C.c_double * 4), ("nIterations", C.c_int),
                ("converged", C.c_int), ("singular", C.c_int), ("normFactor", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "blk_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.blk_create.restype = C.c_void_p
        L.blk_create.argtypes = [C.c_int, C.c_int, ip, ip]
        L.blk_destroy.argtypes = [C.c_void_p]
        L.blk_set_coeffs.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, dp, C.c_int, dp]
        L.blk_amul.argtypes = [C.c_void_p, dp, dp]
        L.blk_precond_setup.argtypes = [C.c_void_p, C.c_int]
        L.blk_get_precon_diag.argtypes = [C.c_void_p, dp, ip]
        L.blk_precondition.argtypes = [C.c_void_p, dp, dp]
        L.blk_solve.argtypes = [C.c_void_p, C.POINTER(BlkOpts), dp, dp, C.POINTER(BlkPerf), dp, C.c_int]
        L.blk_gsumprod.restype = C.c_double
        L.blk_gsumprod.argtypes = [C.c_void_p, dp, dp]
        L.blk_gsumcmptmag.argtypes = [C.c_void_p, dp, dp]
        L.blk_norm_factor.restype = C.c_double
        L.blk_norm_factor.argtypes = [C.c_void_p, dp, dp]
        L.blk_inv4.argtypes = [dp, dp]
        L.blk_set_reduction_mode.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _kind(a: np.ndarray, n: int) -> int:
    k = 1 if a.ndim == 1 else int(np.prod(a.shape[1:]))
    assert k in (1, 4, 16), f"coefficient kind {k}"
    return k


def inv4(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float64).reshape(16)
    out = np.empty(16)
    lib().blk_inv4(_dp(a), _dp(out))
    return out.reshape(4, 4)


class BlockOracle:
    """One block (vector4) LDU system: diag [N], [N,4] or [N,4,4]; upper/lower [F], [F,4] or [F,4,4]; lower None =
    symmetric (lower triangle = transposed upper)."""

    def __init__(self, lowerAddr, upperAddr, nCells, diag, upper, lower=None):
        L = lib()
        self.n = int(nCells)
        self.l = np.ascontiguousarray(lowerAddr, np.int32)
        self.u = np.ascontiguousarray(upperAddr, np.int32)
        self.nf = int(self.l.size)
        self.h = L.blk_create(self.n, self.nf, _ip(self.l), _ip(self.u))
        self.set_coeffs(diag, upper, lower)

    def set_coeffs(self, diag, upper, lower=None):
        d = np.ascontiguousarray(diag, np.float64)
        u = np.ascontiguousarray(upper, np.float64)
        lo = None if lower is None else np.ascontiguousarray(lower, np.float64)
        rc = lib().blk_set_coeffs(self.h, _kind(d, self.n), _dp(d), _kind(u, self.nf), _dp(u),
                                  0 if lo is None else _kind(lo, self.nf), _dp(lo))
        if rc:
            raise RuntimeError("blk_set_coeffs failed")

    def close(self):
        if self.h:
            lib().blk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_reduction_mode(self, mode: int):
        """0: sequential sums (reference); 1: pairwise dot products (summation-order sensitivity probe)."""
        lib().blk_set_reduction_mode(self.h, int(mode))

    def amul(self, x):
        x = np.ascontiguousarray(x, np.float64).reshape(self.n, 4)
        y = np.empty_like(x)
        lib().blk_amul(self.h, _dp(x), _dp(y))
        return y

    def precon_diag(self, precond="Cholesky"):
        lib().blk_precond_setup(self.h, PRECONDS[precond])
        out = np.empty(self.n * 16)
        k = C.c_int(0)
        rc = lib().blk_get_precon_diag(self.h, _dp(out), C.byref(k))
        if rc:
            raise RuntimeError("no preconditioner diagonal")
        return out[: self.n * k.value].reshape(self.n, k.value).copy()

    def precondition(self, r, precond="Cholesky"):
        lib().blk_precond_setup(self.h, PRECONDS[precond])
        r = np.ascontiguousarray(r, np.float64).reshape(self.n, 4)
        w = np.empty_like(r)
        lib().blk_precondition(self.h, _dp(r), _dp(w))
        return w

    def sumprod(self, a, b):
        return lib().blk_gsumprod(self.h, _dp(np.ascontiguousarray(a, np.float64)), _dp(np.ascontiguousarray(b, np.float64)))

    def norm_factor(self, x, b):
        return lib().blk_norm_factor(self.h, _dp(np.ascontiguousarray(x, np.float64)), _dp(np.ascontiguousarray(b, np.float64)))

    def solve(self, x0, b, solver="BiCGStab", precond="Cholesky", tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000):
        x = np.array(x0, np.float64).reshape(self.n, 4).copy()
        b = np.ascontiguousarray(b, np.float64).reshape(self.n, 4)
        o = BlkOpts(SOLVERS[solver], PRECONDS[precond], tolerance, relTol, minIter, maxIter)
        p = BlkPerf()
        cap = maxIter + 2
        hist = np.empty((cap, 4))
        rc = lib().blk_solve(self.h, C.byref(o), _dp(x), _dp(b), C.byref(p), _dp(hist), cap)
        if rc:
            raise RuntimeError("blk_solve failed")
        return x, dict(initialResidual=np.array(p.initialResidual[:]), finalResidual=np.array(p.finalResidual[:]),
                       nIterations=p.nIterations, converged=bool(p.converged), singular=bool(p.singular),
                       normFactor=p.normFactor, history=hist[: p.nIterations + 1].copy())
