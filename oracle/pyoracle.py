"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package ``multiregionfoam_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libldu_oracle.so")

SOLVERS = {"PCG": 0, "CG": 0, "BiCGStab": 1, "PBiCGStab": 1, "PBiCG": 2, "BiCG": 2}
PRECONDS = {"none": 0, "diagonal": 1, "DIC": 2, "FDIC": 2, "DILU": 3, "Cholesky": 4}


class OrcOpts(C.Structure):
    _fields_ = [("solver", C.c_int), ("precond", C.c_int), ("tolerance", C.c_double),
                ("relTol", C.c_double), ("minIter", C.c_int), ("maxIter", C.c_int)]


class OrcPerf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double),
                ("nIterations", C.c_int), ("converged", C.c_int), ("singular", C.c_int),
                ("normFactor", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ldu_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def set_threads(n: int) -> None:
    """Threads of the oracle's row loops (stand-ins for MPI ranks); results are independent of the count."""
    lib().orc_set_threads(int(n))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_row.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
        L.orc_set_coeffs.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        L.orc_add_iface.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, ip, dp, dp, C.c_int, C.c_int, ip, ip, dp]
        L.orc_set_iface_zone.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, ip, ip, ip]
        L.orc_total_cells.argtypes = [C.c_void_p]
        for f in ("orc_amul", "orc_tmul", "orc_precondition", "orc_preconditionT"):
            getattr(L, f).argtypes = [C.c_void_p, dp, dp]
        L.orc_sumA.argtypes = [C.c_void_p, dp]
        L.orc_get_rD.argtypes = [C.c_void_p, dp]
        L.orc_residual.argtypes = [C.c_void_p, dp, dp, dp]
        L.orc_precond_setup.argtypes = [C.c_void_p, C.c_int]
        L.orc_solve.argtypes = [C.c_void_p, C.POINTER(OrcOpts), dp, dp, C.POINTER(OrcPerf), dp, C.c_int]
        L.orc_gs_smooth.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, dp, dp, C.c_int]
        L.orc_gs_smooth.restype = None
        L.orc_gs_solve.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.POINTER(OrcPerf), dp, C.c_int]
        L.orc_gsumprod.restype = C.c_double
        L.orc_gsumprod.argtypes = [C.c_void_p, dp, dp]
        L.orc_gsummag.restype = C.c_double
        L.orc_gsummag.argtypes = [C.c_void_p, dp]
        L.orc_set_reduction_mode.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_threads(1)
        L.orc_ggi_interpolate.argtypes = [C.c_int, ip, ip, dp, dp, C.c_int, dp]
        L.orc_patch_face_to_global.argtypes = [C.c_int, ip, ip, dp, C.c_int, C.c_int, dp]
        L.orc_global_face_to_patch.argtypes = [C.c_int, ip, dp, C.c_int, dp]
        L.orc_direct_map.argtypes = [C.c_int, ip, dp, C.c_int, dp]
        L.orc_version.restype = C.c_char_p
        _lib = L
    return _lib


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class OracleSystem:
    """The coupled system of a ``multiregionfoam_b200.case.Case`` (all ranks) inside the oracle.
    Vectors are concatenated in (rank, region) row order."""

    def __init__(self, case):
        L = lib()
        self.case = case
        nReg = case.nRegions
        self.nRows = case.nRanks * nReg
        self.h = L.orc_create(self.nRows, case.nRanks)
        for rk in case.ranks:
            for ri, reg in enumerate(rk.regions):
                row = rk.rank * nReg + ri
                rc = L.orc_set_row(self.h, row, rk.rank, ri, reg.nCells, reg.nFaces,
                                   _ip(_i32(reg.lowerAddr)), _ip(_i32(reg.upperAddr)))
                if rc:
                    raise RuntimeError(f"orc_set_row failed rc={rc}")
                L.orc_set_coeffs(self.h, row, _dp(_f64(reg.diag)), _dp(_f64(reg.upper)),
                                 None if reg.lower is None else _dp(_f64(reg.lower)))
        for rk in case.ranks:
            for ri, reg in enumerate(rk.regions):
                row = rk.rank * nReg + ri
                for itf in reg.interfaces:
                    peerRow = itf.peerRank * nReg + itf.peerRegion
                    go = None if itf.ggiOffsets is None else _i32(itf.ggiOffsets)
                    ga = None if itf.ggiAddr is None else _i32(itf.ggiAddr)
                    gw = None if itf.ggiWeights is None else _f64(itf.ggiWeights)
                    rc = L.orc_add_iface(self.h, row, itf.kind, itf.nFaces, _ip(_i32(itf.faceCells)),
                                         _dp(_f64(itf.bouCoeffs)), _dp(_f64(itf.intCoeffs)), peerRow,
                                         itf.peerIface, _ip(go), _ip(ga), _dp(gw))
                    if rc < 0:
                        raise RuntimeError(f"orc_add_iface failed rc={rc}")
        # shadow patches spread over ranks (decompose.py): the zone table, once every row has its interfaces
        for rk in case.ranks:
            for ri, reg in enumerate(rk.regions):
                for ii, itf in enumerate(reg.interfaces):
                    if not getattr(itf, "pieces", None):
                        continue
                    nZone = int(itf.nPeerFaces)
                    zr, zi, zp = (np.full(nZone, -1, np.int32) for _ in range(3))
                    for (h, pr, pi, za) in itf.pieces:
                        zr[za], zi[za], zp[za] = h * nReg + pr, pi, np.arange(len(za), dtype=np.int32)
                    rc = L.orc_set_iface_zone(self.h, rk.rank * nReg + ri, ii, nZone, _ip(zr), _ip(zi), _ip(zp))
                    if rc:
                        raise RuntimeError(f"orc_set_iface_zone failed rc={rc}")
        self.n = L.orc_total_cells(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_coeffs(self, rank, region, diag, upper, lower):
        row = rank * self.case.nRegions + region
        lib().orc_set_coeffs(self.h, row, _dp(_f64(diag)), _dp(_f64(upper)),
                             None if lower is None else _dp(_f64(lower)))

    def _vv(self, fn, x):
        x = _f64(x)
        y = np.empty(self.n)
        rc = getattr(lib(), fn)(self.h, _dp(x), _dp(y))
        if rc:
            raise RuntimeError(f"{fn} rc={rc}")
        return y

    def amul(self, x):
        return self._vv("orc_amul", x)

    def tmul(self, x):
        return self._vv("orc_tmul", x)

    def sumA(self):
        y = np.empty(self.n)
        lib().orc_sumA(self.h, _dp(y))
        return y

    def residual(self, x, b):
        x, b = _f64(x), _f64(b)
        r = np.empty(self.n)
        rc = lib().orc_residual(self.h, _dp(x), _dp(b), _dp(r))
        if rc:
            raise RuntimeError(f"orc_residual rc={rc}")
        return r

    def precond_setup(self, name: str):
        rc = lib().orc_precond_setup(self.h, PRECONDS[name])
        if rc:
            raise RuntimeError(f"orc_precond_setup rc={rc}")

    def rD(self):
        y = np.empty(self.n)
        lib().orc_get_rD(self.h, _dp(y))
        return y

    def precondition(self, r):
        return self._vv("orc_precondition", r)

    def preconditionT(self, r):
        return self._vv("orc_preconditionT", r)

    def set_reduction_mode(self, mode: int):
        """0: sequential sums (the reference, default); 1: pairwise sums (sensitivity measurement only)."""
        lib().orc_set_reduction_mode(self.h, mode)

    def gsumprod(self, a, b):
        return lib().orc_gsumprod(self.h, _dp(_f64(a)), _dp(_f64(b)))

    def gsummag(self, a):
        return lib().orc_gsummag(self.h, _dp(_f64(a)))

    def solve(self, x0, b, solver="BiCGStab", precond="DILU", tolerance=1e-6, relTol=0.0,
              minIter=0, maxIter=1000, historyCap=None):
        x = _f64(x0).copy()
        b = _f64(b)
        cap = (maxIter + 1) if historyCap is None else historyCap
        hist = np.full(cap, np.nan)
        opts = OrcOpts(SOLVERS[solver], PRECONDS[precond], tolerance, relTol, minIter, maxIter)
        perf = OrcPerf()
        rc = lib().orc_solve(self.h, C.byref(opts), _dp(x), _dp(b), C.byref(perf), _dp(hist), cap)
        if rc:
            raise RuntimeError(f"orc_solve rc={rc}")
        info = dict(initialResidual=perf.initialResidual, finalResidual=perf.finalResidual,
                    nIterations=perf.nIterations, converged=bool(perf.converged),
                    singular=bool(perf.singular), normFactor=perf.normFactor,
                    history=hist[:min(cap, perf.nIterations + 1)].copy())
        return x, info


def ggi_interpolate(offsets, addr, weights, ff, nComp=1):
    offsets, addr, weights, ff = _i32(offsets), _i32(addr), _f64(weights), _f64(ff)
    nTo = offsets.size - 1
    out = np.empty(nTo * nComp)
    lib().orc_ggi_interpolate(nTo, _ip(offsets), _ip(addr), _dp(weights), _dp(ff), nComp, _dp(out))
    return out.reshape(nTo, nComp) if nComp > 1 else out


def patch_face_to_global(pieceOffsets, faceToGlobalAddr, pField, nZoneFaces, nComp=1):
    po, addr, pf = _i32(pieceOffsets), _i32(faceToGlobalAddr), _f64(pField)
    out = np.empty(nZoneFaces * nComp)
    lib().orc_patch_face_to_global(po.size - 1, _ip(po), _ip(addr), _dp(pf), nComp, nZoneFaces, _dp(out))
    return out.reshape(nZoneFaces, nComp) if nComp > 1 else out


def global_face_to_patch(faceToGlobalAddr, gField, nComp=1):
    addr, g = _i32(faceToGlobalAddr), _f64(gField)
    out = np.empty(addr.size * nComp)
    lib().orc_global_face_to_patch(addr.size, _ip(addr), _dp(g), nComp, _dp(out))
    return out.reshape(addr.size, nComp) if nComp > 1 else out


def direct_map_build(to, from_, tol):
    """directMapInterfaceToInterfaceMapping.C:155-168: -> map (first match, -1 if none), number unmatched."""
    t, f = _f64(to), _f64(from_)
    m = np.empty(t.size // 3, np.int32)
    L = lib()
    L.orc_direct_map_build.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_int)]
    n = L.orc_direct_map_build(m.size, _dp(t), f.size // 3, _dp(f), float(tol), _ip(m))
    return m, n


def direct_map(map_, from_, nComp=1):
    m, f = _i32(map_), _f64(from_)
    out = np.empty(m.size * nComp)
    lib().orc_direct_map(m.size, _ip(m), _dp(f), nComp, _dp(out))
    return out.reshape(m.size, nComp) if nComp > 1 else out


# ---- GaussSeidelSmoother / smoothSolver for one matrix
def gs_smooth(lowerAddr, upperAddr, diag, upper, lower, psi, source, nSweeps=1):
    """GaussSeidelSmoother::smooth: nSweeps sweeps; coupled-patch contributions are expected inside `source`."""
    l, u = _i32(lowerAddr), _i32(upperAddr)
    d, up = _f64(diag), _f64(upper)
    lo = None if lower is None else _f64(lower)
    x = np.array(psi, np.float64).copy()
    lib().orc_gs_smooth(d.size, l.size, _ip(l), _ip(u), _dp(d), _dp(up), _dp(lo), _dp(x), _dp(_f64(source)), int(nSweeps))
    return x


def gs_solve(lowerAddr, upperAddr, diag, upper, lower, psi, source, nSweeps=1, tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000):
    """smoothSolver::solve with a GaussSeidel smoother -> (psi, dict like OracleSystem.solve)."""
    l, u = _i32(lowerAddr), _i32(upperAddr)
    d, up = _f64(diag), _f64(upper)
    lo = None if lower is None else _f64(lower)
    x = np.array(psi, np.float64).copy()
    p = OrcPerf()
    cap = maxIter // max(1, nSweeps) + 3
    hist = np.full(cap, np.nan)
    rc = lib().orc_gs_solve(d.size, l.size, _ip(l), _ip(u), _dp(d), _dp(up), _dp(lo), _dp(x), _dp(_f64(source)), int(nSweeps),
                            tolerance, relTol, minIter, maxIter, C.byref(p), _dp(hist), cap)
    if rc:
        raise RuntimeError(f"orc_gs_solve failed rc={rc}")
    n = p.nIterations // max(1, nSweeps)
    return x, dict(initialResidual=p.initialResidual, finalResidual=p.finalResidual, nIterations=p.nIterations,
                   converged=bool(p.converged), normFactor=p.normFactor, history=hist[: n + 1].copy())
