"""ctypes binding of the C ABI in include/b200_ldu.h (libb200ldu.so) plus thin host objects.

Everything here goes through the ``extern "C"`` entry points with plain pointers and sizes -- the
same calls a foam-extend adapter makes.  There is no CPU fallback: if the library
has not been built, or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .build import LIB_PATH
from .case import RankSystem

SOLVER_PCG, SOLVER_BICGSTAB, SOLVER_PBICG = 0, 1, 2
PRECOND_NONE, PRECOND_DIAGONAL, PRECOND_DIC, PRECOND_DILU, PRECOND_CHOLESKY = 0, 1, 2, 3, 4
KERNEL_CLASSES = ["amul", "iface", "sweep_fwd", "sweep_bwd", "vector", "reduce", "pack", "halo"]

# every symbol include/b200_ldu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "b200_nccl_unique_id", "b200_ctx_create", "b200_ctx_destroy", "b200_last_error", "b200_version",
    "b200_sys_create", "b200_sys_destroy", "b200_sys_set_region", "b200_sys_add_interface", "b200_sys_finalize",
    "b200_sys_set_coeffs", "b200_sys_set_interface_coeffs", "b200_sys_num_cells", "b200_sys_num_faces",
    "b200_solve", "b200_upload", "b200_solve_resident", "b200_download",
    "b200_x_save", "b200_x_restore", "b200_host_register", "b200_host_unregister",
    "b200_amul", "b200_precondition", "b200_get_rD", "b200_reduce", "b200_residual", "b200_sum_a", "b200_smooth",
    "b200_set_profiling", "b200_get_kernel_times", "b200_launch_count", "b200_debug_sweep_stats",
    "b200_ggi_interpolate", "b200_patch_face_to_global", "b200_global_face_to_patch",
    "b200_sys_set_interface_attached", "b200_sys_set_interface_ggi", "b200_sys_set_interface_pieces",
    "b200_sys_set_fv_geometry", "b200_sys_assemble_T",
    "b200_ggi_build", "b200_ggi_fetch", "b200_direct_map_build", "b200_direct_map_transfer",
]

TEQN_CONDUCT, TEQN_TRANSPORT = 0, 1


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200ldu error {code}: {msg}")
        self.code = code


class SolverOpts(C.Structure):
    _fields_ = [("solver", C.c_int), ("precond", C.c_int), ("tolerance", C.c_double), ("relTol", C.c_double),
                ("minIter", C.c_int), ("maxIter", C.c_int)]


class Perf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double), ("nIterations", C.c_int),
                ("converged", C.c_int), ("singular", C.c_int), ("normFactor", C.c_double), ("deviceMs", C.c_double)]


_lib = None
SWEEP_STATS_STRIDE = 16 + 8 * 142  # kStatsStride of csrc/kernels.cuh: 16 totals + 8 time stamps per traced block


def load():
    """dlopen the in-tree library; raises if it is missing (run ``python -m multiregionfoam_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("B200_LDU_LIB", LIB_PATH)  # developer knob: compare differently compiled builds
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not built: run __graft_entry__.build() / python -m multiregionfoam_b200.build "
            "(the CUDA library is the only compute path; there is no CPU fallback)")
    L = C.CDLL(path)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    dpp = C.POINTER(dp)
    L.b200_nccl_unique_id.argtypes = [vp]
    L.b200_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.b200_ctx_destroy.argtypes = [vp]
    L.b200_last_error.argtypes = [vp]
    L.b200_last_error.restype = C.c_char_p
    L.b200_sys_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.b200_sys_destroy.argtypes = [vp]
    L.b200_sys_set_region.argtypes = [vp, C.c_int, C.c_int32, C.c_int32, ip, ip]
    L.b200_sys_add_interface.argtypes = [vp, C.c_int, C.c_int, C.c_int32, ip, C.c_int, C.c_int, C.c_int, C.c_int32, ip, ip, dp]
    L.b200_sys_finalize.argtypes = [vp]
    L.b200_sys_set_coeffs.argtypes = [vp, C.c_int, dp, dp, dp]
    L.b200_sys_set_interface_coeffs.argtypes = [vp, C.c_int, C.c_int, dp, dp]
    L.b200_sys_set_interface_attached.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.b200_sys_set_interface_ggi.argtypes = [vp, C.c_int, C.c_int, C.c_int32, ip, ip, dp]
    L.b200_sys_set_interface_pieces.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip]
    L.b200_sys_set_fv_geometry.argtypes = [vp, C.c_int, dp, dp, dp, C.c_int32, ip, dp, dp]
    L.b200_sys_assemble_T.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp, dp]
    L.b200_sys_num_cells.argtypes = [vp]
    L.b200_sys_num_cells.restype = C.c_int64
    L.b200_sys_num_faces.argtypes = [vp]
    L.b200_sys_num_faces.restype = C.c_int64
    L.b200_solve.argtypes = [vp, C.POINTER(SolverOpts), dpp, dpp, C.POINTER(Perf), dp, C.c_int]
    L.b200_upload.argtypes = [vp, dpp, dpp]
    L.b200_solve_resident.argtypes = [vp, C.POINTER(SolverOpts), C.POINTER(Perf), dp, C.c_int]
    L.b200_download.argtypes = [vp, dpp]
    L.b200_x_save.argtypes = [vp]
    L.b200_x_restore.argtypes = [vp]
    L.b200_host_register.argtypes = [vp, vp, C.c_uint64]
    L.b200_host_unregister.argtypes = [vp, vp]
    L.b200_amul.argtypes = [vp, dpp, dpp, C.c_int]
    L.b200_precondition.argtypes = [vp, C.c_int, dpp, dpp, C.c_int]
    L.b200_residual.argtypes = [vp, dpp, dpp, dpp]
    L.b200_smooth.argtypes = [vp, C.c_int, C.c_int, dpp, dpp]
    L.b200_sum_a.argtypes = [vp, dpp]
    L.b200_get_rD.argtypes = [vp, C.c_int, dpp]
    L.b200_reduce.argtypes = [vp, dpp, dpp, dp]
    L.b200_set_profiling.argtypes = [vp, C.c_int]
    L.b200_get_kernel_times.argtypes = [vp, dp, C.POINTER(C.c_int64), C.c_int]
    L.b200_debug_sweep_stats.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_int]
    L.b200_launch_count.argtypes = [vp]
    L.b200_launch_count.restype = C.c_int64
    L.b200_ggi_build.argtypes = [vp, C.c_int32, ip, ip, C.c_int32, dp, C.c_int32, ip, ip, C.c_int32, dp, C.c_double, C.c_int]
    L.b200_ggi_fetch.argtypes = [vp, C.c_int32, ip, ip, dp]
    L.b200_direct_map_build.argtypes = [vp, C.c_int32, dp, C.c_int32, dp, C.c_double, ip]
    L.b200_direct_map_transfer.argtypes = [vp, C.c_int32, ip, C.c_int32, dp, C.c_int, dp]
    L.b200_ggi_interpolate.argtypes = [vp, C.c_int32, C.c_int32, ip, ip, dp, dp, C.c_int, dp]
    L.b200_patch_face_to_global.argtypes = [vp, C.c_int32, ip, dp, C.c_int, C.c_int32, dp]
    L.b200_global_face_to_patch.argtypes = [vp, C.c_int32, ip, dp, C.c_int, dp]
    _lib = L
    return L


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _dpp(arrs: Sequence[np.ndarray]):
    T = C.POINTER(C.c_double) * len(arrs)
    return T(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs])


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = load().b200_nccl_unique_id(buf)
    if rc:
        raise B200Error(rc, load().b200_last_error(None).decode())
    return buf.raw


class Context:
    """One per process / rank, bound to one GPU (b200_ctx)."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, unique_id: Optional[bytes] = None):
        L = load()
        self.h = C.c_void_p()
        uid = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        rc = L.b200_ctx_create(device, rank, nranks, uid, C.byref(self.h))
        if rc:
            raise B200Error(rc, L.b200_last_error(None).decode())
        self.device, self.rank, self.nranks = device, rank, nranks

    def check(self, rc: int) -> int:
        if rc < 0:
            raise B200Error(rc, load().b200_last_error(self.h).decode())
        return rc

    @property
    def launches(self) -> int:
        return int(load().b200_launch_count(self.h))

    def host_register(self, a: np.ndarray):
        self.check(load().b200_host_register(self.h, a.ctypes.data, a.nbytes))

    def host_unregister(self, a: np.ndarray):
        self.check(load().b200_host_unregister(self.h, a.ctypes.data))

    def close(self):
        if self.h:
            load().b200_ctx_destroy(self.h)
            self.h = C.c_void_p()

    # ---- partitioned-coupling face transfer
    def ggi_build(self, mFaceOffsets, mFaceLabels, mPoints, sFaceOffsets, sFaceLabels, sPoints, nonOverlapTol: float = 1e-15,
                  rescale: bool = True):
        """GGIInterpolation weights of the master faces from the slave patch (ggiInterfaceToInterfaceMapping.C:62-77:
        tolerances SMALL, rescale true): -> offsets, addr, weights as ggi_interpolate / set_interface_ggi take them."""
        mo, ml, so, sl = _i32(mFaceOffsets), _i32(mFaceLabels), _i32(sFaceOffsets), _i32(sFaceLabels)
        mp, sp = _f64(mPoints), _f64(sPoints)
        nM, nS = mo.size - 1, so.size - 1
        nnz = self.check(load().b200_ggi_build(self.h, nM, _ip(mo), _ip(ml), mp.size // 3, _dp(mp), nS, _ip(so), _ip(sl),
                                               sp.size // 3, _dp(sp), float(nonOverlapTol), int(bool(rescale))))
        off, addr, w = np.zeros(nM + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        self.check(load().b200_ggi_fetch(self.h, nM, _ip(off), _ip(addr), _dp(w)))
        return off, addr, w

    def direct_map_build(self, to, from_, tol: float, require_conformal: bool = True) -> np.ndarray:
        """directMapInterfaceToInterfaceMapping::calcZoneAToZoneBFaceMap & co. (directMapInterfaceToInterfaceMapping.C:147-181):
        for every location of ``to`` (n x 3) the first location of ``from_`` closer than tol; an unmatched location is the
        reference's FatalError ("Direct mapping can only be used with conformal interfaces!")."""
        t, f = _f64(to).reshape(-1, 3), _f64(from_).reshape(-1, 3)
        m = np.full(t.shape[0], -1, np.int32)
        unmatched = self.check(load().b200_direct_map_build(self.h, t.shape[0], _dp(t), f.shape[0], _dp(f), float(tol), _ip(m)))
        if unmatched and require_conformal:
            raise B200Error(-1, f"Cannot calculate the map between interfaces: {unmatched} locations have no counterpart "
                                "(direct mapping can only be used with conformal interfaces)")
        return m

    def direct_map_transfer(self, map_, from_) -> np.ndarray:
        """transferFacesZoneToZone / transferPointsZoneToZone of the directMap mapping: to[i] = from[map[i]]."""
        m, f = _i32(map_), _f64(from_)
        nComp = 1 if f.ndim == 1 else f.shape[1]
        out = np.empty((m.size, nComp))
        self.check(load().b200_direct_map_transfer(self.h, m.size, _ip(m), f.shape[0], _dp(f), nComp, _dp(out)))
        return out[:, 0] if f.ndim == 1 else out

    def ggi_interpolate(self, offsets, addr, weights, ff, nFrom: Optional[int] = None) -> np.ndarray:
        offsets, addr, weights = _i32(offsets), _i32(addr), _f64(weights)
        ff = _f64(ff)
        nComp = 1 if ff.ndim == 1 else ff.shape[1]
        nFrom = ff.shape[0] if nFrom is None else nFrom
        nTo = offsets.size - 1
        out = np.empty((nTo, nComp))
        self.check(load().b200_ggi_interpolate(self.h, nTo, nFrom, _ip(offsets), _ip(addr), _dp(weights), _dp(ff), nComp, _dp(out)))
        return out[:, 0] if ff.ndim == 1 else out

    def patch_face_to_global(self, faceToGlobalAddr, pField, nZoneFaces: int) -> np.ndarray:
        addr, pf = _i32(faceToGlobalAddr), _f64(pField)
        nComp = 1 if pf.ndim == 1 else pf.shape[1]
        out = np.empty((nZoneFaces, nComp))
        self.check(load().b200_patch_face_to_global(self.h, addr.size, _ip(addr), _dp(pf), nComp, nZoneFaces, _dp(out)))
        return out[:, 0] if pf.ndim == 1 else out

    def global_face_to_patch(self, faceToGlobalAddr, gField) -> np.ndarray:
        addr, g = _i32(faceToGlobalAddr), _f64(gField)
        nComp = 1 if g.ndim == 1 else g.shape[1]
        out = np.empty((addr.size, nComp))
        self.check(load().b200_global_face_to_patch(self.h, addr.size, _ip(addr), _dp(g), nComp, _dp(out)))
        return out[:, 0] if g.ndim == 1 else out


class LduSystem:
    """The coupledLduMatrix of one rank on the device (b200_sys): one lduMatrix per region plus
    its coupled interfaces.  Per-region host vectors in, per-region host vectors out."""

    def __init__(self, ctx: Context, rs: RankSystem, set_coeffs: bool = True):
        L = load()
        self.ctx = ctx
        self.rs = rs
        self.h = C.c_void_p()
        ctx.check(L.b200_sys_create(ctx.h, len(rs.regions), C.byref(self.h)))
        for r, reg in enumerate(rs.regions):
            l, u = _i32(reg.lowerAddr), _i32(reg.upperAddr)
            ctx.check(L.b200_sys_set_region(self.h, r, reg.nCells, reg.nFaces, _ip(l), _ip(u)))
        for r, reg in enumerate(rs.regions):
            for i, itf in enumerate(reg.interfaces):
                fc = _i32(itf.faceCells)
                go = None if itf.ggiOffsets is None else _i32(itf.ggiOffsets)
                ga = None if itf.ggiAddr is None else _i32(itf.ggiAddr)
                gw = None if itf.ggiWeights is None else _f64(itf.ggiWeights)
                nPeer = itf.nFaces if getattr(itf, "nPeerFaces", None) is None else itf.nPeerFaces
                got = ctx.check(L.b200_sys_add_interface(self.h, r, itf.kind, itf.nFaces, _ip(fc), itf.peerRank, itf.peerRegion,
                                                         itf.peerIface, nPeer, _ip(go), _ip(ga), _dp(gw)))
                assert got == i
                pieces = getattr(itf, "pieces", None)
                if pieces:   # shadow patch spread over ranks: (rank, region, iface, zoneAddr) per piece
                    off = np.zeros(len(pieces) + 1, np.int32)
                    off[1:] = np.cumsum([len(p[3]) for p in pieces])
                    za = _i32(np.concatenate([np.asarray(p[3], np.int32) for p in pieces]) if off[-1] else np.zeros(0, np.int32))
                    pr, pg, pi = (_i32([p[k] for p in pieces]) for k in range(3))
                    ctx.check(L.b200_sys_set_interface_pieces(self.h, r, i, len(pieces), _ip(pr), _ip(pg), _ip(pi), _ip(off), _ip(za)))
        ctx.check(L.b200_sys_finalize(self.h))
        self.sizes = [reg.nCells for reg in rs.regions]
        if set_coeffs:
            self.set_all_coeffs()

    # ---- coefficients
    def set_coeffs(self, r: int, diag, upper, lower=None):
        d, u = _f64(diag), _f64(upper)
        lo = None if lower is None else _f64(lower)
        self.ctx.check(load().b200_sys_set_coeffs(self.h, r, _dp(d), _dp(u), _dp(lo)))

    def set_interface_coeffs(self, r: int, i: int, bou, int_=None):
        b = _f64(bou)
        ic = None if int_ is None else _f64(int_)
        self.ctx.check(load().b200_sys_set_interface_coeffs(self.h, r, i, _dp(b), _dp(ic)))

    def set_interface_attached(self, r: int, i: int, attached: bool):
        """regionInterfaceType::attach()/detach() (regionInterfaceType.C:543-627): while a regionCouple interface is
        detached, amul / solve raise (monolithicCouplingFvPatchField.C:406-413 is fatal for a detached patch)."""
        self.ctx.check(load().b200_sys_set_interface_attached(self.h, r, i, int(bool(attached))))

    def set_interface_ggi(self, r: int, i: int, nPeerFaces: int, offsets=None, addr=None, weights=None):
        """Replace the GGI addressing / weights of a regionCouple interface (re-computed by attach() after mesh motion);
        offsets None: identity pairing.  The device interface tables are rebuilt at the next use."""
        if offsets is None:
            self.ctx.check(load().b200_sys_set_interface_ggi(self.h, r, i, int(nPeerFaces), None, None, None))
        else:
            o, a, w = _i32(offsets), _i32(addr), _f64(weights)
            self.ctx.check(load().b200_sys_set_interface_ggi(self.h, r, i, int(nPeerFaces), _ip(o), _ip(a), _dp(w)))

    # ---- device-side coefficient refresh of the T equations
    def set_fv_geometry(self, r: int, V, magSf, deltaCoeffs, bCells=None, bIntCoeffs=None, bSrcCoeffs=None):
        """Static FV tables of region r: mesh.V(), magSf / deltaCoeffs of the internal faces and the boundary faces in
        patch order with their internalCoeffs / boundary-source contributions (addBoundaryDiag / addBoundarySource)."""
        v, a, d = _f64(V), _f64(magSf), _f64(deltaCoeffs)
        nB = 0 if bCells is None else int(np.asarray(bCells).size)
        bc = _i32(bCells) if nB else None
        bi = _f64(bIntCoeffs) if nB else None
        bs = _f64(bSrcCoeffs) if nB else None
        self.ctx.check(load().b200_sys_set_fv_geometry(self.h, r, _dp(v), _dp(a), _dp(d), nB, _ip(bc), _dp(bi), _dp(bs)))

    def assemble_T(self, r: int, form: int, rhoC: float, rDeltaT: float, kappa: float, kappaFace=None, phi=None):
        """conductTemperature / transportTemperature::setCoupledEqns on the device (conductTemperature.C:135-142,
        transportTemperature.C:129-140): coefficients of region r and its part of the resident b from the resident x
        (= T.oldTime()).  kappaFace / phi None: keep the resident table."""
        kf = None if kappaFace is None else _f64(kappaFace)
        ph = None if phi is None else _f64(phi)
        self.ctx.check(load().b200_sys_assemble_T(self.h, r, int(form), float(rhoC), float(rDeltaT), float(kappa), _dp(kf), _dp(ph)))

    def set_all_coeffs(self):
        for r, reg in enumerate(self.rs.regions):
            self.set_coeffs(r, reg.diag, reg.upper, reg.lower)
            for i, itf in enumerate(reg.interfaces):
                self.set_interface_coeffs(r, i, itf.bouCoeffs, itf.intCoeffs)

    @property
    def nCells(self) -> int:
        return int(load().b200_sys_num_cells(self.h))

    @property
    def nFaces(self) -> int:
        return int(load().b200_sys_num_faces(self.h))

    # ---- helpers
    def split(self, vec: np.ndarray) -> List[np.ndarray]:
        out, o = [], 0
        for n in self.sizes:
            out.append(np.ascontiguousarray(vec[o:o + n], dtype=np.float64))
            o += n
        return out

    def _empty(self) -> List[np.ndarray]:
        return [np.empty(n) for n in self.sizes]

    # ---- single operations (test hooks)
    def amul(self, x: np.ndarray, transpose: bool = False) -> np.ndarray:
        xs, ys = self.split(x), self._empty()
        self.ctx.check(load().b200_amul(self.h, _dpp(xs), _dpp(ys), int(transpose)))
        return np.concatenate(ys) if ys else np.empty(0)

    def residual(self, x: np.ndarray, b: np.ndarray) -> np.ndarray:
        """lduMatrix::residual: b - A x with the reference's per-row rounding (b200_residual)."""
        xs, bs, rs_ = self.split(x), self.split(b), self._empty()
        self.ctx.check(load().b200_residual(self.h, _dpp(xs), _dpp(bs), _dpp(rs_)))
        return np.concatenate(rs_) if rs_ else np.empty(0)

    def sumA(self) -> np.ndarray:
        """lduMatrix::sumA (b200_sum_a)."""
        out = self._empty()
        self.ctx.check(load().b200_sum_a(self.h, _dpp(out)))
        return np.concatenate(out) if out else np.empty(0)

    def precondition(self, precond: int, r: np.ndarray, transpose: bool = False) -> np.ndarray:
        rs_, ws = self.split(r), self._empty()
        self.ctx.check(load().b200_precondition(self.h, precond, _dpp(rs_), _dpp(ws), int(transpose)))
        return np.concatenate(ws) if ws else np.empty(0)

    def smooth(self, precond: int, x: np.ndarray, b: np.ndarray, nSweeps: int = 1) -> np.ndarray:
        """DICSmoother / DILUSmoother::smooth: nSweeps times  psi += M^-1 (source - A psi)  with the preconditioner's sweeps."""
        xs, bs = [np.array(v, copy=True) for v in self.split(x)], self.split(b)
        self.ctx.check(load().b200_smooth(self.h, precond, int(nSweeps), _dpp(xs), _dpp(bs)))
        return np.concatenate(xs) if xs else np.empty(0)

    def rD(self, precond: int) -> np.ndarray:
        out = self._empty()
        self.ctx.check(load().b200_get_rD(self.h, precond, _dpp(out)))
        return np.concatenate(out) if out else np.empty(0)

    def reduce(self, a: np.ndarray, b: np.ndarray):
        as_, bs = self.split(a), self.split(b)
        out = np.zeros(2)
        self.ctx.check(load().b200_reduce(self.h, _dpp(as_), _dpp(bs), _dp(out)))
        return float(out[0]), float(out[1])

    # ---- solve
    @staticmethod
    def _opts(solver, precond, tolerance, relTol, minIter, maxIter) -> SolverOpts:
        return SolverOpts(solver, precond, tolerance, relTol, minIter, maxIter)

    @staticmethod
    def _info(perf: Perf, hist: Optional[np.ndarray]):
        d = dict(initialResidual=perf.initialResidual, finalResidual=perf.finalResidual, nIterations=perf.nIterations,
                 converged=bool(perf.converged), singular=bool(perf.singular), normFactor=perf.normFactor,
                 deviceMs=perf.deviceMs)
        if hist is not None:
            d["history"] = hist[:min(hist.size, perf.nIterations + 1)].copy()
        return d

    def solve(self, x0: np.ndarray, b: np.ndarray, solver=SOLVER_BICGSTAB, precond=PRECOND_DILU, tolerance=1e-6,
              relTol=0.0, minIter=0, maxIter=1000, history: bool = True):
        """b200_solve: host x (in/out) and b per region; H2D and D2H inside the call."""
        xs, bs = [a.copy() for a in self.split(x0)], self.split(b)
        opts, perf = self._opts(solver, precond, tolerance, relTol, minIter, maxIter), Perf()
        hist = np.full(maxIter + 1, np.nan) if history else None
        self.ctx.check(load().b200_solve(self.h, C.byref(opts), _dpp(xs), _dpp(bs), C.byref(perf), _dp(hist),
                                         hist.size if history else 0))
        return (np.concatenate(xs) if xs else np.empty(0)), self._info(perf, hist)

    def upload(self, x0: Optional[np.ndarray], b: Optional[np.ndarray]):
        xs = None if x0 is None else self.split(x0)
        bs = None if b is None else self.split(b)
        self.ctx.check(load().b200_upload(self.h, None if xs is None else _dpp(xs), None if bs is None else _dpp(bs)))

    def solve_resident(self, solver=SOLVER_BICGSTAB, precond=PRECOND_DILU, tolerance=1e-6, relTol=0.0, minIter=0,
                       maxIter=1000, history: bool = False):
        opts, perf = self._opts(solver, precond, tolerance, relTol, minIter, maxIter), Perf()
        hist = np.full(maxIter + 1, np.nan) if history else None
        self.ctx.check(load().b200_solve_resident(self.h, C.byref(opts), C.byref(perf), _dp(hist), hist.size if history else 0))
        return self._info(perf, hist)

    def download(self) -> np.ndarray:
        xs = self._empty()
        self.ctx.check(load().b200_download(self.h, _dpp(xs)))
        return np.concatenate(xs) if xs else np.empty(0)

    def x_save(self):
        self.ctx.check(load().b200_x_save(self.h))

    def x_restore(self):
        self.ctx.check(load().b200_x_restore(self.h))

    # ---- profiling
    def set_profiling(self, enable: bool):
        self.ctx.check(load().b200_set_profiling(self.h, int(enable)))

    def kernel_times(self, reset: bool = True):
        ms = np.zeros(len(KERNEL_CLASSES))
        n = np.zeros(len(KERNEL_CLASSES), dtype=np.int64)
        self.ctx.check(load().b200_get_kernel_times(self.h, _dp(ms), n.ctypes.data_as(C.POINTER(C.c_int64)), int(reset)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_CLASSES)}

    def sweep_stats(self, direction: int, enable: bool = True) -> np.ndarray:
        """Debug counters of the sweep kernel per group, 16 columns: consumer {cycles, wait cycles, start ns, end ns},
        producer polls, nT, general blocks, blocks, producer 0 {cycles, stage-wait, value-wait, spin cycles}, -; then per
        block 8 clock64 stamps: consumer {ready seen, done}, loader issue, producer 0 {step start, next stage landed,
        values checked, delivered}, producer 7 delivered;
        returns what the sweeps since the previous call recorded and (re)arms."""
        n = self.ctx.check(load().b200_debug_sweep_stats(self.h, direction, 0, None, 0))
        out = np.zeros((max(n, 1), SWEEP_STATS_STRIDE), dtype=np.int64)
        self.ctx.check(load().b200_debug_sweep_stats(self.h, direction, int(enable), out.ctypes.data_as(C.POINTER(C.c_longlong)), out.size))
        return out[:n]

    def close(self):
        if self.h:
            load().b200_sys_destroy(self.h)
            self.h = C.c_void_p()
