"""Block-coupled (vector4) systems: ctypes binding of include/b200_blk.h and the host-side mirror of the
reference's ``fvBlockMatrix<vector4>`` interface for this path.

Reference call chain (``/root/reference``):
    pUCoupledIcoFluid::setCoupledEqns            src/regions/pUCoupledIcoFluid/pUCoupledIcoFluid.C:584-621
        fvBlockMatrix<vector4> UpEqn(Up)
        UpEqn.insertEquation(0, UEqn)            filesToReplace/fvBlockMatrix.C:768-795 (:60-160 diag/source, :163-288 upper/lower)
        UpEqn.insertEquation(3, pEqn)
        UpEqn.insertBlockCoupling(0, 3, pInU, true) / (3, 0, UInp, false)    :841-872 -> insertBlock :379-474
    multiRegionSystem::assembleAndSolveEqns      src/multiRegionSystem/multiRegionSystem.C:293
        fvBlockMatrix<vector4>::solve(dict)      filesToReplace/fvBlockMatrix.C:1360-1388
            BlockLduSolver<vector4>::New(psi.name(), *this, dict)->solve(psi, source)

There is no CPU fallback: everything numeric goes through libb200ldu.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import ldu
from .solvers import FatalError, read_controls

SCALAR, LINEAR, SQUARE = 1, 4, 16
SOLVER_CG, SOLVER_BICGSTAB = 0, 1
BLK_KERNEL_CLASSES = ["amul", "sweep_fwd", "sweep_bwd", "vector", "precon_setup"]

# every symbol include/b200_blk.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "b200_blk_create", "b200_blk_destroy", "b200_blk_set_coeffs", "b200_blk_amul", "b200_blk_precondition",
    "b200_blk_get_precon_diag", "b200_blk_reduce", "b200_blk_solve", "b200_blk_upload", "b200_blk_solve_resident",
    "b200_blk_download", "b200_blk_x_save", "b200_blk_x_restore", "b200_blk_set_profiling", "b200_blk_get_kernel_times",
    "b200_blk_add_interface", "b200_blk_set_interface_coeffs",
]

# BlockLduSolver / BlockLduPrecon run-time selection names: foam-extend's own and the cuda* names of the drop-in
BLOCK_SOLVER_TABLE: Dict[str, tuple] = {
    "cudaBlockBiCGStab": (SOLVER_BICGSTAB, False), "BiCGStab": (SOLVER_BICGSTAB, False),
    "cudaBlockCG": (SOLVER_CG, True), "CG": (SOLVER_CG, True),
}
BLOCK_PRECOND_TABLE: Dict[str, int] = {
    "cudaBlockCholesky": ldu.PRECOND_CHOLESKY, "Cholesky": ldu.PRECOND_CHOLESKY,
    "diagonal": ldu.PRECOND_DIAGONAL, "none": ldu.PRECOND_NONE,
}


class BlkPerf(C.Structure):
    _fields_ = [("initialResidual", C.c_double * 4), ("finalResidual", C.c_double * 4), ("nIterations", C.c_int),
                ("converged", C.c_int), ("singular", C.c_int), ("normFactor", C.c_double), ("deviceMs", C.c_double)]


_bound = False


def _lib():
    global _bound
    L = ldu.load()
    if not _bound:
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
        L.b200_blk_create.argtypes = [vp, C.c_int32, C.c_int32, ip, ip, C.POINTER(vp)]
        L.b200_blk_destroy.argtypes = [vp]
        L.b200_blk_set_coeffs.argtypes = [vp, C.c_int, dp, C.c_int, dp, C.c_int, dp]
        L.b200_blk_add_interface.argtypes = [vp, C.c_int32, ip, C.c_int32, C.c_int32, ip]
        L.b200_blk_set_interface_coeffs.argtypes = [vp, C.c_int32, C.c_int, dp]
        L.b200_blk_amul.argtypes = [vp, dp, dp]
        L.b200_blk_precondition.argtypes = [vp, C.c_int, dp, dp]
        L.b200_blk_get_precon_diag.argtypes = [vp, C.c_int, dp, C.POINTER(C.c_int)]
        L.b200_blk_reduce.argtypes = [vp, dp, dp, dp]
        L.b200_blk_solve.argtypes = [vp, C.POINTER(ldu.SolverOpts), dp, dp, C.POINTER(BlkPerf), dp, C.c_int]
        L.b200_blk_upload.argtypes = [vp, dp, dp]
        L.b200_blk_solve_resident.argtypes = [vp, C.POINTER(ldu.SolverOpts), C.POINTER(BlkPerf), dp, C.c_int]
        L.b200_blk_download.argtypes = [vp, dp]
        L.b200_blk_x_save.argtypes = [vp]
        L.b200_blk_x_restore.argtypes = [vp]
        L.b200_blk_set_profiling.argtypes = [vp, C.c_int]
        L.b200_blk_get_kernel_times.argtypes = [vp, dp, C.POINTER(C.c_int64), C.c_int]
        _bound = True
    return L


def coeff_kind(a: np.ndarray) -> int:
    """Active type of a coefficient array: [n] SCALAR, [n,4] LINEAR, [n,4,4] / [n,16] SQUARE."""
    k = 1 if a.ndim == 1 else int(np.prod(a.shape[1:]))
    if k not in (SCALAR, LINEAR, SQUARE):
        raise FatalError(f"coefficient array of shape {a.shape} is not scalar / linear / square for vector4")
    return k


def _perf_dict(p: BlkPerf, hist: Optional[np.ndarray]) -> dict:
    d = dict(initialResidual=np.array(p.initialResidual[:]), finalResidual=np.array(p.finalResidual[:]),
             nIterations=p.nIterations, converged=bool(p.converged), singular=bool(p.singular),
             normFactor=p.normFactor, deviceMs=p.deviceMs)
    if hist is not None:
        d["history"] = hist[: p.nIterations + 1].copy()
    return d


class BlockSystem:
    """One BlockLduMatrix<vector4> on the device (b200_blk)."""

    def __init__(self, ctx: ldu.Context, lowerAddr, upperAddr, nCells: int):
        L = _lib()
        self.ctx = ctx
        self.nCells = int(nCells)
        l, u = ldu._i32(lowerAddr), ldu._i32(upperAddr)
        self.nFaces = int(l.size)
        self.h = C.c_void_p()
        ctx.check(L.b200_blk_create(ctx.h, self.nCells, self.nFaces, ldu._ip(l), ldu._ip(u), C.byref(self.h)))

    def close(self):
        if self.h:
            _lib().b200_blk_destroy(self.h)
            self.h = C.c_void_p()

    def set_coeffs(self, diag, upper, lower=None):
        d, u = ldu._f64(diag), ldu._f64(upper)
        lo = None if lower is None else ldu._f64(lower)
        self.ctx.check(_lib().b200_blk_set_coeffs(self.h, coeff_kind(d), ldu._dp(d), coeff_kind(u), ldu._dp(u),
                                                  0 if lo is None else coeff_kind(lo), ldu._dp(lo)))

    def add_interface(self, faceCells, peerRank: int, peerIface: int) -> int:
        """A processor patch of the block matrix (BlockLduMatrix::interfaces()); peerIface: the neighbour's index of
        the matching patch.  -> the index of this patch."""
        fc = ldu._i32(faceCells)
        idx = C.c_int32(-1)
        self.ctx.check(_lib().b200_blk_add_interface(self.h, int(fc.size), ldu._ip(fc), int(peerRank), int(peerIface), C.byref(idx)))
        return idx.value

    def set_interface_coeffs(self, iface: int, coupleUpper):
        cu = ldu._f64(coupleUpper)
        self.ctx.check(_lib().b200_blk_set_interface_coeffs(self.h, int(iface), coeff_kind(cu), ldu._dp(cu)))

    def _field(self, a) -> np.ndarray:
        a = ldu._f64(a)
        if a.size != 4 * self.nCells:
            raise ValueError(f"field of {a.size} doubles for {self.nCells} vector4 cells")
        return a.reshape(self.nCells, 4)

    def amul(self, x) -> np.ndarray:
        x = self._field(x)
        y = np.empty_like(x)
        self.ctx.check(_lib().b200_blk_amul(self.h, ldu._dp(x), ldu._dp(y)))
        return y

    def precondition(self, precond: int, r) -> np.ndarray:
        r = self._field(r)
        w = np.empty_like(r)
        self.ctx.check(_lib().b200_blk_precondition(self.h, precond, ldu._dp(r), ldu._dp(w)))
        return w

    def precon_diag(self, precond: int) -> np.ndarray:
        out = np.empty(16 * max(self.nCells, 1))
        k = C.c_int(0)
        self.ctx.check(_lib().b200_blk_get_precon_diag(self.h, precond, ldu._dp(out), C.byref(k)))
        return out[: self.nCells * k.value].reshape(self.nCells, k.value).copy()

    def reduce(self, a, b):
        a, b = self._field(a), self._field(b)
        out = np.empty(5)
        self.ctx.check(_lib().b200_blk_reduce(self.h, ldu._dp(a), ldu._dp(b), ldu._dp(out)))
        return out[0], out[1:].copy()

    def _opts(self, solver, precond, tolerance, relTol, minIter, maxIter):
        return ldu.SolverOpts(solver, precond, tolerance, relTol, minIter, maxIter)

    def solve(self, x0, b, solver=SOLVER_BICGSTAB, precond=ldu.PRECOND_CHOLESKY, tolerance=1e-6, relTol=0.0, minIter=0,
              maxIter=1000, history=True):
        x = np.array(self._field(x0), copy=True)
        b = self._field(b)
        o = self._opts(solver, precond, tolerance, relTol, minIter, maxIter)
        p = BlkPerf()
        cap = maxIter + 2 if history else 0
        hist = np.empty((cap, 4)) if history else None
        self.ctx.check(_lib().b200_blk_solve(self.h, C.byref(o), ldu._dp(x), ldu._dp(b), C.byref(p), ldu._dp(hist), cap))
        return x, _perf_dict(p, hist)

    # device-resident variant (bench)
    def upload(self, x0, b):
        x0, b = self._field(x0), self._field(b)
        self.ctx.check(_lib().b200_blk_upload(self.h, ldu._dp(x0), ldu._dp(b)))

    def x_save(self):
        self.ctx.check(_lib().b200_blk_x_save(self.h))

    def x_restore(self):
        self.ctx.check(_lib().b200_blk_x_restore(self.h))

    def solve_resident(self, solver=SOLVER_BICGSTAB, precond=ldu.PRECOND_CHOLESKY, tolerance=1e-6, relTol=0.0, minIter=0,
                       maxIter=1000):
        o = self._opts(solver, precond, tolerance, relTol, minIter, maxIter)
        p = BlkPerf()
        self.ctx.check(_lib().b200_blk_solve_resident(self.h, C.byref(o), C.byref(p), None, 0))
        return _perf_dict(p, None)

    def download(self) -> np.ndarray:
        x = np.empty((self.nCells, 4))
        self.ctx.check(_lib().b200_blk_download(self.h, ldu._dp(x)))
        return x

    def set_profiling(self, on: bool):
        self.ctx.check(_lib().b200_blk_set_profiling(self.h, int(on)))

    def kernel_times(self, reset=False) -> dict:
        ms = np.zeros(5)
        n = np.zeros(5, np.int64)
        self.ctx.check(_lib().b200_blk_get_kernel_times(self.h, ldu._dp(ms), n.ctypes.data_as(C.POINTER(C.c_int64)), int(reset)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(BLK_KERNEL_CLASSES)}


# ------------------------------------------------------------------------------------------ fvBlockMatrix mirror

@dataclass
class ScalarEqn:
    """What insertEquation takes from an ``fvMatrix<scalar>`` / one direction of an ``fvMatrix<vector>`` after
    ``completeAssembly`` + ``addBoundaryDiag`` / ``addBoundarySource`` (fvBlockMatrix.C:60-160): plain LDU arrays."""
    diag: np.ndarray
    source: np.ndarray
    upper: Optional[np.ndarray] = None   # None: diagonal-only matrix (insertUpperLower returns early, :172-176)
    lower: Optional[np.ndarray] = None   # None: symmetric


@dataclass
class BlockCoupling:
    """A ``BlockLduSystem<vector, scalar>`` (e.g. fvm::grad(p), fvm::UDiv(U)): LINEAR (3-component) diag / upper / lower
    and a scalar or vector source (fvBlockMatrix.C:379-474, 477-568)."""
    diag: np.ndarray    # [N, 3]
    upper: np.ndarray   # [F, 3]
    lower: np.ndarray   # [F, 3]
    source: Optional[np.ndarray] = None  # [N] or [N, k]


@dataclass
class BlockSolverPerformance:
    solverName: str
    fieldName: str
    initialResidual: np.ndarray
    finalResidual: np.ndarray
    nIterations: int = 0
    converged: bool = False
    singular: bool = False
    normFactor: float = 0.0
    deviceMs: float = 0.0
    history: Optional[np.ndarray] = None

    def line(self) -> str:
        # BlockSolverPerformance<Type>::print
        v = lambda a: "(" + " ".join(f"{x:g}" for x in a) + ")"
        return (f"{self.solverName}:  Solving for {self.fieldName}, Initial residual = {v(self.initialResidual)}, "
                f"Final residual = {v(self.finalResidual)}, No Iterations {self.nIterations}")


class fvBlockMatrix:
    """``fvBlockMatrix<vector4>``: the block matrix of one region, assembled on the host from scalar equations and
    coupling blocks exactly as the reference does, solved on the device.

    Coefficient arrays start UNALLOCATED and take the narrowest active type that holds what was inserted
    (SCALAR -> LINEAR on the second equation, SQUARE once a cross-coupling block arrives, fvBlockMatrix.C:184-221,
    441-446)."""

    nComp = 4

    def __init__(self, lowerAddr, upperAddr, nCells: int, psi: Optional[np.ndarray] = None, name: str = "Up"):
        self.l = np.ascontiguousarray(lowerAddr, np.int32)
        self.u = np.ascontiguousarray(upperAddr, np.int32)
        self.nCells, self.nFaces, self.name = int(nCells), int(self.l.size), name
        self.psi = np.zeros((self.nCells, 4)) if psi is None else np.array(psi, np.float64).reshape(self.nCells, 4)
        self.source = np.zeros((self.nCells, 4))
        self.diag: Optional[np.ndarray] = None
        self.upper: Optional[np.ndarray] = None
        self.lower: Optional[np.ndarray] = None

    # ---- active-type handling of CoeffField<vector4>
    @staticmethod
    def _as_linear(a: Optional[np.ndarray], n: int) -> np.ndarray:
        if a is None:
            return np.zeros((n, 4))
        if a.ndim == 1:
            return np.repeat(a[:, None], 4, axis=1)
        if a.ndim == 2:
            return a
        raise FatalError("cannot demote a square coefficient to linear")

    @staticmethod
    def _as_square(a: Optional[np.ndarray], n: int) -> np.ndarray:
        if a is None:
            return np.zeros((n, 4, 4))
        if a.ndim == 3:
            return a
        lin = fvBlockMatrix._as_linear(a, n)
        sq = np.zeros((n, 4, 4))
        for i in range(4):
            sq[:, i, i] = lin[:, i]
        return sq

    def symmetric(self) -> bool:
        return self.lower is None

    def insertEquation(self, dir: int, eqn: ScalarEqn, nCmpts: int = 1):
        """fvBlockMatrix.C:768-795 for a scalar equation (nCmpts = 1) or a vector equation whose components share the
        LDU coefficients (nCmpts = 3; ``source`` then [N, 3])."""
        if dir + nCmpts > 4:
            raise FatalError("insertEquation: direction out of range for vector4")
        src = np.asarray(eqn.source, np.float64).reshape(self.nCells, -1)
        # insertDiagSource (:60-160): diag is linear unless already square
        if self.diag is not None and self.diag.ndim == 3:
            for c in range(nCmpts):
                self.diag[:, dir + c, dir + c] = eqn.diag
        else:
            self.diag = self._as_linear(self.diag, self.nCells).copy()
            for c in range(nCmpts):
                self.diag[:, dir + c] = eqn.diag
        for c in range(nCmpts):
            self.source[:, dir + c] += src[:, c if src.shape[1] > 1 else 0]
        # insertUpperLower (:163-288)
        if eqn.upper is None:
            return
        self._insert_ul("upper", np.asarray(eqn.upper, np.float64), dir, nCmpts)
        if eqn.lower is None and self.lower is None:
            return  # "Both matrices are symmetric: inserting only upper triangle"
        if self.lower is None:
            # BlockLduMatrix::lower() allocates the lower triangle as the transposed upper one on first access
            self.lower = self._transposed(self.upper)
        # a symmetric fvMatrix returns its upper coefficients from lower()
        self._insert_ul("lower", np.asarray(eqn.lower if eqn.lower is not None else eqn.upper, np.float64), dir, nCmpts)

    @staticmethod
    def _transposed(a: np.ndarray) -> np.ndarray:
        return a.transpose(0, 2, 1).copy() if a.ndim == 3 else a.copy()

    def _insert_ul(self, nm: str, coef: np.ndarray, dir: int, nCmpts: int):
        cur = getattr(self, nm)
        if cur is None:
            setattr(self, nm, coef.copy())  # UNALLOCATED: asScalar() = coefficients, whatever the direction (:181-184)
        elif cur.ndim == 3:
            for c in range(nCmpts):
                cur[:, dir + c, dir + c] = coef
        else:
            lin = self._as_linear(cur, self.nFaces).copy()
            for c in range(nCmpts):
                lin[:, dir + c] = coef
            setattr(self, nm, lin)

    def insertBlockCoupling(self, dirI: int, dirJ: int, blk: BlockCoupling, incFirst: bool):
        """fvBlockMatrix.C:841-872 -> insertBlock (:379-474) + the source part of insertBoundaryContributions
        (:477-520): cross-coupling forces the SQUARE active type."""
        if dirI == dirJ:
            raise FatalError("Trying to insert coupling in the position where equation should be, since dirI = dirJ. "
                             "Try using insertEquation member function.")
        bd, bu, bl = (np.asarray(a, np.float64).reshape(-1, 3) for a in (blk.diag, blk.upper, blk.lower))
        if self.lower is None and self.upper is not None:
            self.lower = self._transposed(self.upper)
        self.diag = self._as_square(self.diag, self.nCells).copy()
        self.upper = self._as_square(self.upper, self.nFaces).copy()
        self.lower = self._as_square(self.lower, self.nFaces).copy()
        i, j = dirI, dirJ
        for c in range(3):
            self.diag[:, i, j] += bd[:, c]
            self.upper[:, i, j] += bu[:, c]
            self.lower[:, i, j] += bl[:, c]
            if incFirst:
                i += 1
            else:
                j += 1
        if blk.source is not None:
            s = np.asarray(blk.source, np.float64).reshape(self.nCells, -1)
            i = dirI
            for c in range(s.shape[1]):
                self.source[:, i] += s[:, c]
                if incFirst:
                    i += 1

    # ---- the solve (fvBlockMatrix.C:1360-1388)
    def device_system(self, ctx: ldu.Context) -> BlockSystem:
        if self.diag is None or self.upper is None:
            raise FatalError("fvBlockMatrix::solve: no equation inserted")
        S = BlockSystem(ctx, self.l, self.u, self.nCells)
        up, lo = self.upper, self.lower
        if lo is not None and lo.ndim != up.ndim:
            # the device wants one active type for both triangles ("Assuming lower and upper triangle have the same
            # active type", BlockCholeskyPrecon): widen the narrower one
            wide = self._as_square if max(lo.ndim, up.ndim) == 3 else self._as_linear
            up, lo = wide(up, self.nFaces), wide(lo, self.nFaces)
        S.set_coeffs(self.diag, up, lo)
        return S

    def solve(self, ctx: ldu.Context, solverDict: dict, system: Optional[BlockSystem] = None) -> BlockSolverPerformance:
        c = read_controls(solverDict)
        name = c["solver"]
        if name is None:
            raise FatalError("solver dictionary lacks the 'solver' entry")
        if name not in BLOCK_SOLVER_TABLE:
            raise FatalError(f"Unknown matrix solver {name}; valid solvers are {sorted(BLOCK_SOLVER_TABLE)}")
        if c["preconditioner"] not in BLOCK_PRECOND_TABLE:
            raise FatalError(f"Unknown matrix preconditioner {c['preconditioner']}; valid preconditioners are "
                             f"{sorted(BLOCK_PRECOND_TABLE)}")
        solverId, symOnly = BLOCK_SOLVER_TABLE[name]
        if symOnly and not self.symmetric():
            raise FatalError(f"Unknown asymmetric matrix solver {name}")
        own = system is None
        S = system or self.device_system(ctx)
        try:
            x, info = S.solve(self.psi, self.source, solverId, BLOCK_PRECOND_TABLE[c["preconditioner"]], c["tolerance"],
                              c["relTol"], c["minIter"], c["maxIter"])
        finally:
            if own:
                S.close()
        self.psi[...] = x
        return BlockSolverPerformance(name, self.name, info["initialResidual"], info["finalResidual"], info["nIterations"],
                                      info["converged"], info["singular"], info["normFactor"], info["deviceMs"],
                                      info.get("history"))

    def retrieveSolution(self, dir: int, nCmpts: int = 1) -> np.ndarray:
        """fvBlockMatrix.C:739-765: components dir .. dir+nCmpts-1 of the block solution."""
        return self.psi[:, dir] .copy() if nCmpts == 1 else self.psi[:, dir:dir + nCmpts].copy()
