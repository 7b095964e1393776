"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the decomposed block-coupled (vector4) solve on the CPU.

A BlockLduMatrix<vector4> split over subdomains with processor patches, as foam-extend runs fvBlockMatrix<vector4>::solve
(/root/reference/filesToReplace/fvBlockMatrix.C:1360-1388) in parallel:
  * Amul = the subdomain's own product (oracle/blk_oracle.c, blk_amul) followed by
    BlockLduMatrix::updateInterfaces(coupleUpper, Ax, x): per processor patch, in patch order,
    Ax[faceCells[f]] -= coupleUpper[f] * x_neighbour[f]   (processorFvPatchField<Type>::updateInterfaceMatrix for block
    matrices, switchToLhs = false);
  * BlockCholeskyPrecon / BlockDiagonalPrecon use the subdomain's own coefficients only (no interface term);
  * gSumProd / gSum(cmptMag) / gAverage: per subdomain sequential sums, added in subdomain order.
The Krylov loops restate blk_oracle.c's blk_solve_bicgstab / blk_solve_cg (BlockBiCGStabSolver / BlockCGSolver) with
these three global operations.  Parity unpinned, like the rest of the oracle (no foam-extend build here).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from . import pyblk

SMALL, VSMALL, GREAT = 1.0e-15, 1.0e-300, 1.0e15


def _mult(coef: np.ndarray, x: np.ndarray) -> np.ndarray:
    """BlockCoeff<vector4> (scalar [F], linear [F,4] or square [F,4,4]) times x [F,4], component sums left to right."""
    if coef.ndim == 1:
        return coef[:, None] * x
    if coef.ndim == 2:
        return coef * x
    out = coef[:, :, 0] * x[:, 0:1]
    for j in range(1, 4):
        out = out + coef[:, :, j] * x[:, j:j + 1]
    return out


def _seqsum(a: np.ndarray) -> float:
    s = 0.0
    for v in a.tolist():
        s += v
    return s


class MultiBlockOracle:
    """subs: list of dicts  n, l, u, diag, upper, lower (None = symmetric),
    ifaces: list of dicts  faceCells, peer (subdomain index), peerIface, coupleUpper."""

    def __init__(self, subs: Sequence[Dict]):
        self.subs = list(subs)
        self.loc = [pyblk.BlockOracle(s["l"], s["u"], s["n"], s["diag"], s["upper"], s.get("lower")) for s in self.subs]
        self.nGlobal = sum(int(s["n"]) for s in self.subs)

    def close(self):
        for o in self.loc:
            o.close()

    # ---- the three global operations
    def amul(self, xs: List[np.ndarray]) -> List[np.ndarray]:
        ys = [o.amul(x) if o.n else np.zeros((0, 4)) for o, x in zip(self.loc, xs)]
        for d, s in enumerate(self.subs):
            for I in s.get("ifaces", []):
                nb = self.subs[I["peer"]]["ifaces"][I["peerIface"]]
                xn = xs[I["peer"]][np.asarray(nb["faceCells"], np.int64)]
                contrib = _mult(np.asarray(I["coupleUpper"], np.float64), xn)
                fc = np.asarray(I["faceCells"], np.int64)
                for f in range(fc.size):   # patch order; a cell may own several faces of the patch
                    ys[d][fc[f]] -= contrib[f]
        return ys

    def sumprod(self, a, b) -> float:
        tot = 0.0
        for o, x, y in zip(self.loc, a, b):
            tot += o.sumprod(x, y) if o.n else 0.0
        return tot

    def sum_cmptmag(self, a) -> np.ndarray:
        tot = np.zeros(4)
        for x in a:
            tot += np.array([_seqsum(np.abs(x[:, i])) for i in range(4)])
        return tot

    def norm_factor(self, xs, bs) -> float:
        tot = np.zeros(4)
        for x in xs:
            tot += np.array([_seqsum(x[:, i]) for i in range(4)])
        xRef = tot / (self.nGlobal if self.nGlobal else 1)
        wA = self.amul(xs)
        pA = self.amul([np.tile(xRef, (x.shape[0], 1)) for x in xs])
        s = 0.0
        for w, p, b in zip(wA, pA, bs):
            d1, d2 = w - p, b - p

            def mag(d):
                q = d[:, 0] * d[:, 0]
                for i in range(1, 4):
                    q = q + d[:, i] * d[:, i]
                return np.sqrt(q)
            s += _seqsum(mag(d1) + mag(d2))
        return s + SMALL

    def precondition(self, rs, precond):
        return [o.precondition(r, precond) if o.n else np.zeros((0, 4)) for o, r in zip(self.loc, rs)]

    # ---- BlockBiCGStabSolver / BlockCGSolver
    def solve(self, xs0, bs, solver="BiCGStab", precond="Cholesky", tolerance=1e-6, relTol=0.0, minIter=0, maxIter=1000):
        xs = [np.array(x, np.float64).reshape(-1, 4).copy() for x in xs0]
        bs = [np.asarray(b, np.float64).reshape(-1, 4) for b in bs]
        nf = self.norm_factor(xs, bs)
        perf = dict(normFactor=nf, nIterations=0, converged=False, singular=False)
        rs = [b - y for b, y in zip(bs, self.amul(xs))]
        perf["initialResidual"] = self.sum_cmptmag(rs) / nf
        perf["finalResidual"] = perf["initialResidual"].copy()
        hist = [perf["initialResidual"].copy()]

        def stop():
            if perf["nIterations"] < minIter:
                return False
            fin, ini = perf["finalResidual"].max(), perf["initialResidual"].max()
            perf["converged"] = bool(fin < tolerance or (relTol > 1e-15 and fin <= relTol * ini))
            return perf["nIterations"] >= maxIter or perf["converged"]

        if not stop():
            zeros = lambda: [np.zeros_like(x) for x in xs]
            if solver == "BiCGStab":
                rho, alpha, omega = GREAT, 0.0, GREAT
                p, v = zeros(), zeros()
                rw = [r.copy() for r in rs]
                while True:
                    rhoOld = rho
                    rho = self.sumprod(rw, rs)
                    beta = rho / rhoOld * (alpha / omega)
                    if rho == 0:
                        rw = [r.copy() for r in rs]
                        rho = self.sumprod(rw, rs)
                        alpha = omega = beta = 0.0
                    p = [r + beta * pp - beta * omega * vv for r, pp, vv in zip(rs, p, v)]
                    ph = self.precondition(p, precond)
                    v = self.amul(ph)
                    alpha = rho / self.sumprod(rw, v)
                    sv = [r - alpha * vv for r, vv in zip(rs, v)]
                    sh = self.precondition(sv, precond)
                    t = self.amul(sh)
                    omega = self.sumprod(t, sv) / self.sumprod(t, t)
                    xs = [x + alpha * a + omega * c for x, a, c in zip(xs, ph, sh)]
                    rs = [s_ - omega * tt for s_, tt in zip(sv, t)]
                    perf["finalResidual"] = self.sum_cmptmag(rs) / nf
                    perf["nIterations"] += 1
                    hist.append(perf["finalResidual"].copy())
                    if stop():
                        break
            else:
                rho = GREAT
                p = zeros()
                while True:
                    rhoOld = rho
                    w = self.precondition(rs, precond)
                    rho = self.sumprod(w, rs)
                    beta = rho / rhoOld
                    p = [ww + beta * pp for ww, pp in zip(w, p)]
                    w = self.amul(p)
                    wApA = self.sumprod(w, p)
                    if not (abs(wApA) / nf > VSMALL):
                        perf["singular"] = True
                        break
                    alpha = rho / wApA
                    xs = [x + alpha * pp for x, pp in zip(xs, p)]
                    rs = [r - alpha * ww for r, ww in zip(rs, w)]
                    perf["finalResidual"] = self.sum_cmptmag(rs) / nf
                    perf["nIterations"] += 1
                    hist.append(perf["finalResidual"].copy())
                    if stop():
                        break
        perf["history"] = np.array(hist)
        return xs, perf


def split_block_system(n, l, u, diag, upper, lower, owner: np.ndarray):
    """Decompose one block system into subdomains (owner[cell] = subdomain) the way decomposePar does for the matrix:
    cells keep their relative order; a face between two subdomains becomes one face of a processor patch on either
    side (one patch per neighbouring subdomain, patches in ascending neighbour order, faces in global face order).  The
    owner (lower-address) side sees the neighbour through  coupleUpper = -upper[f]  and the neighbour side through
    coupleUpper = -lower[f]  (the transposed upper for a symmetric matrix): with  Ax -= coupleUpper * xNbr  the
    decomposed Amul equals the global one.  -> (subs for MultiBlockOracle, cells of each subdomain)."""
    owner = np.asarray(owner)
    nd = int(owner.max()) + 1 if owner.size else 1
    cells = [np.nonzero(owner == d)[0] for d in range(nd)]
    local = np.empty(n, np.int64)
    for d in range(nd):
        local[cells[d]] = np.arange(cells[d].size)
    ol, ou = owner[l], owner[u]
    F = l.size

    def tr(a):
        return a.transpose(0, 2, 1) if a.ndim == 3 else a

    low = tr(upper) if lower is None else lower
    subs = []
    for d in range(nd):
        inner = np.nonzero((ol == d) & (ou == d))[0]
        subs.append(dict(n=int(cells[d].size), l=local[l[inner]].astype(np.int32), u=local[u[inner]].astype(np.int32),
                         diag=np.ascontiguousarray(diag[cells[d]]), upper=np.ascontiguousarray(upper[inner]),
                         lower=None if lower is None else np.ascontiguousarray(lower[inner]), ifaces=[], _nbr=[]))
    cut = np.nonzero(ol != ou)[0] if F else np.zeros(0, np.int64)
    for d in range(nd):
        for e in range(nd):
            if e == d:
                continue
            f_own = cut[(ol[cut] == d) & (ou[cut] == e)]   # d owns the face
            f_nei = cut[(ol[cut] == e) & (ou[cut] == d)]   # e owns the face
            ff = np.sort(np.concatenate([f_own, f_nei]))
            if ff.size == 0:
                continue
            mine = ol[ff] == d
            fc = np.where(mine, local[l[ff]], local[u[ff]]).astype(np.int32)
            cu = np.where(mine.reshape((-1,) + (1,) * (upper.ndim - 1)), -upper[ff], -low[ff])
            subs[d]["ifaces"].append(dict(faceCells=fc, peer=e, peerIface=-1, coupleUpper=np.ascontiguousarray(cu)))
            subs[d]["_nbr"].append(e)
    for d in range(nd):
        for I in subs[d]["ifaces"]:
            I["peerIface"] = subs[I["peer"]]["_nbr"].index(d)
    for s in subs:
        del s["_nbr"]
    return subs, cells
