"""Block-coupled (vector4) fixtures shared by the CPU and GPU tests (seeded, deterministic)."""
from __future__ import annotations

import numpy as np

from multiregionfoam_b200.assembly import pu_block_matrix
from multiregionfoam_b200.mesh import Block, StructuredRegion


def box_addr(nx, ny, nz):
    m = StructuredRegion("box", [Block(nx, 0.0, 1.0, 1.0)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
    return m.nCells, np.ascontiguousarray(m.lowerAddr, np.int32), np.ascontiguousarray(m.upperAddr, np.int32)


def chain_addr(n):
    l = np.arange(n - 1, dtype=np.int32)
    return n, l, l + 1


def golden(golden_addr, key):
    return int(golden_addr[f"{key}_nCells"]), golden_addr[f"{key}_l"].astype(np.int32), golden_addr[f"{key}_u"].astype(np.int32)


def addressings(golden_addr):
    e = np.empty(0, np.int32)
    return {
        "box3d": box_addr(17, 9, 5),
        "box2d": box_addr(40, 23, 1),
        "bubbleA": golden(golden_addr, "bubbleA"),          # polyhedral (2dRisingBubble fluidA)
        "duineveld0": golden(golden_addr, "duineveld0"),    # polyhedral 3-D
        "chain": chain_addr(700),
        "one_cell": (1, e, e),
        "no_faces": (37, e, e),
    }


ADDR_NAMES = ["box3d", "box2d", "bubbleA", "duineveld0", "chain", "one_cell", "no_faces"]

# (diag kind, upper/lower kind, symmetric)
KIND_COMBOS = [(16, 16, False), (16, 16, True), (4, 4, False), (4, 1, True), (1, 1, False), (16, 4, False), (4, 16, False),
               (1, 16, True), (16, 1, False)]


def random_block_coeffs(n, l, u, dK, uK, symmetric, seed=5):
    """Seeded, block diagonally dominant coefficients with the requested active types."""
    rng = np.random.default_rng(seed)
    F = l.size

    def offdiag(k):
        if k == 1:
            return -(0.5 + rng.random(F))
        if k == 4:
            return -(0.5 + rng.random((F, 4)))
        a = 0.15 * rng.standard_normal((F, 4, 4))
        for i in range(4):
            a[:, i, i] = -(0.5 + rng.random(F))
        return a

    upper = offdiag(uK)
    lower = None if symmetric else offdiag(uK)
    deg = np.bincount(l, minlength=n) + np.bincount(u, minlength=n)
    base = 1.0 + 1.6 * deg
    if dK == 1:
        diag = base + rng.random(n)
    elif dK == 4:
        diag = base[:, None] + rng.random((n, 4))
    else:
        diag = 0.2 * rng.standard_normal((n, 4, 4))
        for i in range(4):
            diag[:, i, i] = base + rng.random(n)
    return diag, upper, lower


def as_square(a, n):
    if a.ndim == 3:
        return a
    sq = np.zeros((n, 4, 4))
    for i in range(4):
        sq[:, i, i] = a if a.ndim == 1 else a[:, i]
    return sq


def scipy_block_matrix(n, l, u, diag, upper, lower):
    """The 4n x 4n sparse matrix the block coefficients stand for (independent of the oracle)."""
    import scipy.sparse as sp
    F = l.size
    D, U = as_square(diag, n), as_square(upper, F)
    L = U.transpose(0, 2, 1) if lower is None else as_square(lower, F)
    rows, cols, vals = [], [], []
    ii, jj = np.meshgrid(np.arange(4), np.arange(4), indexing="ij")

    def put(r, c, blocks):
        rows.append((4 * r[:, None, None] + ii).ravel())
        cols.append((4 * c[:, None, None] + jj).ravel())
        vals.append(blocks.ravel())

    put(np.arange(n), np.arange(n), D)
    put(u, l, L)
    put(l, u, U)
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(4 * n, 4 * n))


def pu_matrix(n, l, u, seed=2024):
    return pu_block_matrix(n, l, u, seed=seed)
