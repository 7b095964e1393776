"""The decomposed block-coupled oracle (oracle/pyblk_multi.py) against the one-domain oracle: a BlockLduMatrix<vector4>
split over subdomains with processor patches must give the same Amul, and - with a preconditioner that has no
off-diagonal part - the same Krylov iterates up to summation order (the parallel form of
/root/reference/filesToReplace/fvBlockMatrix.C:1360-1388)."""
import numpy as np
import pytest

from block_helpers import box_addr, random_block_coeffs
from oracle import pyblk
from oracle.pyblk_multi import MultiBlockOracle, split_block_system


def owners(n, nd, kind, seed=3):
    if kind == "slabs":
        return (np.arange(n) * nd // n).astype(np.int64)
    return np.random.default_rng(seed).integers(0, nd, n)


@pytest.mark.parametrize("kinds", [(16, 16, False), (16, 16, True), (4, 1, True), (16, 4, False)])
@pytest.mark.parametrize("nd,how", [(2, "slabs"), (3, "slabs"), (3, "random")])
def test_decomposed_amul_equals_global(kinds, nd, how):
    dK, uK, sym = kinds
    n, l, u = box_addr(9, 7, 4)
    diag, upper, lower = random_block_coeffs(n, l, u, dK, uK, sym)
    own = owners(n, nd, how)
    subs, cells = split_block_system(n, l, u, diag, upper, lower, own)
    for d, s in enumerate(subs):   # the patch pairs match face by face
        for I in s["ifaces"]:
            J = subs[I["peer"]]["ifaces"][I["peerIface"]]
            assert J["peer"] == d and J["faceCells"].size == I["faceCells"].size
    x = np.random.default_rng(1).standard_normal((n, 4))
    G = pyblk.BlockOracle(l, u, n, diag, upper, lower)
    M = MultiBlockOracle(subs)
    ys = M.amul([x[c] for c in cells])
    y = np.empty_like(x)
    for c, yy in zip(cells, ys):
        y[c] = yy
    ref = G.amul(x)
    assert np.max(np.abs(y - ref)) <= 1e-13 * np.max(np.abs(ref))


@pytest.mark.parametrize("solver,sym", [("BiCGStab", False), ("CG", True)])
def test_decomposed_solve_with_diagonal_precon_follows_global(solver, sym):
    n, l, u = box_addr(8, 6, 5)
    diag, upper, lower = random_block_coeffs(n, l, u, 16, 16 if not sym else 1, sym)
    if sym:  # CG needs a symmetric positive definite matrix: symmetric diagonal blocks
        diag = 0.5 * (diag + diag.transpose(0, 2, 1))
    rng = np.random.default_rng(2)
    b, x0 = rng.standard_normal((n, 4)), rng.standard_normal((n, 4))
    own = owners(n, 3, "slabs")
    subs, cells = split_block_system(n, l, u, diag, upper, lower, own)
    G = pyblk.BlockOracle(l, u, n, diag, upper, lower)
    M = MultiBlockOracle(subs)
    xg, pg = G.solve(x0, b, solver, "diagonal", tolerance=1e-10, maxIter=200)
    xs, pm = M.solve([x0[c] for c in cells], [b[c] for c in cells], solver, "diagonal", tolerance=1e-10, maxIter=200)
    assert abs(pm["normFactor"] - pg["normFactor"]) <= 1e-12 * pg["normFactor"]
    k = min(15, pg["history"].shape[0], pm["history"].shape[0])
    assert np.max(np.abs(pm["history"][:k] - pg["history"][:k]) / np.maximum(pg["history"][:k], 1e-300)) < 1e-8
    xm = np.empty_like(xg)
    for c, xx in zip(cells, xs):
        xm[c] = xx
    assert np.linalg.norm(xm - xg) <= 1e-7 * np.linalg.norm(xg)


def test_cholesky_precon_is_subdomain_local():
    """BlockCholeskyPrecon sees the subdomain's own faces only: the decomposed run converges, along another path."""
    n, l, u = box_addr(8, 6, 5)
    diag, upper, lower = random_block_coeffs(n, l, u, 16, 16, False)
    rng = np.random.default_rng(4)
    b, x0 = rng.standard_normal((n, 4)), np.zeros((n, 4))
    subs, cells = split_block_system(n, l, u, diag, upper, lower, owners(n, 2, "slabs"))
    M = MultiBlockOracle(subs)
    xs, pm = M.solve([x0[c] for c in cells], [b[c] for c in cells], "BiCGStab", "Cholesky", tolerance=1e-9, maxIter=200)
    assert pm["converged"]
    G = pyblk.BlockOracle(l, u, n, diag, upper, lower)
    xm = np.empty((n, 4))
    for c, xx in zip(cells, xs):
        xm[c] = xx
    r = b - G.amul(xm)
    assert np.abs(r).sum() / pm["normFactor"] < 1e-8
