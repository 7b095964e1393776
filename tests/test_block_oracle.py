"""CPU tests of the block-coupled (vector4) oracle and of the fvBlockMatrix host mirror.

The reference holds no golden vector for this path (parity unpinned), so the oracle is pinned to
independent restatements: scipy block-CSR products, a dense numpy evaluation of the BlockCholesky recurrences, the
factored form (P^-1 + L) P (P^-1 + U) of the preconditioner and known solutions."""
import numpy as np
import pytest

from block_helpers import (ADDR_NAMES, KIND_COMBOS, addressings, as_square, box_addr, pu_matrix, random_block_coeffs,
                           scipy_block_matrix)
from multiregionfoam_b200 import blockldu
from multiregionfoam_b200.blockldu import BlockCoupling, ScalarEqn, fvBlockMatrix
from multiregionfoam_b200.solvers import FatalError
from oracle import pyblk


@pytest.mark.parametrize("combo", KIND_COMBOS)
@pytest.mark.parametrize("name", ["box3d", "bubbleA", "chain", "no_faces"])
def test_amul_matches_scipy(golden_addr, name, combo):
    n, l, u = addressings(golden_addr)[name]
    dK, uK, sym = combo
    diag, upper, lower = random_block_coeffs(n, l, u, dK, uK, sym)
    O = pyblk.BlockOracle(l, u, n, diag, upper, lower)
    A = scipy_block_matrix(n, l, u, diag, upper, lower)
    x = np.random.default_rng(1).standard_normal((n, 4))
    y = O.amul(x)
    ref = (A @ x.ravel()).reshape(n, 4)
    assert np.allclose(y, ref, rtol=1e-13, atol=1e-12)


def test_inv4_against_numpy():
    rng = np.random.default_rng(3)
    for _ in range(50):
        a = rng.standard_normal((4, 4)) + 3 * np.eye(4)
        assert np.allclose(pyblk.inv4(a), np.linalg.inv(a), rtol=1e-11, atol=1e-13)
    # needs pivoting: zero leading entry
    a = np.array([[0.0, 2, 0, 0], [1, 0, 0, 0], [0, 0, 0, 4], [0, 0, 5, 1]])
    assert np.allclose(pyblk.inv4(a) @ a, np.eye(4), atol=1e-15)


def dense_cholesky_diag(n, l, u, diag, upper, lower):
    """BlockCholeskyPrecon::calcPreconDiag written out with numpy (square working type)."""
    F = l.size
    D = as_square(diag, n).copy()
    U = as_square(upper, F)
    L = U.transpose(0, 2, 1) if lower is None else as_square(lower, F)
    for f in range(F):
        D[u[f]] -= L[f] @ np.linalg.inv(D[l[f]]) @ U[f]
    return np.linalg.inv(D)


@pytest.mark.parametrize("combo", KIND_COMBOS)
def test_cholesky_diag_and_factored_form(golden_addr, combo):
    n, l, u = addressings(golden_addr)["box3d"]
    dK, uK, sym = combo
    diag, upper, lower = random_block_coeffs(n, l, u, dK, uK, sym)
    O = pyblk.BlockOracle(l, u, n, diag, upper, lower)
    pD = O.precon_diag("Cholesky")
    assert pD.shape[1] == max(dK, uK)
    P = as_square(pD if pD.shape[1] != 16 else pD.reshape(n, 4, 4), n) if pD.shape[1] != 1 else as_square(pD[:, 0], n)
    assert np.allclose(P, dense_cholesky_diag(n, l, u, diag, upper, lower), rtol=1e-10, atol=1e-12)
    # M w = r with M = (P^-1 + L) P (P^-1 + U)
    import scipy.sparse as sp
    Pinv = np.linalg.inv(P)
    z = np.zeros((l.size, 4, 4))
    e = np.empty(0, np.int32)
    A = scipy_block_matrix(n, l, u, diag, upper, lower)
    Lm = sp.tril(A, k=-1)
    Um = sp.triu(A, k=1)
    # strictly block-lower / block-upper parts: remove the in-block off-diagonal entries of the diagonal blocks
    Dblk = scipy_block_matrix(n, e, e, as_square(diag, n), z[:0], z[:0])
    Lm = Lm - sp.tril(Dblk, k=-1)
    Um = Um - sp.triu(Dblk, k=1)
    Pm = scipy_block_matrix(n, e, e, P, z[:0], z[:0])
    Pim = scipy_block_matrix(n, e, e, Pinv, z[:0], z[:0])
    r = np.random.default_rng(2).standard_normal((n, 4))
    w = O.precondition(r, "Cholesky")
    Mw = (Pim + Lm) @ (Pm @ ((Pim + Um) @ w.ravel()))
    assert np.allclose(Mw, r.ravel(), rtol=1e-10, atol=1e-10)
    # diagonal preconditioner = inverse of the diagonal blocks
    wd = O.precondition(r, "diagonal")
    assert np.allclose(wd, np.einsum("nij,nj->ni", np.linalg.inv(as_square(diag, n)), r), rtol=1e-12, atol=1e-13)
    assert np.array_equal(O.precondition(r, "none"), r)


def test_cholesky_exact_on_block_chain():
    """On a chain the incomplete factorisation is complete: one application solves the system."""
    n = 60
    l = np.arange(n - 1, dtype=np.int32)
    u = l + 1
    diag, upper, lower = random_block_coeffs(n, l, u, 16, 16, False, seed=9)
    O = pyblk.BlockOracle(l, u, n, diag, upper, lower)
    x = np.random.default_rng(4).standard_normal((n, 4))
    assert np.allclose(O.precondition(O.amul(x), "Cholesky"), x, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("pre", ["none", "diagonal", "Cholesky"])
def test_bicgstab_recovers_known_solution(pre):
    n, l, u = box_addr(14, 9, 6)
    M = pu_matrix(n, l, u)
    O = pyblk.BlockOracle(M.l, M.u, n, M.diag, M.upper, M.lower)
    x, info = O.solve(M.psi, M.source, "BiCGStab", pre, tolerance=1e-11, maxIter=500)
    assert info["converged"] and info["history"].shape == (info["nIterations"] + 1, 4)
    assert np.linalg.norm(x - M.xstar) / np.linalg.norm(M.xstar) < 1e-8
    # true residual agrees with the recurrence residual the solver reports
    nf = O.norm_factor(M.psi, M.source)
    true = np.abs(M.source - O.amul(x)).sum(axis=0) / nf
    assert np.all(true < 50 * max(info["finalResidual"].max(), 1e-13))


def test_cg_on_symmetric_block_system(golden_addr):
    n, l, u = addressings(golden_addr)["box3d"]
    rng = np.random.default_rng(8)
    # SPD: symmetric diagonal blocks, lower = upper^T
    diag, upper, _ = random_block_coeffs(n, l, u, 16, 16, True, seed=8)
    diag = 0.5 * (diag + diag.transpose(0, 2, 1))
    O = pyblk.BlockOracle(l, u, n, diag, upper, None)
    xs = rng.standard_normal((n, 4))
    b = O.amul(xs)
    x, info = O.solve(np.zeros((n, 4)), b, "CG", "Cholesky", tolerance=1e-12, maxIter=300)
    assert info["converged"] and np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-9
    x2, info2 = O.solve(np.zeros((n, 4)), b, "CG", "none", tolerance=1e-12, maxIter=300)
    assert info2["nIterations"] > info["nIterations"]


def test_stop_rules():
    n, l, u = box_addr(8, 5, 3)
    M = pu_matrix(n, l, u)
    O = pyblk.BlockOracle(M.l, M.u, n, M.diag, M.upper, M.lower)
    _, info = O.solve(M.psi, M.source, "BiCGStab", "Cholesky", tolerance=1e-30, maxIter=7)
    assert info["nIterations"] == 7 and not info["converged"]
    _, info = O.solve(M.xstar, M.source, "BiCGStab", "Cholesky", tolerance=1e-6, minIter=3, maxIter=50)
    assert info["nIterations"] == 3  # already converged, minIter forces three iterations
    _, info = O.solve(M.psi, M.source, "BiCGStab", "Cholesky", tolerance=0.0, relTol=1e-3, maxIter=100)
    assert info["converged"] and info["finalResidual"].max() <= 1e-3 * info["initialResidual"].max()


# ------------------------------------------------------------------------------------ fvBlockMatrix host mirror
def test_fvblockmatrix_active_types_follow_the_reference():
    n, l, u = box_addr(6, 4, 2)
    F = l.size
    rng = np.random.default_rng(0)
    M = fvBlockMatrix(l, u, n)
    upU, loU = rng.random(F), rng.random(F)
    M.insertEquation(0, ScalarEqn(rng.random(n) + 5, np.zeros((n, 3)), upU, loU), nCmpts=3)
    # first insertion: UNALLOCATED -> SCALAR upper (fvBlockMatrix.C:181-184); diag linear; lower allocated from upper
    assert M.diag.shape == (n, 4) and M.upper.shape == (F,) and M.lower.shape == (F, 4)
    assert np.array_equal(M.lower[:, 0], loU) and np.array_equal(M.lower[:, 3], upU)
    upP = rng.random(F)
    M.insertEquation(3, ScalarEqn(rng.random(n) + 5, np.ones(n), upP, None))
    assert M.upper.shape == (F, 4) and np.array_equal(M.upper[:, 3], upP) and np.array_equal(M.upper[:, 1], upU)
    assert np.array_equal(M.lower[:, 3], upP) and np.array_equal(M.source[:, 3], np.ones(n))
    g = BlockCoupling(rng.random((n, 3)), rng.random((F, 3)), rng.random((F, 3)))
    M.insertBlockCoupling(0, 3, g, True)
    M.insertBlockCoupling(3, 0, g, False)
    assert M.diag.shape == (n, 4, 4) and M.upper.shape == (F, 4, 4) and M.lower.shape == (F, 4, 4)
    assert np.array_equal(M.upper[:, :3, 3], g.upper) and np.array_equal(M.lower[:, 3, :3], g.lower)
    assert np.count_nonzero(M.upper[0]) == 10
    with pytest.raises(FatalError):
        M.insertBlockCoupling(2, 2, g, True)


def test_fvblockmatrix_symmetric_stays_symmetric():
    n, l, u = box_addr(5, 3, 1)
    M = fvBlockMatrix(l, u, n)
    M.insertEquation(0, ScalarEqn(np.ones(n), np.zeros(n), -np.ones(l.size), None))
    M.insertEquation(1, ScalarEqn(np.ones(n), np.zeros(n), -2 * np.ones(l.size), None))
    assert M.symmetric() and M.upper.shape == (l.size, 4)


def test_unknown_block_solver_names_are_fatal():
    n, l, u = box_addr(4, 3, 1)
    M = pu_matrix(n, l, u)
    with pytest.raises(FatalError, match="Unknown matrix solver"):
        M.solve(None, {"solver": "GMRES", "preconditioner": "Cholesky"})
    with pytest.raises(FatalError, match="Unknown matrix preconditioner"):
        M.solve(None, {"solver": "BiCGStab", "preconditioner": "ILUC0"})
    with pytest.raises(FatalError, match="asymmetric"):
        M.solve(None, {"solver": "CG", "preconditioner": "Cholesky"})
    assert "cudaBlockBiCGStab" in blockldu.BLOCK_SOLVER_TABLE and "cudaBlockCholesky" in blockldu.BLOCK_PRECOND_TABLE
