"""The Gauss-Seidel smoother / smoothSolver oracle (oracle/ldu_oracle.c orc_gs_smooth / orc_gs_solve)
against known answers: one sweep equals the textbook  (D + L) psi_new = b - U psi_old,  the iteration converges to the
direct solution on diagonally dominant matrices, a triangular matrix is solved in one sweep."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from block_helpers import box_addr, chain_addr
from oracle import pyoracle


def coeffs(n, l, u, symmetric, seed=0):
    rng = np.random.default_rng(seed)
    F = l.size
    upper = -(0.5 + rng.random(F))
    lower = None if symmetric else -(0.5 + rng.random(F))
    deg = np.bincount(l, minlength=n) + np.bincount(u, minlength=n)
    diag = 1.0 + 1.7 * deg + rng.random(n)
    return diag, upper, lower


def csr(n, l, u, diag, upper, lower):
    lo = upper if lower is None else lower
    return sp.csr_matrix((np.concatenate([diag, upper, lo]), (np.concatenate([np.arange(n), l, u]), np.concatenate([np.arange(n), u, l]))), shape=(n, n))


@pytest.mark.parametrize("symmetric", [True, False])
def test_one_sweep_is_the_forward_gauss_seidel_step(symmetric):
    n, l, u = box_addr(7, 5, 4)
    diag, upper, lower = coeffs(n, l, u, symmetric)
    A = csr(n, l, u, diag, upper, lower)
    rng = np.random.default_rng(1)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    x1 = pyoracle.gs_smooth(l, u, diag, upper, lower, x0, b, 1)
    DL, U = sp.tril(A, 0).tocsr(), sp.triu(A, 1).tocsr()
    ref = sla.spsolve_triangular(DL, b - U @ x0, lower=True)
    assert np.max(np.abs(x1 - ref)) <= 1e-13 * np.max(np.abs(ref))
    x3 = pyoracle.gs_smooth(l, u, diag, upper, lower, x0, b, 3)
    y = x0
    for _ in range(3):
        y = pyoracle.gs_smooth(l, u, diag, upper, lower, y, b, 1)
    assert np.array_equal(x3, y)


def test_smooth_solver_converges_to_the_direct_solution():
    n, l, u = box_addr(9, 6, 5)
    diag, upper, lower = coeffs(n, l, u, False, seed=3)
    A = csr(n, l, u, diag, upper, lower)
    rng = np.random.default_rng(2)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    x, info = pyoracle.gs_solve(l, u, diag, upper, lower, x0, b, nSweeps=2, tolerance=1e-12, maxIter=500)
    assert info["converged"] and info["nIterations"] % 2 == 0
    assert np.all(np.diff(info["history"]) < 0)
    assert np.linalg.norm(x - sla.spsolve(A.tocsc(), b)) <= 1e-9 * np.linalg.norm(x)
    # normFactor and the initial residual as lduMatrix::solver::normFactor defines them
    Ax, pA = A @ x0, A @ np.full(n, x0.mean())
    nf = np.abs(Ax - pA).sum() + np.abs(b - pA).sum() + 1e-20
    assert abs(info["normFactor"] - nf) <= 1e-12 * nf
    assert abs(info["initialResidual"] - np.abs(b - Ax).sum() / nf) <= 1e-12


def test_lower_triangular_matrix_is_solved_by_one_sweep_and_limits():
    n, l, u = chain_addr(50)
    diag, upper, lower = coeffs(n, l, u, False)
    upper[:] = 0.0
    b = np.random.default_rng(5).standard_normal(n)
    x = pyoracle.gs_smooth(l, u, diag, upper, lower, np.zeros(n), b, 1)
    assert np.max(np.abs(csr(n, l, u, diag, upper, lower) @ x - b)) < 1e-13
    # maxIter bounds the number of sweeps; minIter forces them
    _, info = pyoracle.gs_solve(l, u, diag, upper, lower, np.zeros(n), b, nSweeps=3, tolerance=1e-30, maxIter=7)
    assert info["nIterations"] == 9 and not info["converged"]
    _, info = pyoracle.gs_solve(l, u, diag, upper, lower, x, b, nSweeps=1, tolerance=1.0, minIter=2, maxIter=10)
    assert info["nIterations"] == 2
