"""GaussSeidelSmoother / smoothSolver through include/b200_smooth.h (first piece) against the oracle:
a sweep is BIT-EXACT (the level-scheduled kernel subtracts a row's terms in the reference order), smoothSolver histories
within 1e-10, on structured, 2-D, polyhedral and degenerate addressings, symmetric and asymmetric."""
import numpy as np
import pytest

from block_helpers import ADDR_NAMES, addressings
from multiregionfoam_b200 import ldu, smoother
from oracle import pyoracle
from test_gs_oracle import coeffs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("symmetric", [True, False])
@pytest.mark.parametrize("name", ADDR_NAMES)
def test_gauss_seidel_sweeps_bit_exact(gpu_ctx, golden_addr, name, symmetric):
    n, l, u = addressings(golden_addr)[name]
    diag, upper, lower = coeffs(n, l, u, symmetric, seed=4)
    G = smoother.GaussSeidel(gpu_ctx, l, u, n)
    try:
        G.set_coeffs(diag, upper, lower)
        rng = np.random.default_rng(8)
        b, x0 = rng.standard_normal(n), rng.standard_normal(n) * 3 + 1
        for k in (1, 2, 5):
            assert np.array_equal(G.smooth(x0, b, k), pyoracle.gs_smooth(l, u, diag, upper, lower, x0, b, k)), (name, k)
        assert np.array_equal(G.sweep(x0, b), pyoracle.gs_smooth(l, u, diag, upper, lower, x0, b, 1))
    finally:
        G.close()


@pytest.mark.parametrize("name,nSweeps", [("box3d", 1), ("bubbleA", 2), ("duineveld0", 4), ("chain", 1)])
def test_smooth_solver_history(gpu_ctx, golden_addr, name, nSweeps):
    n, l, u = addressings(golden_addr)[name]
    diag, upper, lower = coeffs(n, l, u, name != "duineveld0", seed=6)
    G = smoother.GaussSeidel(gpu_ctx, l, u, n)
    try:
        G.set_coeffs(diag, upper, lower)
        rng = np.random.default_rng(9)
        b, x0 = rng.standard_normal(n), rng.standard_normal(n)
        xg, ig = G.solve(x0, b, nSweeps, tolerance=1e-10, maxIter=400)
        xo, io = pyoracle.gs_solve(l, u, diag, upper, lower, x0, b, nSweeps, tolerance=1e-10, maxIter=400)
        assert ig["nIterations"] == io["nIterations"] and ig["converged"] == io["converged"]
        assert abs(ig["normFactor"] - io["normFactor"]) <= 1e-12 * io["normFactor"]
        assert np.max(np.abs(ig["history"] - io["history"]) / io["history"]) < 1e-10
        assert np.array_equal(xg, xo)      # every sweep is bit-exact, so is the field
    finally:
        G.close()


def test_smoother_errors(gpu_ctx):
    with pytest.raises(ldu.B200Error):
        smoother.GaussSeidel(gpu_ctx, np.array([1], np.int32), np.array([0], np.int32), 2)   # not upper-triangular
    G = smoother.GaussSeidel(gpu_ctx, np.array([0], np.int32), np.array([1], np.int32), 2)
    try:
        with pytest.raises(ldu.B200Error):
            G.smooth(np.zeros(2), np.zeros(2), 1)                                            # coefficients not set
    finally:
        G.close()


@pytest.mark.parametrize("pre,name", [(ldu.PRECOND_DILU, "DILU"), (ldu.PRECOND_DIC, "DIC")])
def test_dic_dilu_smoother_is_residual_precondition_add(gpu_ctx, pre, name):
    """DICSmoother / DILUSmoother::smooth = the preconditioner's sweeps applied to lduMatrix::residual, added to psi - through
    b200_smooth on the coupled two-region CHT system (regionCouple interface in the residual), bit-exact against the oracle's
    residual + precondition, sweep by sweep."""
    from multiregionfoam_b200.assembly import cht_case
    case, _, _ = cht_case(1, 3)
    if name == "DIC":   # DIC needs symmetric rows
        for reg in case.ranks[0].regions:
            reg.lower = None
    O = pyoracle.OracleSystem(case)
    S = ldu.LduSystem(gpu_ctx, case.ranks[0])
    try:
        x0, b = case.concat("psi"), case.concat("source")
        O.precond_setup(name)
        ref = x0.copy()
        for k in range(1, 4):
            ref = ref + O.precondition(O.residual(ref, b))
            assert np.array_equal(S.smooth(pre, x0, b, k), ref), (name, k)
        # it is a smoother: the residual norm goes down
        r0, r3 = np.abs(O.residual(x0, b)).sum(), np.abs(O.residual(ref, b)).sum()
        assert r3 < 0.5 * r0
    finally:
        S.close()


def test_smooth_solver_mirror_selects_by_dictionary(gpu_ctx):
    """`solver smoothSolver; smoother ...; nSweeps n;` as the tutorials' fvSolution write it, through the Python mirror of the
    selection (solvers.smoothSolver): GaussSeidel and DICGaussSeidel on one symmetric region, DILU on the coupled asymmetric
    CHT system; every variant against the oracle composition of the same steps."""
    from multiregionfoam_b200 import solvers
    from multiregionfoam_b200.assembly import cht_case
    from multiregionfoam_b200.case import Case, RankSystem
    case, _, _ = cht_case(1, 2)
    d = solvers.parse_dictionary("T { solver smoothSolver; smoother DILU; nSweeps 2; tolerance 1e-7; relTol 0; maxIter 60; }")["T"]
    S = ldu.LduSystem(gpu_ctx, case.ranks[0])
    O = pyoracle.OracleSystem(case)
    try:
        x, b = case.concat("psi").copy(), case.concat("source")
        perf = solvers.smoothSolver.New("T", S, d).solve(x, b)
        O.precond_setup("DILU")
        ref = case.concat("psi").copy()
        for _ in range(perf.nIterations):
            ref = ref + O.precondition(O.residual(ref, b))
        assert perf.nIterations % 2 == 0 and perf.nIterations > 0
        assert np.array_equal(x, ref)
        assert perf.finalResidual < perf.initialResidual
        assert "smoothSolver:  Solving for T" in perf.line()
        with pytest.raises(solvers.FatalError):
            solvers.smoothSolver.New("T", S, dict(d, smoother="DIC"))       # asymmetric rows
        with pytest.raises(solvers.FatalError):
            solvers.smoothSolver.New("T", S, dict(d, smoother="symGaussSeidel"))
    finally:
        S.close()
    # one symmetric region without coupled patches: GaussSeidel and DICGaussSeidel
    solid = case.ranks[0].regions[1]
    import copy
    reg = copy.copy(solid)
    reg.interfaces = []
    one = Case("solid", [RankSystem(0, 1, [reg])])
    S1, O1 = ldu.LduSystem(gpu_ctx, one.ranks[0]), pyoracle.OracleSystem(one)
    try:
        for name in ("GaussSeidel", "DICGaussSeidel"):
            sol = solvers.smoothSolver.New("T", S1, dict(solver="smoothSolver", smoother=name, nSweeps=1, tolerance=0.0, maxIter=4))
            x, b = reg.psi.copy(), reg.source
            perf = sol.solve(x, b)
            sol.close()
            assert perf.nIterations == 4
            ref = reg.psi.copy()
            O1.precond_setup("DIC")
            for _ in range(4):
                if name == "DICGaussSeidel":
                    ref = ref + O1.precondition(O1.residual(ref, b))
                ref = pyoracle.gs_smooth(reg.lowerAddr, reg.upperAddr, reg.diag, reg.upper, reg.lower, ref, b, 1)
            assert np.array_equal(x, ref), name
    finally:
        S1.close()
