"""The CPU oracle against independent known answers (scipy CSR algebra, exact solutions).
The reference ships no golden vectors for the linear solve (parity unpinned, oracle/ldu_oracle.h);
these tests pin the restatement to algebraic identities instead."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from helpers import chain_region, ggi_case, golden_region, random_vec, rel_l2
from multiregionfoam_b200.assembly import cht_case, single_region_case
from multiregionfoam_b200.decompose import decompose_cht_zslabs
from oracle import pyoracle


def csr_of_case(case):
    """Global sparse matrix of a (possibly decomposed) case incl. interface couplings."""
    offs = case.row_offsets()
    nReg = case.nRegions
    rows, cols, vals = [], [], []
    for rk in case.ranks:
        for ri, reg in enumerate(rk.regions):
            o = offs[rk.rank * nReg + ri]
            lo = reg.upper if reg.lower is None else reg.lower
            rows += [o + np.arange(reg.nCells), o + reg.lowerAddr, o + reg.upperAddr]
            cols += [o + np.arange(reg.nCells), o + reg.upperAddr, o + reg.lowerAddr]
            vals += [reg.diag, reg.upper, lo]
            for itf in reg.interfaces:
                peer = case.ranks[itf.peerRank].regions[itf.peerRegion]
                po = offs[itf.peerRank * nReg + itf.peerRegion]
                pfc = peer.interfaces[itf.peerIface].faceCells
                if itf.ggiOffsets is None:
                    rows.append(o + itf.faceCells)
                    cols.append(po + pfc)
                    vals.append(-itf.bouCoeffs)
                else:
                    for i in range(itf.nFaces):
                        for k in range(itf.ggiOffsets[i], itf.ggiOffsets[i + 1]):
                            rows.append(np.array([o + itf.faceCells[i]]))
                            cols.append(np.array([po + pfc[itf.ggiAddr[k]]]))
                            vals.append(np.array([-itf.bouCoeffs[i] * itf.ggiWeights[k]]))
    n = offs[-1]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


@pytest.fixture(scope="module")
def cht1():
    return cht_case(1, 1)[0]


def test_amul_matches_csr(cht1):
    O = pyoracle.OracleSystem(cht1)
    A = csr_of_case(cht1)
    x = random_vec(O.n, 1)
    assert rel_l2(O.amul(x), A @ x) < 1e-14
    assert rel_l2(O.tmul(x), A.T @ x) < 1e-14  # bouCoeffs == intCoeffs here
    assert rel_l2(O.sumA(), A @ np.ones(O.n)) < 1e-13
    b = random_vec(O.n, 2)
    assert rel_l2(O.residual(x, b), b - A @ x) < 1e-13


def test_amul_ggi_nonconformal_matches_csr(golden_addr):
    case = ggi_case(golden_addr)
    O = pyoracle.OracleSystem(case)
    A = csr_of_case(case)
    x = random_vec(O.n, 3)
    assert rel_l2(O.amul(x), A @ x) < 1e-14


def test_decomposed_amul_equals_serial():
    case, fluid, solid = cht_case(1, 4)
    dec = decompose_cht_zslabs(case, fluid, solid, 2)
    Os, Od = pyoracle.OracleSystem(case), pyoracle.OracleSystem(dec)
    xg = [random_vec(r.nCells, 10 + i) for i, r in enumerate(case.ranks[0].regions)]
    xd = np.concatenate([xg[ri][reg.globalCells] for rk in dec.ranks for ri, reg in enumerate(rk.regions)])
    ys = Os.amul(np.concatenate(xg))
    yd = dec.to_global(Od.amul(xd), [r.nCells for r in case.ranks[0].regions])
    assert rel_l2(np.concatenate(yd), ys) < 1e-14


@pytest.mark.parametrize("sym", [True, False])
def test_dic_dilu_are_exact_on_a_chain(sym):
    # for a tridiagonal matrix the diagonal-based incomplete factorisation is the exact LDU
    reg = chain_region(500, symmetric=sym)
    case = single_region_case(reg)
    O = pyoracle.OracleSystem(case)
    A = csr_of_case(case)
    O.precond_setup("DIC" if sym else "DILU")
    r = random_vec(O.n, 5)
    w = O.precondition(r)
    assert rel_l2(A @ w, r) < 1e-10
    if not sym:
        assert rel_l2(A.T @ O.preconditionT(r), r) < 1e-10


def test_dilu_matches_explicit_factor(golden_addr):
    # M = (D + L) D^-1 (D + U) with D from the DILU recurrence; check M w = r on an unstructured mesh
    reg = golden_region(golden_addr, "bubbleA", symmetric=False)
    case = single_region_case(reg)
    O = pyoracle.OracleSystem(case)
    O.precond_setup("DILU")
    rD = O.rD()
    n = reg.nCells
    L = sp.csr_matrix((reg.lower, (reg.upperAddr, reg.lowerAddr)), shape=(n, n))
    U = sp.csr_matrix((reg.upper, (reg.lowerAddr, reg.upperAddr)), shape=(n, n))
    D = sp.diags(1.0 / rD)
    M = (D + L) @ sp.diags(rD) @ (D + U)
    r = random_vec(n, 6)
    assert rel_l2(M @ O.precondition(r), r) < 1e-11
    # and the recurrence itself: D_u = diag_u - sum_{l<u} upper*lower/D_l
    d = reg.diag.copy()
    for f in range(reg.nFaces):
        d[reg.upperAddr[f]] -= reg.upper[f] * reg.lower[f] / d[reg.lowerAddr[f]]
    assert np.array_equal(1.0 / d, rD)


@pytest.mark.parametrize("solver,precond", [("PCG", "DIC"), ("BiCGStab", "DILU"), ("PBiCG", "DILU"), ("BiCGStab", "Cholesky"),
                                            ("PCG", "diagonal"), ("BiCGStab", "none")])
def test_solvers_converge_to_direct_solution(golden_addr, solver, precond):
    reg = golden_region(golden_addr, "bubbleB", symmetric=(solver == "PCG"))
    case = single_region_case(reg)
    O = pyoracle.OracleSystem(case)
    A = csr_of_case(case)
    x, info = O.solve(reg.psi, reg.source, solver, precond, tolerance=1e-13, maxIter=2000)
    assert info["converged"], info
    xd = spla.spsolve(A.tocsc(), reg.source)
    assert rel_l2(x, xd) < 1e-8
    h = info["history"]
    assert h.size == info["nIterations"] + 1 and h[0] == info["initialResidual"] and h[-1] == info["finalResidual"]


def test_coupled_cht_solution_and_norm_factor(cht1):
    O = pyoracle.OracleSystem(cht1)
    A = csr_of_case(cht1)
    x0, b = cht1.concat("psi"), cht1.concat("source")
    x, info = O.solve(x0, b, "BiCGStab", "Cholesky", tolerance=1e-15, maxIter=200)
    xd = spla.spsolve(A.tocsc(), b)
    assert rel_l2(x, xd) < 1e-9
    # normFactor definition
    xRef = x0.mean()
    tmp = A @ np.full(O.n, xRef)
    nf = np.abs(A @ x0 - tmp).sum() + np.abs(b - tmp).sum() + 1e-20
    assert abs(info["normFactor"] - nf) / nf < 1e-12
    assert abs(info["initialResidual"] - np.abs(b - A @ x0).sum() / nf) < 1e-12


def test_two_slab_conduction_exact():
    """1-D steady conduction through two slabs in series, coupled by a regionCouple interface with the
    harmonic conductance: the discrete solution is piecewise linear and known in closed form."""
    from multiregionfoam_b200.case import Case, Interface, RankSystem, REGION_COUPLE, Region
    n, kA, kB, h = 20, 5.0, 100.0, 0.05
    TL, TR = 300.0, 310.0

    def slab(k, Tleft=None, Tright=None):
        l = np.arange(n - 1, dtype=np.int32)
        D = np.full(n - 1, k / h)
        diag = np.bincount(l, weights=D, minlength=n) + np.bincount(l + 1, weights=D, minlength=n)
        src = np.zeros(n)
        if Tleft is not None:
            diag[0] += 2 * k / h
            src[0] += 2 * k / h * Tleft
        if Tright is not None:
            diag[-1] += 2 * k / h
            src[-1] += 2 * k / h * Tright
        return Region("slab", n, l, l + 1, diag, -D, None, src, np.full(n, 305.0))

    A, B = slab(kA, Tleft=TL), slab(kB, Tright=TR)
    c = 1.0 / (0.5 * h / kA + 0.5 * h / kB)
    A.diag[-1] += c
    B.diag[0] += c
    cc = np.array([c])
    A.interfaces.append(Interface(REGION_COUPLE, np.array([n - 1], np.int32), cc, cc, 0, 1, 0))
    B.interfaces.append(Interface(REGION_COUPLE, np.array([0], np.int32), cc, cc, 0, 0, 0))
    case = Case("slabs", [RankSystem(0, 1, [A, B])])
    O = pyoracle.OracleSystem(case)
    x, info = O.solve(case.concat("psi"), case.concat("source"), "PCG", "DIC", tolerance=1e-14, maxIter=500)
    q = (TR - TL) / (n * h / kA + n * h / kB)           # heat flux
    xc = (np.arange(n) + 0.5) * h
    exact = np.concatenate([TL + q * xc / kA, TR - q * (n * h - xc) / kB])
    assert np.max(np.abs(x - exact)) < 1e-8


def test_stop_rules_and_defaults(golden_addr):
    reg = golden_region(golden_addr, "bubbleB", symmetric=True)
    O = pyoracle.OracleSystem(single_region_case(reg))
    # maxIter 0: no iteration, residuals equal
    x, info = O.solve(reg.psi, reg.source, "PCG", "DIC", tolerance=0.0, maxIter=0)
    assert info["nIterations"] == 0 and np.array_equal(x, reg.psi)
    # minIter forces iterations even when already converged
    xs, _ = O.solve(reg.psi, reg.source, "PCG", "DIC", tolerance=1e-14, maxIter=500)
    _, info2 = O.solve(xs, reg.source, "PCG", "DIC", tolerance=1.0, minIter=3, maxIter=10)
    assert info2["nIterations"] == 3
    # relTol
    _, info3 = O.solve(reg.psi, reg.source, "PCG", "DIC", tolerance=0.0, relTol=1e-3, maxIter=500)
    assert info3["converged"] and info3["finalResidual"] <= 1e-3 * info3["initialResidual"]


def test_face_transfer_functions():
    rng = np.random.default_rng(0)
    # GGI weighted gather, vector field, ragged incl. an empty row
    offs = np.array([0, 2, 2, 5], np.int32)
    addr = np.array([0, 3, 1, 2, 3], np.int32)
    w = rng.random(5)
    ff = rng.random((4, 3))
    out = pyoracle.ggi_interpolate(offs, addr, w, ff, nComp=3)
    exp = np.array([ff[0] * w[0] + ff[3] * w[1], np.zeros(3), ff[1] * w[2] + ff[2] * w[3] + ff[3] * w[4]])
    assert np.allclose(out, exp, rtol=0, atol=1e-16)
    # patchFaceToGlobal over 3 ranks / globalFaceToPatch round trip
    perm = rng.permutation(10).astype(np.int32)
    po = np.array([0, 4, 4, 10], np.int32)
    pf = rng.random(10)
    g = pyoracle.patch_face_to_global(po, perm, pf, 10)
    assert np.array_equal(g[perm], pf)
    assert np.array_equal(pyoracle.global_face_to_patch(perm[4:], g), pf[4:])
    assert np.array_equal(pyoracle.direct_map(perm, pf), pf[perm])


def test_threaded_oracle_is_independent_of_thread_count():
    """The worker pool standing in for MPI ranks (bench.py's decomposed CPU arm) must not change a single bit:
    per-row partial sums are combined in row / rank order whatever thread produced them."""
    from multiregionfoam_b200.assembly import cht_rank_slab
    from multiregionfoam_b200.case import Case
    case = Case("slabs", [cht_rank_slab(1, 2, g, 4) for g in range(4)])
    O = pyoracle.OracleSystem(case)
    x0, b = case.concat("psi"), case.concat("source")
    ref = None
    try:
        for th in (1, 3, 8):
            pyoracle.set_threads(th)
            x, info = O.solve(x0.copy(), b, "BiCGStab", "DILU", tolerance=1e-12, maxIter=60)
            if ref is None:
                ref = (x, info["history"])
            else:
                assert np.array_equal(x, ref[0]) and np.array_equal(info["history"], ref[1])
    finally:
        pyoracle.set_threads(1)
