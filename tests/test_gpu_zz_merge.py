"""Sweep warps filled across plane ends (schedule.hpp, mergeSequences): the layout large systems get (C3, 64 M cells),
forced here on a small case (B200_MERGE_MIN_GROUPS=0) at several gaps - Amul, the DILU / DIC reciprocal diagonal and the
sweeps must stay bit-exact against the oracle, the BiCGStab history within 1e-10."""
import numpy as np
import pytest

from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case
from oracle import pyoracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gap", ["0.5", "1.5", "4"])
def test_merged_warps_parity(gpu_ctx, monkeypatch, gap):
    monkeypatch.setenv("B200_MERGE_MIN_GROUPS", "0")
    monkeypatch.setenv("B200_MERGE_GAP", gap)
    case, _, _ = cht_case(1, 14)
    O = pyoracle.OracleSystem(case)
    S = ldu.LduSystem(gpu_ctx, case.ranks[0])
    try:
        x0, b = case.concat("psi"), case.concat("source")
        v = np.random.default_rng(5).standard_normal(O.n) * 10 + 300
        assert np.array_equal(S.amul(v), O.amul(v))
        for pre, name in ((ldu.PRECOND_DILU, "DILU"), (ldu.PRECOND_DIC, "DIC")):
            O.precond_setup(name)
            assert np.array_equal(S.rD(pre), O.rD())
            assert np.array_equal(S.precondition(pre, v), O.precondition(v))
        xo, io = O.solve(x0, b, "BiCGStab", "DILU", tolerance=1e-12, maxIter=300)
        xg, ig = S.solve(x0, b, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=1e-12, maxIter=300)
        k = min(21, ig["history"].size, io["history"].size)
        assert np.max(np.abs(ig["history"][:k] - io["history"][:k]) / np.abs(io["history"][:k])) < 1e-10
        assert np.linalg.norm(xg - xo) <= 1e-8 * np.linalg.norm(xo)
    finally:
        S.close()
