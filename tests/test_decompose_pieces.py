"""regionCouple pairs whose two regions are cut at different places (the shipped ``n (2 1 2)`` decomposition,
/root/reference/tutorials/conjugateHeatTransfer/flowOverHeatedPlate/system/fluid/decomposeParDict:17-27): the
decomposition describes such a pair through zone pieces (decompose.py), the oracle evaluates it on the receiving side
(orc_set_iface_zone).  The decomposed product must equal the serial one: same terms, per row the same order."""
import numpy as np
import pytest

from multiregionfoam_b200.assembly import cht_case
from multiregionfoam_b200.decompose import decompose_cht_simple
from oracle import pyoracle


@pytest.mark.parametrize("n", [(2, 1, 2), (3, 1, 1), (2, 1, 1)])
def test_decomposed_pieces_amul_equals_serial(n):
    case, fluid, solid = cht_case(1, 4)
    dec = decompose_cht_simple(case, fluid, solid, n)
    # the two regions are cut at different x: some interface faces face another rank
    spread = [itf for rk in dec.ranks for reg in rk.regions for itf in reg.interfaces if getattr(itf, "pieces", None)]
    assert spread and any(any(p[0] != rk.rank for p in itf.pieces) for rk in dec.ranks for reg in rk.regions for itf in reg.interfaces
                          if getattr(itf, "pieces", None))
    S, D = pyoracle.OracleSystem(case), pyoracle.OracleSystem(dec)
    rng = np.random.default_rng(0)
    xg = [rng.standard_normal(r.nCells) for r in case.ranks[0].regions]
    yg = S.amul(np.concatenate(xg))
    off = np.cumsum([0] + [r.nCells for r in case.ranks[0].regions])
    xd = np.concatenate([xg[ri][reg.globalCells] for rk in dec.ranks for ri, reg in enumerate(rk.regions)])
    yd = D.amul(xd)
    pos = 0
    for rk in dec.ranks:
        for ri, reg in enumerate(rk.regions):
            ref = yg[off[ri] + reg.globalCells]
            got = yd[pos:pos + reg.nCells]
            pos += reg.nCells
            assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(yg))
    # and the decomposed solve converges to the serial solution (block-Jacobi DILU: another path, same fixed point)
    b = np.concatenate([r.source for r in case.ranks[0].regions])
    bd = np.concatenate([reg.source for rk in dec.ranks for reg in rk.regions])
    xs, _ = S.solve(np.zeros_like(b), b, "BiCGStab", "DILU", tolerance=1e-12, maxIter=500)
    xdd, info = D.solve(np.zeros_like(bd), bd, "BiCGStab", "DILU", tolerance=1e-12, maxIter=500)
    pos = 0
    for rk in dec.ranks:
        for ri, reg in enumerate(rk.regions):
            assert np.linalg.norm(xdd[pos:pos + reg.nCells] - xs[off[ri] + reg.globalCells]) <= 1e-8 * np.linalg.norm(xs)
            pos += reg.nCells
