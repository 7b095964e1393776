"""BASELINE config 4 on the GPU, scaled down (HronTurekFsi3 topology at refinement 1, two z-layers: 11 932 cells; the
bench runs it at 7.9 M): two Dirichlet-Neumann coupling iterations of the partitioned loop - three PBiCG + DILU component
solves of U, PCG + DIC for p, the interface field through the non-conformal GGI (84 x 2 fluid faces against 216 x 2 solid
faces), three PCG + DIC component solves of D, the displacement back - every solve and every transfer through the C ABI,
against the CPU oracle doing the same loop: transfers bit-exact, residual histories within 1e-10, fields within 1e-8."""
import numpy as np
import pytest

from helpers import rel_l2
from multiregionfoam_b200 import fsi
from multiregionfoam_b200.case import Case, RankSystem
from oracle import pyoracle

gpu = pytest.mark.gpu


@gpu
def test_partitioned_fsi_loop_on_hron_turek_topology(gpu_ctx):
    case = fsi.fsi_case(*fsi.FSI_WORKLOADS["C4-mini"])
    assert (case.nFluid, case.nSolid) == (5336 * 2, 630 * 2)
    O = {k: pyoracle.OracleSystem(Case(k, [RankSystem(0, 1, [reg])])) for k, reg in (("U", case.fluidU), ("p", case.fluidP), ("D", case.solidD))}
    hist_o, transfers_o = [], []

    def solve_o(key, x0, b, solver, precond, iters):
        x, info = O[key].solve(x0, b, solver, precond, tolerance=0.0, minIter=iters, maxIter=iters)
        hist_o.append(info["history"])
        return x

    def transfer_o(tab, f):
        out = pyoracle.ggi_interpolate(tab[0], tab[1], tab[2], f, 3)
        transfers_o.append(out)
        return out

    D = fsi.DeviceFsi(gpu_ctx, case)
    transfers_g = []

    def transfer_g(tab, f):
        out = D.transfer(tab, f)
        transfers_g.append(out)
        return out

    try:
        so, sg = fsi.initial_state(case), fsi.initial_state(case)
        for _ in range(2):
            so = fsi.coupling_iteration(case, solve_o, transfer_o, so)
            sg = fsi.coupling_iteration(case, D.solve, transfer_g, sg)
        assert len(D.histories) == len(hist_o) == 14
        for hg, ho in zip(D.histories, hist_o):
            k = min(21, hg.size, ho.size)
            assert np.max(np.abs(hg[:k] - ho[:k]) / np.abs(ho[:k])) < 1e-10
        for key in ("U", "p", "D"):
            assert rel_l2(sg[key].ravel(), so[key].ravel()) < 1e-8
        # the transfers see inputs that agree to ~1e-15, and are themselves exact: compare on identical input
        f = np.random.default_rng(1).random((case.fluidFaceCells.size, 3))
        assert np.array_equal(D.transfer(case.toSolid, f), pyoracle.ggi_interpolate(*case.toSolid, f, 3))
        for tg, to in zip(transfers_g, transfers_o):
            assert rel_l2(tg.ravel(), to.ravel()) < 1e-12
    finally:
        D.close()


def test_topology_matches_the_blockmeshdict():
    """5 336 + 630 cells (C4 base sizes), every block face matched, interface 84 against 216 faces, GGI rows sum to one."""
    from multiregionfoam_b200.blockmesh2d import HT_FLUID_INTERFACE, HT_SOLID_INTERFACE, hron_turek
    from multiregionfoam_b200.mesh import is_upper_triangular
    fluid, solid = hron_turek(1, 1)
    assert (fluid.nCells, solid.nCells) == (5336, 630)
    assert is_upper_triangular(fluid.lowerAddr, fluid.upperAddr) and is_upper_triangular(solid.lowerAddr, solid.upperAddr)
    deg = np.bincount(fluid.lowerAddr, minlength=fluid.nCells) + np.bincount(fluid.upperAddr, minlength=fluid.nCells)
    assert deg.max() == 4 and deg.min() == 2
    assert fluid.patch_cells(HT_FLUID_INTERFACE).size == 84 and solid.patch_cells(HT_SOLID_INTERFACE).size == 216
    case = fsi.fsi_case(1, 1)
    for off, addr, w in (case.toSolid, case.toFluid):
        assert np.allclose(np.add.reduceat(w, off[:-1]), 1.0, rtol=0, atol=1e-14)
