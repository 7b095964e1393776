"""Processor decomposition (decomposePar semantics) and the directly assembled rank slabs."""
import numpy as np

from helpers import random_vec, rel_l2
from multiregionfoam_b200.assembly import cht_case, cht_rank_slab
from multiregionfoam_b200.case import Case, PROCESSOR
from multiregionfoam_b200.decompose import decompose_cht_zslabs
from multiregionfoam_b200.mesh import is_upper_triangular
from oracle import pyoracle


def test_decomposed_addressing_is_upper_triangular_and_patches_last():
    case, fluid, solid = cht_case(1, 6)
    dec = decompose_cht_zslabs(case, fluid, solid, 3)
    assert dec.nCells == case.nCells
    for rk in dec.ranks:
        for reg in rk.regions:
            assert is_upper_triangular(reg.lowerAddr, reg.upperAddr)
            kinds = [i.kind for i in reg.interfaces]
            assert kinds == sorted(kinds)  # processor patches after the regionCouple patch
            for itf in reg.interfaces:
                peer = dec.ranks[itf.peerRank].regions[itf.peerRegion].interfaces[itf.peerIface]
                assert peer.peerRank == rk.rank and peer.nFaces == itf.nFaces
                if itf.kind == PROCESSOR:
                    # both sides list the cut faces in the same (global face) order
                    g_mine = reg.globalCells[itf.faceCells]
                    g_peer = dec.ranks[itf.peerRank].regions[itf.peerRegion].globalCells[peer.faceCells]
                    src = case.ranks[0].regions[itf.peerRegion]
                    key = src.lowerAddr.astype(np.int64) * src.nCells + src.upperAddr
                    pair = np.minimum(g_mine, g_peer).astype(np.int64) * src.nCells + np.maximum(g_mine, g_peer)
                    assert np.all(np.isin(pair, key))  # every pair is a face of the undecomposed mesh


def test_rank_slabs_equal_decomposition_of_the_global_case():
    n, Lloc = 3, 2
    case, fluid, solid = cht_case(1, Lloc * n, z1=0.4 * n)
    dec = decompose_cht_zslabs(case, fluid, solid, n)
    slabs = Case("slabs", [cht_rank_slab(1, Lloc, g, n) for g in range(n)])
    for a, b in zip(dec.ranks, slabs.ranks):
        for ra, rb in zip(a.regions, b.regions):
            assert np.array_equal(ra.lowerAddr, rb.lowerAddr) and np.array_equal(ra.upperAddr, rb.upperAddr)
            assert np.allclose(ra.diag, rb.diag, rtol=1e-13, atol=0)
            assert np.allclose(ra.upper, rb.upper, rtol=1e-13, atol=0)
            assert (ra.lower is None) == (rb.lower is None)
            assert np.allclose(ra.source, rb.source, rtol=1e-13, atol=0)
            assert len(ra.interfaces) == len(rb.interfaces)
            for ia, ib in zip(ra.interfaces, rb.interfaces):
                assert (ia.kind, ia.peerRank, ia.peerRegion, ia.peerIface) == (ib.kind, ib.peerRank, ib.peerRegion, ib.peerIface)
                assert np.array_equal(ia.faceCells, ib.faceCells)
                assert np.allclose(ia.bouCoeffs, ib.bouCoeffs, rtol=1e-13, atol=0)
                assert np.allclose(ia.intCoeffs, ib.intCoeffs, rtol=1e-13, atol=0)
    # and the oracle accepts them as one coupled system
    O = pyoracle.OracleSystem(slabs)
    Od = pyoracle.OracleSystem(dec)
    x = random_vec(O.n, 1)
    assert rel_l2(O.amul(x), Od.amul(x)) < 1e-13


def test_decomposed_solve_is_block_jacobi_but_converges_to_the_same_field():
    case, fluid, solid = cht_case(1, 4)
    dec = decompose_cht_zslabs(case, fluid, solid, 2)
    Os, Od = pyoracle.OracleSystem(case), pyoracle.OracleSystem(dec)
    xs, infos = Os.solve(case.concat("psi"), case.concat("source"), "BiCGStab", "DILU", tolerance=1e-13, maxIter=300)
    xd, infod = Od.solve(dec.concat("psi"), dec.concat("source"), "BiCGStab", "DILU", tolerance=1e-13, maxIter=300)
    assert infos["converged"] and infod["converged"]
    glob = dec.to_global(xd, [r.nCells for r in case.ranks[0].regions])
    assert rel_l2(np.concatenate(glob), xs) < 1e-9
