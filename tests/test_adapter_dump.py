"""The foam-extend adapter's interface extraction and dump writer, run for real against mock objects.

adapters/b200LduSolvers/b200Binding.C (describe + dump) is compiled with tests/cpp/adapter_dump_harness.cpp, which backs
the stand-in foam-extend declarations (adapters/foamStub) with plain tables: two coupled rows joined by a regionCouple pair
with a non-conformal GGI (3 master faces against 2 slave faces), a processor patch per row on rank 1 of 2.  The B200LDU1
file the C++ writes is read back by multiregionfoam_b200.dumpio (the reader a foam-extend dump will meet) and must
describe exactly that system; the oracle then applies it, which checks the GGI tables end to end."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("adapter") / "adapter_dump_harness")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-Wno-unused-parameter", "-I", os.path.join(ROOT, "adapters", "foamStub"),
           "-I", os.path.join(ROOT, "adapters", "b200LduSolvers"), "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(ROOT, "tests", "cpp", "adapter_dump_harness.cpp"), os.path.join(ROOT, "adapters", "b200LduSolvers", "b200Binding.C")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_adapter_describes_and_dumps_the_coupled_system(harness, tmp_path):
    from multiregionfoam_b200 import dumpio
    from multiregionfoam_b200.case import PROCESSOR, REGION_COUPLE
    path = str(tmp_path / "T_proc1.b200ldu")
    r = subprocess.run([harness, path, "1"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stderr == "", r.stderr
    rs, ctl, hist = dumpio.read_dump(path)
    assert (rs.rank, rs.nRanks, len(rs.regions)) == (1, 2, 2)
    assert ctl == dict(solver="BiCGStab", preconditioner="Cholesky", tolerance=1e-15, relTol=0.0, minIter=0, maxIter=200)
    assert np.array_equal(hist, [0.5, 0.25, 0.125])
    fluid, solid = rs.regions
    assert (fluid.nCells, fluid.nFaces, fluid.lower is not None) == (6, 5, True)      # asymmetric row: lower present
    assert (solid.nCells, solid.nFaces, solid.lower is None) == (4, 3, True)          # symmetric row: upper only
    assert np.array_equal(fluid.lowerAddr, [0, 1, 2, 3, 4]) and np.array_equal(fluid.upperAddr, [1, 2, 3, 4, 5])
    assert np.array_equal(fluid.diag, [10, 11, 12, 13, 14, 15]) and np.array_equal(fluid.lower, [-0.5, -0.75, -1, -1.25, -1.5])
    assert np.array_equal(fluid.source, [1, 2, 3, 4, 5, 6]) and np.array_equal(solid.psi, [310, 311, 312, 313])
    # the uncoupled wall patch (fluid patch 0) is no interface; the regionCouple patch is interface 0, the processor patch 1
    fi, fp = fluid.interfaces
    si, sp = solid.interfaces
    assert (fi.kind, fi.peerRank, fi.peerRegion, fi.peerIface, fi.nPeerFaces) == (REGION_COUPLE, 1, 1, 0, 2)
    assert (si.kind, si.peerRank, si.peerRegion, si.peerIface, si.nPeerFaces) == (REGION_COUPLE, 1, 0, 0, 3)
    assert np.array_equal(fi.faceCells, [1, 2, 3]) and np.array_equal(si.faceCells, [0, 1])
    assert np.array_equal(fi.bouCoeffs, [0.1, 0.2, 0.3]) and np.array_equal(fi.intCoeffs, [1.1, 1.2, 1.3])
    # master side: masterAddr / masterWeights of patchToPatch(); slave side: slaveAddr / slaveWeights of the master's
    assert np.array_equal(fi.ggiOffsets, [0, 1, 3, 4]) and np.array_equal(fi.ggiAddr, [0, 0, 1, 1])
    assert np.allclose(fi.ggiWeights, [1, .5, .5, 1], rtol=0, atol=0)
    assert np.array_equal(si.ggiOffsets, [0, 2, 4]) and np.array_equal(si.ggiAddr, [0, 1, 1, 2])
    assert np.array_equal(si.ggiWeights, [2 / 3, 1 / 3, 1 / 3, 2 / 3])
    # processor patches: the neighbour is rank 0, whose interface index (exchanged through Pstream) is 1 in both rows
    assert (fp.kind, fp.peerRank, fp.peerRegion, fp.peerIface, fp.nPeerFaces) == (PROCESSOR, 0, 0, 1, 1)
    assert (sp.kind, sp.peerRank, sp.peerRegion, sp.peerIface, sp.nPeerFaces) == (PROCESSOR, 0, 1, 1, 2)
    assert np.array_equal(sp.faceCells, [2, 3]) and np.array_equal(sp.bouCoeffs, [0.8, 0.9])


def test_dumped_region_couple_tables_drive_the_oracle(harness, tmp_path):
    """The regionCouple part of the dumped system through the oracle's coupled Amul: the interface update of the fluid
    row is coeffs * (GGI-weighted solid values), monolithicCouplingFvPatchField.C:400-404, 443-453."""
    from multiregionfoam_b200 import dumpio
    from multiregionfoam_b200.case import Case, RankSystem
    from oracle import pyoracle
    path = str(tmp_path / "T.b200ldu")
    assert subprocess.run([harness, path, "1"]).returncode == 0
    rs, _, _ = dumpio.read_dump(path)
    for reg in rs.regions:                       # single-rank replay: drop the processor patches
        reg.interfaces = [i for i in reg.interfaces if i.peerRank == rs.rank]
        for i in reg.interfaces:
            i.peerRank = 0
    case = Case("dump", [RankSystem(0, 1, rs.regions)])
    O = pyoracle.OracleSystem(case)
    x = np.concatenate([reg.psi for reg in rs.regions])
    y = O.amul(x)
    fluid, solid = rs.regions
    xf, xs_ = x[:6], x[6:]
    yf = fluid.diag * xf
    np.add.at(yf, fluid.upperAddr, fluid.lower * xf[fluid.lowerAddr])
    np.add.at(yf, fluid.lowerAddr, fluid.upper * xf[fluid.upperAddr])
    own = xs_[solid.interfaces[0].faceCells]                                    # the solid side's patchInternalField
    pnf = np.array([own[0], 0.5 * own[0] + 0.5 * own[1], own[1]])              # onto the 3 fluid faces
    yf[fluid.interfaces[0].faceCells] -= fluid.interfaces[0].bouCoeffs * pnf
    assert np.allclose(y[:6], yf, rtol=1e-14, atol=0)


def test_adapter_describes_a_pair_spread_over_processors(harness, tmp_path):
    """!localParallel(): the adapter takes the rows zoneAddressing() of the ZONE-level interpolator, the size of the shadow zone,
    and - through the second Pstream::gatherList - who holds which shadow zone faces (b200_sys_set_interface_pieces)."""
    r = subprocess.run([harness, str(tmp_path / "unused"), "1", "zone"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stderr == "", r.stderr
    lines = {tuple(l.split()[1:3]): l for l in r.stdout.splitlines() if l.startswith("iface ")}
    fluid, solid = lines[("0", "0")], lines[("1", "0")]
    # master side, local faces = master zone faces 2, 3, 4 -> rows 2, 3, 4 of masterAddr; shadow zone (slave) has 4 faces,
    # of which rank 0 holds 0, 1 (its interface 0 of row 1) and this rank 2, 3
    assert "kind 0 zoneMode 1 nPeerFaces 4 peer 1 0 offsets 0 2 4 5 addr 1 2 2 3 3 w 0.25 0.75 0.5 0.5 1 pieces [0 0: 0 1] [1 0: 2 3]" in fluid
    # slave side, local faces = slave zone faces 2, 3 -> rows 2, 3 of the master's slaveAddr; shadow zone (master) has 5 faces
    assert "kind 0 zoneMode 1 nPeerFaces 5 peer 0 0 offsets 0 2 4 addr 2 3 3 4 w 0.375 0.625 0.5 0.5 pieces [0 0: 0 1] [1 0: 2 3 4]" in solid
    # the processor patches are described as before
    assert "kind 1 zoneMode 0" in lines[("0", "1")] and "kind 1 zoneMode 0" in lines[("1", "1")]
