"""The LDU dump format round-trips a coupled system (incl. non-conformal GGI tables) and replays in the oracle."""
import numpy as np

from helpers import ggi_case
from multiregionfoam_b200.case import Case
from multiregionfoam_b200.dumpio import read_dump, write_dump
from oracle import pyoracle


def test_dump_roundtrip_and_replay(tmp_path, golden_addr):
    case = ggi_case(golden_addr)
    O = pyoracle.OracleSystem(case)
    ctl = dict(solver="BiCGStab", preconditioner="DILU", tolerance=1e-10, relTol=0.0, minIter=0, maxIter=400)
    x, info = O.solve(case.concat("psi"), case.concat("source"), "BiCGStab", "DILU", tolerance=1e-10, maxIter=400)
    p = str(tmp_path / "sys.b200ldu")
    write_dump(p, case.ranks[0], ctl, info["history"])
    rs, ctl2, hist = read_dump(p)
    assert ctl2 == ctl and np.array_equal(hist, info["history"])
    for a, b in zip(case.ranks[0].regions, rs.regions):
        assert np.array_equal(a.lowerAddr, b.lowerAddr) and np.array_equal(a.diag, b.diag)
        assert (a.lower is None) == (b.lower is None)
        for ia, ib in zip(a.interfaces, b.interfaces):
            assert np.array_equal(ia.ggiWeights, ib.ggiWeights) and ia.nPeerFaces == ib.nPeerFaces
    O2 = pyoracle.OracleSystem(Case("replay", [rs]))
    x2, info2 = O2.solve(np.concatenate([r.psi for r in rs.regions]), np.concatenate([r.source for r in rs.regions]),
                         ctl2["solver"], ctl2["preconditioner"], tolerance=ctl2["tolerance"], maxIter=ctl2["maxIter"])
    assert np.array_equal(info2["history"], hist) and np.array_equal(x2, x)
