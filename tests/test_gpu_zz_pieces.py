"""regionCouple pairs spread over ranks by the decomposition (the shipped ``simple; n (2 1 2)`` cuts the
fluid and the plate at different x): the zone-piece interface of the C ABI
(b200_sys_set_interface_pieces) on ONE device - all sub-domains as regions of one system - and on four processes that
share cuda:0 and exchange the patch values over the peer-to-peer transport.  Against the oracle on the same decomposition
(tests/test_decompose_pieces.py ties that to the serial case)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case
from multiregionfoam_b200.decompose import decompose_cht_simple, flatten_ranks
from oracle import pyoracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [(2, 1, 2), (3, 1, 1)])
def test_zone_pieces_on_one_device(gpu_ctx, n):
    case, fluid, solid = cht_case(1, 4)
    dec = decompose_cht_simple(case, fluid, solid, n)
    flat = flatten_ranks(dec)
    assert any(getattr(itf, "pieces", None) for reg in flat.ranks[0].regions for itf in reg.interfaces)
    O = pyoracle.OracleSystem(dec)
    S = ldu.LduSystem(gpu_ctx, flat.ranks[0])
    try:
        x0, b = dec.concat("psi"), dec.concat("source")
        v = np.random.default_rng(3).standard_normal(O.n) * 10 + 300
        assert np.array_equal(S.amul(v), O.amul(v))
        assert np.array_equal(S.residual(v, b), O.residual(v, b))
        xo, io = O.solve(x0, b, "BiCGStab", "DILU", tolerance=1e-12, maxIter=400)
        xg, ig = S.solve(x0, b, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=1e-12, maxIter=400)
        k = min(21, ig["history"].size, io["history"].size)
        assert k > 5
        assert np.max(np.abs(ig["history"][:k] - io["history"][:k]) / np.abs(io["history"][:k])) < 1e-10
        assert np.linalg.norm(xg - xo) <= 1e-8 * np.linalg.norm(xo)
    finally:
        S.close()


def test_piece_errors(gpu_ctx):
    case, fluid, solid = cht_case(1, 2)
    dec = decompose_cht_simple(case, fluid, solid, (2, 1, 1))
    flat = flatten_ranks(dec)
    import copy
    bad = copy.deepcopy(flat.ranks[0])
    itf = next(i for reg in bad.regions for i in reg.interfaces if getattr(i, "pieces", None))
    h, pr, pi, za = itf.pieces[0]
    itf.pieces[0] = (h, pr, pi, np.concatenate([za[:-1], za[:1]]))     # a zone face held twice
    with pytest.raises(ldu.B200Error):
        ldu.LduSystem(gpu_ctx, bad)
    bad = copy.deepcopy(flat.ranks[0])
    itf = next(i for reg in bad.regions for i in reg.interfaces if getattr(i, "pieces", None) and i.ggiAddr.size)
    itf.pieces = [p for p in itf.pieces if itf.ggiAddr[0] not in p[3]]   # the tables address a face nobody holds
    with pytest.raises(ldu.B200Error):
        ldu.LduSystem(gpu_ctx, bad)


def test_selfpeer_2x1x2_decomposition():
    world, uid = 4, os.urandom(128).hex()
    with tempfile.TemporaryDirectory() as work:
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "scripts", "selfpeer_pieces.py"), work, str(rk), str(world), uid, "2", "1", "2"],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for rk in range(world)]
        outs = []
        try:
            for p in procs:
                outs.append(p.communicate(timeout=600)[0])
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
        assert all(p.returncode == 0 for p in procs), "\n".join(f"[rank {i} rc {p.returncode}] " + o[-2500:] for i, (p, o) in enumerate(zip(procs, outs)))
        assert "pieces_selfpeer_ok=True" in outs[0], outs[0][-2000:]
