"""GPU test of the device-side T-equation assembly on unstructured addressings (b200_sys_assemble_T with a face-interpolated
conductivity, both flux signs, several boundary faces per cell) through the C ABI against oracle/fv_oracle.c.  Kept in a
file of its own that sorts after the B200-verified GPU files: it was written after round 1's last GPU minute (its
arithmetic is covered bit-exactly by the CPU emulator, tests/test_fv_assemble.py)."""
import numpy as np
import pytest

from multiregionfoam_b200 import ldu
from oracle import pyfv, pyoracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [3, 57, 4000])   # (one isolated cell without boundary faces starts at its exact solution: 0/0 residual)
@pytest.mark.parametrize("transport", [False, True])
def test_unstructured_face_conductivity(gpu_ctx, n, transport):
    """Seeded unstructured addressing, face-interpolated conductivity, both flux signs, several boundary faces per cell:
    Amul with the device-assembled matrix is bit-exact against the oracle's; matrix and resident right-hand side together:
    the solve is bit-identical (iterations, history, field) to the one of the oracle-assembled system copied in."""
    from multiregionfoam_b200.case import Case, RankSystem, Region
    from test_fv_assemble import random_fv_case
    t = random_fv_case(n, 100 + n, transport)
    form = ldu.TEQN_TRANSPORT if transport else ldu.TEQN_CONDUCT
    rhoC, rdt, kappa = 250.0, 100.0, 5.0
    kw = dict(kappaFace=t["kappaFace"], phi=t["phi"], bCells=t["bCells"], bInt=t["bInt"], bSrc=t["bSrc"])
    d, up, lo, src = pyfv.assemble_T(int(transport), t["l"], t["u"], rhoC, rdt, kappa, t["V"], t["magSf"], t["delta"], t["Told"], **kw)
    reg = Region("r", n, t["l"], t["u"], d, up, lo, src, t["Told"].copy())
    case = Case("fv_unstructured", [RankSystem(0, 1, [reg])])
    S = ldu.LduSystem(gpu_ctx, case.ranks[0], set_coeffs=False)
    H = ldu.LduSystem(gpu_ctx, case.ranks[0])
    try:
        S.set_fv_geometry(0, t["V"], t["magSf"], t["delta"], t["bCells"], t["bInt"], t["bSrc"])
        S.upload(t["Told"], None)
        S.assemble_T(0, form, rhoC, rdt, kappa, kappaFace=t["kappaFace"], phi=t["phi"])
        S.assemble_T(0, form, rhoC, rdt, kappa)          # NULL tables: the resident kappa_f / phi are kept
        O = pyoracle.OracleSystem(case)
        x = 300.0 + 10.0 * np.random.default_rng(n).random(n)
        assert np.array_equal(S.amul(x), O.amul(x))
        H.upload(t["Told"], src)
        opts = dict(solver=ldu.SOLVER_BICGSTAB, precond=ldu.PRECOND_DILU, tolerance=1e-12, maxIter=50, history=True)
        iS, iH = S.solve_resident(**opts), H.solve_resident(**opts)
        assert iS["nIterations"] == iH["nIterations"]
        assert np.array_equal(iS["history"], iH["history"], equal_nan=True)
        assert np.array_equal(S.download(), H.download(), equal_nan=True)
    finally:
        S.close()
        H.close()
