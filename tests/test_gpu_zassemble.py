"""GPU tests of the device-side T-equation assembly (b200_sys_set_fv_geometry / b200_sys_assemble_T, SURVEY 8(f)
rank 3) through the C ABI, against oracle/fv_oracle.c.

The assembled coefficients never leave the device, so they are checked through what the solver does with them:
 * Amul with the device-assembled matrix == oracle Amul with the oracle-assembled matrix: BIT-EXACT;
 * the coupled solve from the device-assembled matrix and right-hand side == the solve of a second system whose
   coefficients and b were assembled by the oracle and copied in with b200_sys_set_coeffs / b200_upload:
   identical iteration counts, residual histories and fields (same kernels, same inputs -> bit-identical);
 * a second time step (T.oldTime() = the resident solution, nothing but the call crosses the bus) against the same
   step driven from the host.
"""
import copy

import numpy as np
import pytest

from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import assemble_cht, cht_fv_tables
from multiregionfoam_b200.mesh import flow_over_heated_plate
from oracle import pyfv, pyoracle

pytestmark = pytest.mark.gpu


def oracle_step(case, meshes, tables, Told_regions):
    """The case with every region's coefficients / source re-assembled by the fv oracle from Told."""
    out = copy.deepcopy(case)
    for reg, mesh, t, Told in zip(out.ranks[0].regions, meshes, tables, Told_regions):
        d, up, lo, src = pyfv.assemble_T(t["form"], mesh.lowerAddr, mesh.upperAddr, t["rhoC"], t["rDeltaT"], t["kappa"], t["V"],
                                         t["magSf"], t["deltaCoeffs"], Told, phi=t["phi"], bCells=t["bCells"], bInt=t["bInt"],
                                         bSrc=t["bSrc"])
        reg.diag, reg.upper, reg.lower, reg.source, reg.psi = d, up, lo, src, np.array(Told, dtype=np.float64)
    return out


def device_assemble(S, tables, first):
    for r, t in enumerate(tables):
        S.assemble_T(r, t["form"], t["rhoC"], t["rDeltaT"], t["kappa"], phi=t["phi"] if first else None)


@pytest.mark.parametrize("r,layers", [(1, 1), (1, 4)])
def test_device_assembly_matches_oracle(gpu_ctx, r, layers):
    fluid, solid = flow_over_heated_plate(r, layers)
    meshes, tables = (fluid, solid), cht_fv_tables(fluid, solid)
    case = assemble_cht(fluid, solid)
    rng = np.random.default_rng(3)
    T0 = [t["T0"] + rng.random(m.nCells) for t, m in zip(tables, meshes)]
    ref = oracle_step(case, meshes, tables, T0)

    S = ldu.LduSystem(gpu_ctx, case.ranks[0], set_coeffs=False)   # device-assembled
    H = ldu.LduSystem(gpu_ctx, ref.ranks[0])                       # host-assembled (oracle) and copied in
    try:
        for ri, t in enumerate(tables):
            S.set_fv_geometry(ri, t["V"], t["magSf"], t["deltaCoeffs"], t["bCells"], t["bInt"], t["bSrc"])
            for i, itf in enumerate(case.ranks[0].regions[ri].interfaces):
                S.set_interface_coeffs(ri, i, itf.bouCoeffs, itf.intCoeffs)
        launches0 = gpu_ctx.launches
        S.upload(np.concatenate(T0), None)
        device_assemble(S, tables, first=True)
        assert gpu_ctx.launches > launches0

        # 1. the matrix: Amul bit-exact against the oracle's product with the oracle-assembled coefficients
        O = pyoracle.OracleSystem(ref)
        x = 300.0 + 10.0 * rng.random(O.n)
        assert np.array_equal(S.amul(x), O.amul(x))

        # 2. matrix and right-hand side: two time steps, device-driven vs host-driven
        S.upload(np.concatenate(T0), None)
        device_assemble(S, tables, first=False)   # amul used the vectors; assemble again from the re-uploaded field
        H.upload(np.concatenate(T0), ref.concat("source"))
        kw = dict(solver=ldu.SOLVER_BICGSTAB, precond=ldu.PRECOND_DILU, tolerance=1e-12, maxIter=200, history=True)
        for step in range(2):
            iS, iH = S.solve_resident(**kw), H.solve_resident(**kw)
            xS, xH = S.download(), H.download()
            assert iS["nIterations"] == iH["nIterations"] and iS["nIterations"] > 0
            assert np.array_equal(iS["history"][: iS["nIterations"] + 1], iH["history"][: iH["nIterations"] + 1])
            assert np.array_equal(xS, xH)
            # and against the oracle solver on the oracle matrix (north_star tolerances)
            xo, io = pyoracle.OracleSystem(ref).solve(ref.concat("psi"), ref.concat("source"), "BiCGStab", "DILU",
                                                      tolerance=1e-12, maxIter=200)
            assert np.linalg.norm(xS - xo) / np.linalg.norm(xo) < 1e-8
            # next step: the device side needs nothing from the host; the host side re-assembles and copies
            device_assemble(S, tables, first=False)
            ref = oracle_step(case, meshes, tables, H.split(xH))
            for ri, reg in enumerate(ref.ranks[0].regions):
                H.set_coeffs(ri, reg.diag, reg.upper, reg.lower)
            H.upload(None, ref.concat("source"))
    finally:
        S.close()
        H.close()


def test_assemble_errors(gpu_ctx):
    fluid, solid = flow_over_heated_plate(1, 1)
    tables = cht_fv_tables(fluid, solid)
    case = assemble_cht(fluid, solid)
    S = ldu.LduSystem(gpu_ctx, case.ranks[0], set_coeffs=False)
    try:
        with pytest.raises(ldu.B200Error) as e:   # no geometry yet
            S.assemble_T(0, ldu.TEQN_TRANSPORT, 250.0, 100.0, 5.0)
        assert e.value.code == -5
        t = tables[0]
        S.set_fv_geometry(0, t["V"], t["magSf"], t["deltaCoeffs"], t["bCells"], t["bInt"], t["bSrc"])
        with pytest.raises(ldu.B200Error) as e:   # transport form without a flux
            S.assemble_T(0, ldu.TEQN_TRANSPORT, 250.0, 100.0, 5.0)
        assert e.value.code == -5
        with pytest.raises(ldu.B200Error) as e:   # unknown form
            S.assemble_T(0, 7, 250.0, 100.0, 5.0)
        assert e.value.code == -1
        with pytest.raises(ldu.B200Error) as e:   # boundary cell out of range
            S.set_fv_geometry(0, t["V"], t["magSf"], t["deltaCoeffs"], np.array([fluid.nCells], np.int32), np.ones(1), np.ones(1))
        assert e.value.code == -1
    finally:
        S.close()
