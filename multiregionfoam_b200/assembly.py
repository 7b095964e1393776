"""Synthetic finite-volume assembly of the monolithic conjugate-heat-transfer T system.

Produces the matrices that ``coupledFvMatrix<scalar>::solve`` hands to the coupled Krylov
solver for the flowOverHeatedPlate topology:

  fluid  (transportTemperature, /root/reference/src/regions/transportTemperature/transportTemperature.C:129-150)
         rho*cp*(ddt(T) + div(phi,T)) = laplacian(k,T),  U = (1,0,0) prescribed, upwind  -> asymmetric
  solid  (conductTemperature, /root/reference/src/regions/conductTemperature/conductTemperature.C:135-152)
         rho*cv*ddt(T) = laplacian(k,T)                                                  -> symmetric
  fluid 'interface' <-> solid 'top': regionCouple patch pair; series (harmonic) conductance on both
         sides (monolithicThermalDiffusivityFvPatchScalarField.C:257-276 without radiation), conformal
         => identity GGI addressing, weights 1.

Constants: constant/{fluid,solid}/transportProperties (rho=1, cp=250, k=5 | rho=1, cv=100, k=100),
system/controlDict deltaT 1e-2, 0/*/orig/monolithic/T (inlet 300 K, solid bottom 310 K).
The matrices are representative of, not identical to, the reference's (identical ones need the
foam-extend dump path).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .case import Case, Interface, PROCESSOR, RankSystem, Region, REGION_COUPLE
from .mesh import StructuredRegion, flow_over_heated_plate


def _laplacian_ddt(reg: StructuredRegion, rho_c: float, k: float, dt: float, T_old: float):
    D = k * reg.faceArea / reg.faceDelta
    l, u = reg.lowerAddr, reg.upperAddr
    diag = rho_c * reg.volume / dt
    source = diag * T_old
    diag = diag + np.bincount(l, weights=D, minlength=reg.nCells) + np.bincount(u, weights=D, minlength=reg.nCells)
    return D, diag, source


def assemble_cht(fluid: StructuredRegion, solid: StructuredRegion, *, dt: float = 1e-2,
                 Ux: float = 1.0, name: str = "cht") -> Case:
    rho_f, cp_f, k_f = 1.0, 250.0, 5.0
    rho_s, cv_s, k_s = 1.0, 100.0, 100.0
    T_f0, T_s0, T_in, T_bot = 300.0, 310.0, 300.0, 310.0

    # ------------------------------------------------------------------ fluid
    D, diag_f, src_f = _laplacian_ddt(fluid, rho_f * cp_f, k_f, dt, T_f0)
    l, u = fluid.lowerAddr, fluid.upperAddr
    # convective flux rho*cp*U.S_f through +x faces (owner -> neighbour), upwind
    F = np.where(fluid.faceDir == 0, rho_f * cp_f * Ux * fluid.faceArea, 0.0)
    upper_f = -D + np.minimum(F, 0.0)
    lower_f = -D - np.maximum(F, 0.0)
    diag_f = diag_f + np.bincount(l, weights=np.maximum(F, 0.0), minlength=fluid.nCells) \
                    - np.bincount(u, weights=np.minimum(F, 0.0), minlength=fluid.nCells)
    # inlet: fixedValue T_in (diffusion + convective inflow)
    c, a, d = fluid.side_xmin()
    Db = k_f * a / d
    Fb = -rho_f * cp_f * Ux * a            # outward normal is -x
    np.add.at(diag_f, c, Db + np.maximum(Fb, 0.0))
    np.add.at(src_f, c, (Db - np.minimum(Fb, 0.0)) * T_in)
    # outlet: zeroGradient, convective outflow
    c, a, d = fluid.side_xmax()
    np.add.at(diag_f, c, np.maximum(rho_f * cp_f * Ux * a, 0.0))

    # ------------------------------------------------------------------ solid
    Ds, diag_s, src_s = _laplacian_ddt(solid, rho_s * cv_s, k_s, dt, T_s0)
    upper_s = -Ds
    c, a, d = solid.side_y(0, top=False)   # bottom: fixedValue 310
    Db = k_s * a / d
    np.add.at(diag_s, c, Db)
    np.add.at(src_s, c, Db * T_bot)

    # ------------------------------------------------------------------ interface
    fc_f, a_f, d_f = fluid.side_y(1, top=False)  # fluid block 2 bottom ('interface')
    fc_s, a_s, d_s = solid.side_y(0, top=True)   # solid top
    assert fc_f.size == fc_s.size
    kOwn = k_f / d_f
    kNei = k_s / d_s
    cond = a_f * kOwn * kNei / (kOwn + kNei)
    np.add.at(diag_f, fc_f, cond)                # addBoundaryDiag(internalCoeffs)
    np.add.at(diag_s, fc_s, cond)

    reg_f = Region("fluid", fluid.nCells, l, u, diag_f, upper_f, lower_f, src_f,
                   np.full(fluid.nCells, T_f0))
    reg_s = Region("solid", solid.nCells, solid.lowerAddr, solid.upperAddr, diag_s, upper_s, None,
                   src_s, np.full(solid.nCells, T_s0))
    reg_f.interfaces.append(Interface(REGION_COUPLE, fc_f, cond.copy(), cond.copy(), 0, 1, 0, name="interface"))
    reg_s.interfaces.append(Interface(REGION_COUPLE, fc_s, cond.copy(), cond.copy(), 0, 0, 0, name="top"))
    return Case(name, [RankSystem(0, 1, [reg_f, reg_s])])


def cht_fv_tables(fluid: StructuredRegion, solid: StructuredRegion, *, dt: float = 1e-2, Ux: float = 1.0):
    """What a foam-extend adapter hands to ``b200_sys_set_fv_geometry`` / ``b200_sys_assemble_T`` for the two T regions of
    the CHT case (same physics and constants as :func:`assemble_cht`): per region a dict with the static geometry
    (``V``, ``magSf``, ``deltaCoeffs``), the boundary faces in patch order (``bCells`` with their internalCoeffs ``bInt``
    and boundary-source contributions ``bSrc``; the regionCouple patch contributes its internalCoeffs), the equation form
    and constants, the face flux ``phi`` (fluid), the interface coefficients and the initial field."""
    from .ldu import TEQN_CONDUCT, TEQN_TRANSPORT
    rho_f, cp_f, k_f = 1.0, 250.0, 5.0
    rho_s, cv_s, k_s = 1.0, 100.0, 100.0
    T_f0, T_s0, T_in, T_bot = 300.0, 310.0, 300.0, 310.0
    rc_f = rho_f * cp_f
    fc_f, a_f, d_f = fluid.side_y(1, top=False)
    fc_s, a_s, d_s = solid.side_y(0, top=True)
    kOwn, kNei = k_f / d_f, k_s / d_s
    cond = a_f * kOwn * kNei / (kOwn + kNei)
    # fluid: inlet (fixedValue), outlet (zeroGradient), interface (regionCouple)
    ci, ai, di = fluid.side_xmin()
    co, ao, _ = fluid.side_xmax()
    Db = k_f * ai / di
    Fb = -rc_f * Ux * ai
    bCells_f = np.concatenate([ci, co, fc_f]).astype(np.int32)
    bInt_f = np.concatenate([Db + np.maximum(Fb, 0.0), np.maximum(rc_f * Ux * ao, 0.0), cond])
    bSrc_f = np.concatenate([(Db - np.minimum(Fb, 0.0)) * T_in, np.zeros(co.size), np.zeros(fc_f.size)])
    phi = np.where(fluid.faceDir == 0, Ux * fluid.faceArea, 0.0)  # volumetric flux U.S_f; rho*cp scales the operator
    # solid: bottom (fixedValue), top (regionCouple)
    cb, ab, db = solid.side_y(0, top=False)
    Dbs = k_s * ab / db
    bCells_s = np.concatenate([cb, fc_s]).astype(np.int32)
    bInt_s = np.concatenate([Dbs, cond])
    bSrc_s = np.concatenate([Dbs * T_bot, np.zeros(fc_s.size)])
    tf = dict(form=TEQN_TRANSPORT, rhoC=rc_f, kappa=k_f, rDeltaT=1.0 / dt, V=fluid.volume, magSf=fluid.faceArea,
              deltaCoeffs=1.0 / fluid.faceDelta, phi=phi, bCells=bCells_f, bInt=bInt_f, bSrc=bSrc_f,
              ifaceCoeffs=cond.copy(), T0=np.full(fluid.nCells, T_f0))
    ts = dict(form=TEQN_CONDUCT, rhoC=rho_s * cv_s, kappa=k_s, rDeltaT=1.0 / dt, V=solid.volume, magSf=solid.faceArea,
              deltaCoeffs=1.0 / solid.faceDelta, phi=None, bCells=bCells_s, bInt=bInt_s, bSrc=bSrc_s,
              ifaceCoeffs=cond.copy(), T0=np.full(solid.nCells, T_s0))
    return tf, ts


def cht_fv_tables_slab(r: int, layers_per_rank: int, rank: int, nranks: int):
    """:func:`cht_fv_tables` for rank ``rank`` of the z-slab decomposition :func:`cht_rank_slab` builds: the processor
    patches (after the regionCouple patch, ordered by neighbour rank) add their internalCoeffs to the boundary-face
    list, exactly what ``processorFvPatchField`` contributes through ``addBoundaryDiag``; they carry no source."""
    fluid, solid = flow_over_heated_plate(r, layers_per_rank)
    tables = cht_fv_tables(fluid, solid)
    for t, mesh in zip(tables, (fluid, solid)):
        for nb in (rank - 1, rank + 1):
            if 0 <= nb < nranks:
                c, a, d = mesh.side_z(top=(nb > rank))
                D = t["kappa"] * a / (2.0 * d)
                t["bCells"] = np.concatenate([t["bCells"], c]).astype(np.int32)
                t["bInt"] = np.concatenate([t["bInt"], D])
                t["bSrc"] = np.concatenate([t["bSrc"], np.zeros(c.size)])
    return tables, (fluid, solid)


def cht_case(r: int = 1, layers: int = 1, z1: float = 0.4, **kw) -> Tuple[Case, StructuredRegion, StructuredRegion]:
    fluid, solid = flow_over_heated_plate(r, layers, z1)
    return assemble_cht(fluid, solid, name=f"cht_r{r}_L{layers}", **kw), fluid, solid


def cht_rank_slab(r: int, layers_per_rank: int, rank: int, nranks: int) -> RankSystem:
    """Rank ``rank`` of the CHT case with ``layers_per_rank * nranks`` z-layers decomposed into z-slabs
    (``simple; n (1 1 nranks)``), assembled directly from the local slab: identical to
    ``decompose_cht_zslabs(cht_case(r, L*nranks, z1=0.4*nranks), ...).ranks[rank]`` (tests/test_decompose.py)
    without ever building the global mesh.  Every rank holds the same number of cells (weak scaling);
    processor patches (rank-1, rank+1) follow the regionCouple patch, ordered by neighbour rank."""
    fluid, solid = flow_over_heated_plate(r, layers_per_rank)
    rs = assemble_cht(fluid, solid, name=f"cht_r{r}_L{layers_per_rank}x{nranks}").ranks[0]
    rs.rank, rs.nRanks = rank, nranks
    k_of = {0: 5.0, 1: 100.0}
    for ri, (reg, mesh) in enumerate(zip(rs.regions, (fluid, solid))):
        for itf in reg.interfaces:
            itf.peerRank = rank
        nbrs = [nb for nb in (rank - 1, rank + 1) if 0 <= nb < nranks]
        base = len(reg.interfaces)
        for k, nb in enumerate(nbrs):
            c, a, d = mesh.side_z(top=(nb > rank))
            D = k_of[ri] * a / (2.0 * d)
            np.add.at(reg.diag, c, D)
            # the peer's patch list: [regionCouple, (its lower neighbour), (its upper neighbour)]
            peerNbrs = [q for q in (nb - 1, nb + 1) if 0 <= q < nranks]
            peerIface = base + peerNbrs.index(rank)
            reg.interfaces.append(Interface(PROCESSOR, c, D.copy(), D.copy(), nb, ri, peerIface,
                                            name=f"procBoundary{rank}to{nb}"))
    return rs


WORKLOADS = {
    # BASELINE.json configs; (r, layers)
    "C1": (1, 1),        # as-shipped flowOverHeatedPlate mesh, 21 812 cells
    "C2": (3, 22),       # 3-D, 4 318 776 cells   (metric config)
    "C2-2D": (14, 1),    # 2-D-faithful, 4 275 152 cells
    "C3": (4, 184),      # 3-D, 64 214 528 cells (184 = 8 x 23 layers: the case bench.py splits over 1 / 2 / 4 / 8 GPUs)
    "C3-2D": (54, 1),
    "C3-slab8": (4, 23), # one of the 8 z-slabs of C3 (r = 4, 8 x 23 = 184 layers: 64 234 496 cells on 8 ranks), 8 029 312 cells per rank
}


def synthetic_coeffs(nCells: int, l: np.ndarray, u: np.ndarray, *, symmetric: bool, seed: int = 12345,
                     name: str = "addr") -> Region:
    """Addressing-only fixture: seeded diagonally dominant coefficients on a
    given LDU addressing; b = A x*, x0 = 300."""
    rng = np.random.default_rng(seed)
    F = l.size
    upper = -(0.5 + rng.random(F))
    lower = None if symmetric else upper * (1.0 + 0.2 * (2.0 * rng.random(F) - 1.0))
    lo = upper if lower is None else lower
    diag = 0.01 - np.bincount(l, weights=upper, minlength=nCells) - np.bincount(u, weights=lo, minlength=nCells)
    xstar = 300.0 + 10.0 * rng.random(nCells)
    b = diag * xstar
    b += np.bincount(l, weights=upper * xstar[u], minlength=nCells)
    b += np.bincount(u, weights=lo * xstar[l], minlength=nCells)
    return Region(name, nCells, l.astype(np.int32), u.astype(np.int32), diag, upper, lower, b,
                  np.full(nCells, 300.0))


def single_region_case(reg: Region, name: str = "single") -> Case:
    return Case(name, [RankSystem(0, 1, [reg])])


def pu_block_matrix(nCells: int, l: np.ndarray, u: np.ndarray, *, seed: int = 2024, coupling: float = 0.05,
                    name: str = "Up"):
    """Synthetic pressure-velocity block system (BASELINE config 5) on a given LDU addressing, assembled through the
    ``fvBlockMatrix<vector4>`` mirror the way ``pUCoupledIcoFluid::setCoupledEqns`` does
    (/root/reference/src/regions/pUCoupledIcoFluid/pUCoupledIcoFluid.C:584-621): ``insertEquation(0, UEqn)``
    (asymmetric momentum matrix shared by the three components), ``insertEquation(3, pEqn)`` (symmetric pressure
    Laplacian), ``insertBlockCoupling(0, 3, grad p, true)`` and ``insertBlockCoupling(3, 0, div U, false)``: 10 of the
    16 entries of every block are structurally non-zero.  Seeded, diagonally dominant; the source is
    ``A x*`` for a known ``x*`` (returned as ``M.xstar``) and the initial guess is zero."""
    from .blockldu import BlockCoupling, ScalarEqn, fvBlockMatrix
    rng = np.random.default_rng(seed)
    F = l.size
    l = np.ascontiguousarray(l, np.int32)
    u = np.ascontiguousarray(u, np.int32)
    upU = -(0.5 + rng.random(F))
    loU = upU * (1.0 + 0.2 * (2.0 * rng.random(F) - 1.0))
    dU = 0.5 - np.bincount(l, weights=upU, minlength=nCells) - np.bincount(u, weights=loU, minlength=nCells)
    upP = -(0.1 + 0.2 * rng.random(F))
    dP = 0.02 - np.bincount(l, weights=upP, minlength=nCells) - np.bincount(u, weights=upP, minlength=nCells)
    Sf = coupling * rng.standard_normal((F, 3))
    w = 0.3 + 0.4 * rng.random(F)
    # fvm::grad(p): owner gets +w Sf p_P + (1-w) Sf p_N, neighbour the opposite sign
    gUp = (1.0 - w)[:, None] * Sf
    gLo = -w[:, None] * Sf
    gD = np.zeros((nCells, 3))
    for c in range(3):
        gD[:, c] = np.bincount(l, weights=w * Sf[:, c], minlength=nCells) - np.bincount(u, weights=(1.0 - w) * Sf[:, c], minlength=nCells)
    # fvm::UDiv(U): the same stencil acting on U, feeding the p row
    M = fvBlockMatrix(l, u, nCells, name=name)
    M.insertEquation(0, ScalarEqn(dU, np.zeros((nCells, 3)), upU, loU), nCmpts=3)
    M.insertEquation(3, ScalarEqn(dP, np.zeros(nCells), upP, None))
    M.insertBlockCoupling(0, 3, BlockCoupling(gD, gUp, gLo), True)
    M.insertBlockCoupling(3, 0, BlockCoupling(gD, gUp, gLo), False)
    xstar = np.empty((nCells, 4))
    xstar[:, :3] = 1.0 + 0.1 * rng.standard_normal((nCells, 3))
    xstar[:, 3] = 10.0 + rng.standard_normal(nCells)
    b = np.einsum("nij,nj->ni", M.diag, xstar)
    np.add.at(b, u, np.einsum("fij,fj->fi", M.lower, xstar[l]))
    np.add.at(b, l, np.einsum("fij,fj->fi", M.upper, xstar[u]))
    M.source[...] = b
    M.xstar = xstar
    return M
