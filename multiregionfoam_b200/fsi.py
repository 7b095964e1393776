"""BASELINE config 4: partitioned fluid-structure coupling on the HronTurekFsi3 topology.

The reference runs this case with PARTITIONED coupling only ("monolithic coupling for fsi cases currently not supported",
``tutorials/fluidStructureInteraction/HronTurekFsi3/Allrun:53-59``): inside every Dirichlet-Neumann iteration
(``multiRegionSystem::solve`` ``src/multiRegionSystem/multiRegionSystem.C:572-614``) the fluid region solves its segregated
equations, ``assembleAndSolveEqns`` (``:193-324``, ``eqn.solve()`` at ``:293``) solves the solid's displacement equation, and
between the solves the interface fields cross the GGI (``regionInterfaceType::interpolateFacesFromA/B``,
``src/regionInterfaces/regionInterface/regionInterfaceTypeTemplates.C:35-133`` ->
``ggiInterfaceToInterfaceMapping::transferFacesZoneToZone``).  Solver selection as shipped
(``HronTurekFsi3/system/fluid/fvSolution:17-70``, ``system/solid/fvSolution:17-28``):

    U   PBiCG + DILU      (vector: three component solves sharing the off-diagonals)
    p   GAMG              -> PCG + DIC here (GAMG is out of scope)
    D   PCG + FDIC (DIC)  (vector: three component solves)

One *coupling iteration* of this module is that sequence with fixed iteration counts, the traction going fluid -> solid
and the displacement solid -> fluid through ``b200_ggi_interpolate`` with the weights of the non-conformal flag
interface (84 r L fluid faces against 216 r L solid faces).  The matrices are seeded, diagonally dominant
addressing-only fixtures on the generated addressing: the hot path does not care where coefficients come from.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Tuple

import numpy as np

from .assembly import synthetic_coeffs
from .blockmesh2d import HT_FLUID_INTERFACE, HT_SOLID_INTERFACE, hron_turek, hron_turek_interface_intervals, interval_ggi
from .case import Region

FSI_WORKLOADS = {"C4": (7, 27), "C4-2D": (37, 1), "C4-mini": (1, 2)}   # (refinement, z-layers): 7.89 M / 8.17 M / 11 932 cells
N_U, N_P, N_D = 5, 30, 20   # Krylov iterations per component solve of U, of p, per component solve of D


@dataclass
class FsiCase:
    fluidU: Region     # momentum matrix (asymmetric); the three components share it
    fluidP: Region     # pressure matrix (symmetric)
    solidD: Region     # displacement matrix (symmetric)
    bU: np.ndarray     # [3, Nf] sources of the momentum components
    bD: np.ndarray     # [3, Ns]
    fluidFaceCells: np.ndarray
    solidFaceCells: np.ndarray
    toSolid: Tuple[np.ndarray, np.ndarray, np.ndarray]   # GGI tables fluid faces -> solid faces
    toFluid: Tuple[np.ndarray, np.ndarray, np.ndarray]
    r: int
    layers: int

    @property
    def nFluid(self) -> int:
        return self.fluidU.nCells

    @property
    def nSolid(self) -> int:
        return self.solidD.nCells

    def cell_iterations(self) -> int:
        return self.nFluid * (3 * N_U + N_P) + self.nSolid * 3 * N_D


def fsi_case(r: int, layers: int) -> FsiCase:
    fluid, solid = hron_turek(r, layers)
    U = synthetic_coeffs(fluid.nCells, fluid.lowerAddr, fluid.upperAddr, symmetric=False, seed=4101, name="fluidU")
    P = synthetic_coeffs(fluid.nCells, fluid.lowerAddr, fluid.upperAddr, symmetric=True, seed=4102, name="fluidP")
    D = synthetic_coeffs(solid.nCells, solid.lowerAddr, solid.upperAddr, symmetric=True, seed=4103, name="solidD")
    rng = np.random.default_rng(4104)
    bU = np.stack([U.source * (1.0 + 0.1 * c) + rng.standard_normal(U.nCells) for c in range(3)])
    bD = np.stack([D.source * (1.0 - 0.1 * c) + rng.standard_normal(D.nCells) for c in range(3)])
    tf, ts = hron_turek_interface_intervals(fluid, solid)
    return FsiCase(U, P, D, bU, bD, fluid.patch_cells(HT_FLUID_INTERFACE), solid.patch_cells(HT_SOLID_INTERFACE),
                   interval_ggi(ts, tf), interval_ggi(tf, ts), r, layers)


def coupling_iteration(case: FsiCase, solve: Callable, transfer: Callable, state: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """One Dirichlet-Neumann iteration.  ``solve(key, x0, b, solver, precond, iters) -> x`` with key in {"U", "p", "D"},
    ``transfer(tables, field[nFrom, 3]) -> field[nTo, 3]``; ``state`` holds U [3, Nf], p [Nf], D [3, Ns] and is updated.
    The interface data flow is the reference's: the fluid's interface values (here: its momentum solution next to the
    flag, standing in for the traction sigma) load the solid's interface cells (regionCoupledTraction,
    ``HronTurekFsi3/0/solid/orig/partitioned/D``), the solid's interface displacement feeds back into the fluid's sources."""
    fc, sc = case.fluidFaceCells, case.solidFaceCells
    dFaces = transfer(case.toFluid, np.ascontiguousarray(state["D"][:, sc].T))            # solid D -> fluid faces
    for c in range(3):
        b = case.bU[c].copy()
        np.add.at(b, fc, 1e-3 * dFaces[:, c])
        state["U"][c] = solve("U", state["U"][c], b, "PBiCG", "DILU", N_U)
    bp = case.fluidP.source + 1e-3 * (state["U"][0] - state["U"][1])
    state["p"] = solve("p", state["p"], bp, "PCG", "DIC", N_P)
    traction = transfer(case.toSolid, np.ascontiguousarray(state["U"][:, fc].T))          # fluid -> solid faces
    for c in range(3):
        b = case.bD[c].copy()
        np.add.at(b, sc, 1e-3 * traction[:, c])
        state["D"][c] = solve("D", state["D"][c], b, "PCG", "DIC", N_D)
    return state


def initial_state(case: FsiCase) -> Dict[str, np.ndarray]:
    return {"U": np.stack([case.fluidU.psi.copy() for _ in range(3)]), "p": case.fluidP.psi.copy(),
            "D": np.stack([case.solidD.psi.copy() for _ in range(3)])}


class DeviceFsi:
    """The three LDU systems of the case on the device and the two callbacks of ``coupling_iteration`` through the C ABI."""

    def __init__(self, ctx, case: FsiCase):
        from . import ldu
        from .case import RankSystem
        self.ctx, self.case, self.ldu = ctx, case, ldu
        self.sys = {k: ldu.LduSystem(ctx, RankSystem(0, 1, [reg])) for k, reg in (("U", case.fluidU), ("p", case.fluidP), ("D", case.solidD))}
        self.device_ms = 0.0
        self.iterations = 0
        self.histories = []

    def solve(self, key, x0, b, solver, precond, iters):
        ldu = self.ldu
        S = self.sys[key]
        sid = {"PBiCG": ldu.SOLVER_PBICG, "PCG": ldu.SOLVER_PCG, "BiCGStab": ldu.SOLVER_BICGSTAB}[solver]
        pid = {"DILU": ldu.PRECOND_DILU, "DIC": ldu.PRECOND_DIC}[precond]
        x, info = S.solve(x0, b, sid, pid, tolerance=0.0, minIter=iters, maxIter=iters)
        self.device_ms += info["deviceMs"]
        self.iterations += info["nIterations"]
        self.histories.append(info["history"])
        return x

    def transfer(self, tables, field):
        return self.ctx.ggi_interpolate(tables[0], tables[1], tables[2], field)

    def close(self):
        for S in self.sys.values():
            S.close()
