"""Host-side mirror of the solver-selection interface the reference reaches the hot path through.

In the reference the linear solver is chosen by an ``fvSolution`` sub-dictionary and created from
a run-time selection table:

    coupledFvMatrix<T>::solve(dict "<fld>coupled")   /root/reference/src/multiRegionSystem/multiRegionSystem.C:150-153
    fvMatrix<T>::solve()                              /root/reference/src/multiRegionSystem/multiRegionSystem.C:293
    e.g. solvers { Tcoupled { solver BiCGStab; preconditioner { preconditioner Cholesky; }
                              tolerance 1e-15; relTol 0; minIter 0; maxIter 200; } }
         /root/reference/tutorials/conjugateHeatTransfer/flowOverHeatedPlate/system/fluid/fvSolution:43-57

This module keeps the same vocabulary: a table of solver type names (the reference's own names and
the ``cuda*`` names the drop-in library registers), a table of preconditioner names, the
dictionary defaults of ``lduMatrix::solver::readControls`` (tolerance 1e-6, relTol 0, minIter 0,
maxIter 1000) and an ``lduSolverPerformance``-like result with the reference's print format.
Unknown names raise, like ``FatalIOError`` from the selection table ("Unknown ... solver").
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Union

import numpy as np

from . import ldu
from .case import RankSystem

# name -> (solver id, symmetric-only)
SOLVER_TABLE: Dict[str, tuple] = {
    # names registered by the drop-in library (north_star)
    "cudaPCG": (ldu.SOLVER_PCG, True),
    "cudaPBiCGStab": (ldu.SOLVER_BICGSTAB, False),
    "cudaPBiCG": (ldu.SOLVER_PBICG, False),
    # names used by the shipped dictionaries; the adapter registers aliases so existing cases run unchanged
    "PCG": (ldu.SOLVER_PCG, True),
    "CG": (ldu.SOLVER_PCG, True),
    "BiCGStab": (ldu.SOLVER_BICGSTAB, False),
    "PBiCGStab": (ldu.SOLVER_BICGSTAB, False),
    "PBiCG": (ldu.SOLVER_PBICG, False),
    "BiCG": (ldu.SOLVER_PBICG, False),
}

PRECOND_TABLE: Dict[str, int] = {
    "cudaDIC": ldu.PRECOND_DIC,
    "cudaDILU": ldu.PRECOND_DILU,
    "DIC": ldu.PRECOND_DIC,
    "FDIC": ldu.PRECOND_DIC,   # identical products, pre-multiplied
    "DILU": ldu.PRECOND_DILU,
    "Cholesky": ldu.PRECOND_CHOLESKY,
    "diagonal": ldu.PRECOND_DIAGONAL,
    "none": ldu.PRECOND_NONE,
}


class FatalError(RuntimeError):
    """Mirror of ``FatalErrorIn(...) << ... << abort(FatalError)``."""


@dataclass
class lduSolverPerformance:
    solverName: str
    fieldName: str
    initialResidual: float = 0.0
    finalResidual: float = 0.0
    nIterations: int = 0
    converged: bool = False
    singular: bool = False
    normFactor: float = 0.0
    deviceMs: float = 0.0
    history: Optional[np.ndarray] = None

    def line(self) -> str:
        # lduSolverPerformance::print
        return (f"{self.solverName}:  Solving for {self.fieldName}, Initial residual = {self.initialResidual:g}, "
                f"Final residual = {self.finalResidual:g}, No Iterations {self.nIterations}")


# --------------------------------------------------------------------------- dictionaries

def parse_dictionary(text: str) -> dict:
    """Minimal OpenFOAM dictionary reader (nested ``{}``, ``key value;``, quoted regex keys,
    C/C++ comments) -- enough for ``fvSolution``."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    toks = re.findall(r'"[^"]*"|[{};]|[^\s{};]+', text)
    pos = 0

    def conv(v: str):
        try:
            return int(v)
        except ValueError:
            pass
        try:
            return float(v)
        except ValueError:
            return v

    def block() -> dict:
        nonlocal pos
        d: dict = {}
        while pos < len(toks):
            t = toks[pos]
            if t == "}":
                pos += 1
                return d
            key = t.strip('"')
            pos += 1
            if pos < len(toks) and toks[pos] == "{":
                pos += 1
                d[key] = block()
                continue
            vals: List[str] = []
            while pos < len(toks) and toks[pos] != ";":
                if toks[pos] in "{}":
                    raise FatalError(f"malformed dictionary entry '{key}'")
                vals.append(toks[pos])
                pos += 1
            pos += 1  # ;
            d[key] = conv(vals[0]) if len(vals) == 1 else [conv(v) for v in vals]
        return d

    return block()


def lookup_solver_dict(fvSolution: dict, name: str) -> dict:
    """``solutionDict().solver(name)``: exact key first, then regex keys ("(U|p)Final")."""
    solvers = fvSolution.get("solvers", fvSolution)
    if name in solvers:
        return solvers[name]
    for k, v in solvers.items():
        try:
            if isinstance(v, dict) and re.fullmatch(k, name):
                return v
        except re.error:
            continue
    raise FatalError(f"keyword {name} is undefined in dictionary solvers")


def read_controls(d: dict) -> dict:
    """lduMatrix::solver::readControls defaults."""
    pre = d.get("preconditioner", "none")
    if isinstance(pre, dict):  # "preconditioner { preconditioner Cholesky; }" form
        pre = pre.get("preconditioner", pre.get("type"))
        if pre is None:
            raise FatalError("preconditioner sub-dictionary lacks the 'preconditioner' entry")
    return dict(solver=d.get("solver", d.get("type")), preconditioner=pre, tolerance=float(d.get("tolerance", 1e-6)),
                relTol=float(d.get("relTol", 0.0)), minIter=int(d.get("minIter", 0)), maxIter=int(d.get("maxIter", 1000)))


# --------------------------------------------------------------------------- the solver object

class coupledLduSolver:
    """``coupledLduSolver::New(fieldName, matrix, bouCoeffs, intCoeffs, interfaces, dict)->solve(x, b)``
    on one rank.  ``matrix`` is the device system (:class:`multiregionfoam_b200.ldu.LduSystem`)."""

    def __init__(self, fieldName: str, system: ldu.LduSystem, solverDict: dict):
        c = read_controls(solverDict)
        name = c["solver"]
        if name is None:
            raise FatalError("solver dictionary lacks the 'solver' entry")
        if name not in SOLVER_TABLE:
            raise FatalError(f"Unknown coupled matrix solver {name}; valid solvers are {sorted(SOLVER_TABLE)}")
        if c["preconditioner"] not in PRECOND_TABLE:
            raise FatalError(f"Unknown preconditioner {c['preconditioner']}; valid preconditioners are {sorted(PRECOND_TABLE)}")
        self.solverName, self.fieldName, self.system, self.controls = name, fieldName, system, c
        self.solverId, symOnly = SOLVER_TABLE[name]
        self.precondId = PRECOND_TABLE[c["preconditioner"]]
        if symOnly and any(not r.symmetric for r in system.rs.regions):
            # the reference looks the name up in the asymMatrix constructor table, where PCG is absent
            raise FatalError(f"Unknown asymmetric matrix solver {name}")

    @classmethod
    def New(cls, fieldName: str, system: ldu.LduSystem, solverDict: dict) -> "coupledLduSolver":
        return cls(fieldName, system, solverDict)

    def solve(self, x: np.ndarray, b: np.ndarray, history: bool = True) -> lduSolverPerformance:
        """x is updated in place (concatenated per-region fields of this rank)."""
        c = self.controls
        xs, info = self.system.solve(x, b, self.solverId, self.precondId, c["tolerance"], c["relTol"], c["minIter"],
                                     c["maxIter"], history=history)
        x[...] = xs
        return lduSolverPerformance(self.solverName, self.fieldName, info["initialResidual"], info["finalResidual"],
                                    info["nIterations"], info["converged"], info["singular"], info["normFactor"],
                                    info["deviceMs"], info.get("history"))


def solve_coupled(ctx: ldu.Context, rs: RankSystem, fieldName: str, fvSolution: Union[dict, str],
                  system: Optional[ldu.LduSystem] = None) -> lduSolverPerformance:
    """``coupledFvMatrix<scalar>::solve(solutionDict().solver(fieldName + "coupled"))`` for the regions of
    one rank: psi of every region is updated in place (multiRegionSystem.C:150-153)."""
    if isinstance(fvSolution, str):
        fvSolution = parse_dictionary(fvSolution)
    d = lookup_solver_dict(fvSolution, fieldName + "coupled")
    own = system is None
    system = system or ldu.LduSystem(ctx, rs)
    try:
        x = np.concatenate([r.psi for r in rs.regions])
        b = np.concatenate([r.source for r in rs.regions])
        perf = coupledLduSolver.New(fieldName, system, d).solve(x, b)
        o = 0
        for r in rs.regions:
            r.psi[...] = x[o:o + r.nCells]
            o += r.nCells
        return perf
    finally:
        if own:
            system.close()


# --------------------------------------------------------------------------- smoothSolver

class smoothSolver:
    """``solver smoothSolver; smoother <name>; nSweeps n;`` on the coupled device system of one rank
    (foam-extend smoothSolver.C: smooth nSweeps times, check ``gSumMag(residual) / normFactor``, repeat).

    DIC / DILU smoothers run through ``LduSystem.smooth`` (b200_smooth: the preconditioner's sweeps applied to the residual),
    GaussSeidel through :class:`multiregionfoam_b200.smoother.GaussSeidel` (one region without coupled patches), DICGaussSeidel
    = both in turn.  Residuals use ``LduSystem.residual`` (lduMatrix::residual) and ``LduSystem.amul``; the norms are summed on
    the host in this mirror (the foam-extend adapters leave them to smoothSolver itself)."""

    def __init__(self, fieldName: str, system: ldu.LduSystem, solverDict: dict):
        from . import smoother as sm
        name = solverDict.get("smoother")
        if name not in sm.SMOOTHER_TABLE:
            raise FatalError(f"Unknown smoother {name}; valid smoothers are {sorted(sm.SMOOTHER_TABLE)}")
        self.kind = sm.SMOOTHER_TABLE[name]
        self.fieldName, self.system = fieldName, system
        self.nSweeps = int(solverDict.get("nSweeps", 1))
        c = read_controls(dict(solverDict, solver="smoothSolver"))
        self.controls = c
        regs = system.rs.regions
        sym = all(r.symmetric for r in regs)
        if self.kind in ("dic", "dic+gs") and not sym:
            raise FatalError(f"Unknown asymmetric matrix smoother {name}")
        if self.kind == "dilu" and sym:
            raise FatalError(f"Unknown symmetric matrix smoother {name}")
        self.gs = None
        if "gs" in self.kind:
            if len(regs) != 1 or regs[0].interfaces:
                raise FatalError("the GaussSeidel smoother of this mirror serves one region without coupled patches "
                                 "(with coupled patches: adapters/b200LduSolvers/cudaGaussSeidelSmoother.C builds bPrime on the host)")
            r = regs[0]
            self.gs = sm.GaussSeidel(system.ctx, r.lowerAddr, r.upperAddr, r.nCells)
            self.gs.set_coeffs(r.diag, r.upper, r.lower)

    @classmethod
    def New(cls, fieldName: str, system: ldu.LduSystem, solverDict: dict) -> "smoothSolver":
        return cls(fieldName, system, solverDict)

    def close(self):
        if self.gs is not None:
            self.gs.close()
            self.gs = None

    def _smooth(self, x: np.ndarray, b: np.ndarray, n: int) -> np.ndarray:
        if self.kind in ("dic", "dic+gs"):
            x = self.system.smooth(ldu.PRECOND_DIC, x, b, n)
        elif self.kind == "dilu":
            x = self.system.smooth(ldu.PRECOND_DILU, x, b, n)
        if "gs" in self.kind:
            x = self.gs.smooth(x, b, n)
        return x

    def solve(self, x: np.ndarray, b: np.ndarray) -> lduSolverPerformance:
        """x is updated in place."""
        c, S = self.controls, self.system
        perf = lduSolverPerformance("smoothSolver", self.fieldName)
        if self.nSweeps < 0:
            x[...] = self._smooth(x, b, -self.nSweeps)
            perf.nIterations = -self.nSweeps
            return perf
        Ax = S.amul(x)
        pA = S.amul(np.full_like(x, x.sum() / max(1, x.size)))
        perf.normFactor = float(np.abs(Ax - pA).sum() + np.abs(b - pA).sum() + 1e-20)
        perf.initialResidual = perf.finalResidual = float(np.abs(b - Ax).sum() / perf.normFactor)
        hist = [perf.initialResidual]

        def stop() -> bool:
            if perf.nIterations < c["minIter"]:
                return False
            perf.converged = bool(perf.finalResidual < c["tolerance"]
                                  or (c["relTol"] > 1e-15 and perf.finalResidual <= c["relTol"] * perf.initialResidual))
            return perf.nIterations >= c["maxIter"] or perf.converged

        if not stop():
            while True:
                x[...] = self._smooth(x, b, self.nSweeps)
                perf.finalResidual = float(np.abs(S.residual(x, b)).sum() / perf.normFactor)
                perf.nIterations += self.nSweeps
                hist.append(perf.finalResidual)
                if stop():
                    break
        perf.history = np.array(hist)
        return perf
