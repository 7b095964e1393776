"""Host-side mirror of the fvSolution / run-time selection interface (no GPU needed)."""
import pytest

from multiregionfoam_b200 import ldu, solvers

# the entries of solvers/Tcoupled as shipped with the CHT tutorials (sub-dictionary preconditioner form),
# a word-form entry and a regex key, written out here (the reference tree is not read at test time)
FVSOLUTION = r"""
FoamFile { version 2.0; format ascii; class dictionary; object fvSolution; }
solvers
{
    // monolithic temperature
    Tcoupled
    {
        solver          BiCGStab;
        preconditioner
        {
            preconditioner Cholesky;
        }
        tolerance       1e-15;
        relTol          0;
        minIter         0;
        maxIter         200;
    }
    "(U|k)Final" { solver cudaPBiCGStab; preconditioner cudaDILU; tolerance 1e-09; /* relTol default */ }
    p { solver cudaPCG; preconditioner DIC; relTol 0.01; }
}
"""


def test_parse_and_lookup():
    d = solvers.parse_dictionary(FVSOLUTION)
    t = solvers.read_controls(solvers.lookup_solver_dict(d, "Tcoupled"))
    assert t == dict(solver="BiCGStab", preconditioner="Cholesky", tolerance=1e-15, relTol=0.0, minIter=0, maxIter=200)
    u = solvers.read_controls(solvers.lookup_solver_dict(d, "UFinal"))
    assert u["solver"] == "cudaPBiCGStab" and u["preconditioner"] == "cudaDILU" and u["tolerance"] == 1e-9
    assert u["relTol"] == 0.0 and u["minIter"] == 0 and u["maxIter"] == 1000  # readControls defaults
    p = solvers.read_controls(solvers.lookup_solver_dict(d, "p"))
    assert p["tolerance"] == 1e-6 and p["relTol"] == 0.01
    with pytest.raises(solvers.FatalError):
        solvers.lookup_solver_dict(d, "Tfluid")


def test_selection_tables():
    assert solvers.SOLVER_TABLE["cudaPBiCGStab"][0] == solvers.SOLVER_TABLE["BiCGStab"][0] == ldu.SOLVER_BICGSTAB
    assert solvers.SOLVER_TABLE["cudaPCG"][0] == solvers.SOLVER_TABLE["PCG"][0] == ldu.SOLVER_PCG
    assert solvers.PRECOND_TABLE["cudaDILU"] == solvers.PRECOND_TABLE["DILU"] == ldu.PRECOND_DILU
    assert solvers.PRECOND_TABLE["cudaDIC"] == solvers.PRECOND_TABLE["FDIC"] == ldu.PRECOND_DIC


def test_performance_print_format():
    p = solvers.lduSolverPerformance("BiCGStab", "T", 0.5, 1e-16, 12)
    assert p.line() == "BiCGStab:  Solving for T, Initial residual = 0.5, Final residual = 1e-16, No Iterations 12"
