/*---------------------------------------------------------------------------*\
  cudaLduSolver.C -- see cudaLduSolver.H.  NOT compiled in this repository.
\*---------------------------------------------------------------------------*/
#include "cudaLduSolver.H"
#include "b200Binding.H"
#include "addToRunTimeSelectionTable.H"

namespace Foam
{
    template<> const word cudaPCG::typeName("cudaPCG");
    template<> int cudaPCG::debug(0);
    template<> const word cudaPBiCGStab::typeName("cudaPBiCGStab");
    template<> int cudaPBiCGStab::debug(0);

    // constructor tables of lduMatrix::solver (lduMatrix.H: declareRunTimeSelectionTable ... symMatrix / asymMatrix)
    lduMatrix::solver::addsymMatrixConstructorToTable<cudaPCG> addcudaPCGSymMatrixConstructorToTable_;
    lduMatrix::solver::addsymMatrixConstructorToTable<cudaPBiCGStab> addcudaPBiCGStabSymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<cudaPBiCGStab> addcudaPBiCGStabAsymMatrixConstructorToTable_;
}


template<int SolverId>
Foam::cudaLduSolver<SolverId>::cudaLduSolver
(
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& coupleBouCoeffs,
    const FieldField<Field, scalar>& coupleIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const dictionary& dict
)
:
    lduMatrix::solver(fieldName, matrix, coupleBouCoeffs, coupleIntCoeffs, interfaces, dict)
{
    readControls();
}


template<int SolverId>
Foam::lduSolverPerformance Foam::cudaLduSolver<SolverId>::solve
(
    scalarField& x,
    const scalarField& b,
    const direction cmpt
) const
{
    lduSolverPerformance solverPerf(typeName, fieldName());

    UPtrList<const lduMatrix> matrices(1);
    matrices.set(0, &matrix_);
    List<lduInterfaceFieldPtrsList> ifaces(1, interfaces_);
    b200_sys* sys = b200Binding::system(matrices, ifaces);           // cached on the addressing

    List<const FieldField<Field, scalar>*> bou(1, &coupleBouCoeffs_), inte(1, &coupleIntCoeffs_);
    b200Binding::setCoeffs(sys, matrices, bou, inte);                // b200_sys_set_coeffs / _set_interface_coeffs

    b200_solver_opts opts;
    opts.solver = SolverId;
    opts.precond = b200Binding::precondId(dict());
    opts.tolerance = tolerance();
    opts.relTol = relTolerance();
    opts.minIter = minIter();
    opts.maxIter = maxIter();

    double* xp[1] = { x.begin() };
    const double* bp[1] = { b.begin() };
    b200_perf perf;
    b200Binding::check
    (
        b200_solve(sys, &opts, xp, bp, &perf, NULL, 0),
        "cudaLduSolver::solve(scalarField&, const scalarField&, const direction) const"
    );

    solverPerf.initialResidual() = perf.initialResidual;
    solverPerf.finalResidual() = perf.finalResidual;
    solverPerf.nIterations() = perf.nIterations;
    solverPerf.converged() = perf.converged;
    solverPerf.singular() = perf.singular;
    return solverPerf;
}

template class Foam::cudaLduSolver<0>;
template class Foam::cudaLduSolver<1>;
