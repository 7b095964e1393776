/*---------------------------------------------------------------------------*\
  cudaCoupledLduSolver.C -- see cudaCoupledLduSolver.H.  NOT compiled in this repository.
\*---------------------------------------------------------------------------*/
#include "cudaCoupledLduSolver.H"
#include "b200Binding.H"
#include "addToRunTimeSelectionTable.H"

namespace Foam
{
    defineTypeNameAndDebug(cudaCoupledPBiCGStab, 0);

    coupledLduSolver::addsymMatrixConstructorToTable<cudaCoupledPBiCGStab>
        addcudaCoupledPBiCGStabSymMatrixConstructorToTable_;
    coupledLduSolver::addasymMatrixConstructorToTable<cudaCoupledPBiCGStab>
        addcudaCoupledPBiCGStabAsymMatrixConstructorToTable_;
}


Foam::cudaCoupledPBiCGStab::cudaCoupledPBiCGStab
(
    const word& fieldName,
    const coupledLduMatrix& matrix,
    const PtrList<FieldField<Field, scalar> >& bouCoeffs,
    const PtrList<FieldField<Field, scalar> >& intCoeffs,
    const lduInterfaceFieldPtrsListList& interfaces,
    const dictionary& solverData
)
:
    coupledIterativeSolver(fieldName, matrix, bouCoeffs, intCoeffs, interfaces, solverData)
{}


Foam::coupledSolverPerformance Foam::cudaCoupledPBiCGStab::solve
(
    FieldField<Field, scalar>& x,
    const FieldField<Field, scalar>& b,
    const direction cmpt
) const
{
    coupledSolverPerformance solverPerf(typeName, fieldName());

    const label nRows = matrix_.size();
    UPtrList<const lduMatrix> matrices(nRows);
    List<lduInterfaceFieldPtrsList> ifaces(nRows);
    List<const FieldField<Field, scalar>*> bou(nRows), inte(nRows);
    forAll (matrix_, rowI)
    {
        matrices.set(rowI, &matrix_[rowI]);
        ifaces[rowI] = interfaces_[rowI];
        bou[rowI] = &bouCoeffs_[rowI];
        inte[rowI] = &intCoeffs_[rowI];
    }
    b200_sys* sys = b200Binding::system(matrices, ifaces);
    b200Binding::setCoeffs(sys, matrices, bou, inte);

    b200_solver_opts opts;
    opts.solver = B200_SOLVER_BICGSTAB;
    opts.precond = b200Binding::precondId(dict());
    opts.tolerance = tolerance();
    opts.relTol = relTolerance();
    opts.minIter = minIter();
    opts.maxIter = maxIter();

    List<double*> xp(nRows);
    List<const double*> bp(nRows);
    forAll (x, rowI)
    {
        xp[rowI] = x[rowI].begin();
        bp[rowI] = b[rowI].begin();
    }
    b200_perf perf;
    b200Binding::check
    (
        b200_solve(sys, &opts, xp.begin(), bp.begin(), &perf, NULL, 0),
        "cudaCoupledPBiCGStab::solve(FieldField<Field, scalar>&, const FieldField<Field, scalar>&, const direction) const"
    );

    solverPerf.initialResidual() = perf.initialResidual;
    solverPerf.finalResidual() = perf.finalResidual;
    solverPerf.nIterations() = perf.nIterations;
    solverPerf.converged() = perf.converged;
    solverPerf.singular() = perf.singular;
    return solverPerf;
}
