/*---------------------------------------------------------------------------*\
  cudaCoupledLduSolver.C -- see cudaCoupledLduSolver.H.
\*---------------------------------------------------------------------------*/
#include "cudaCoupledLduSolver.H"
#include "b200Binding.H"
#include "addToRunTimeSelectionTable.H"
#include "Pstream.H"

namespace Foam
{
    defineTypeNameAndDebug(cudaCoupledPCG, 0);
    defineTypeNameAndDebug(cudaCoupledPBiCGStab, 0);
    defineTypeNameAndDebug(cudaCoupledPBiCG, 0);

    coupledLduSolver::addsymMatrixConstructorToTable<cudaCoupledPCG>
        addcudaCoupledPCGSymMatrixConstructorToTable_;
    coupledLduSolver::addsymMatrixConstructorToTable<cudaCoupledPBiCGStab>
        addcudaCoupledPBiCGStabSymMatrixConstructorToTable_;
    coupledLduSolver::addasymMatrixConstructorToTable<cudaCoupledPBiCGStab>
        addcudaCoupledPBiCGStabAsymMatrixConstructorToTable_;
    coupledLduSolver::addasymMatrixConstructorToTable<cudaCoupledPBiCG>
        addcudaCoupledPBiCGAsymMatrixConstructorToTable_;
}


Foam::cudaCoupledLduSolverBase::cudaCoupledLduSolverBase
(
    const int solverId,
    const word& fieldName,
    const coupledLduMatrix& matrix,
    const PtrList<FieldField<Field, scalar> >& bouCoeffs,
    const PtrList<FieldField<Field, scalar> >& intCoeffs,
    const lduInterfaceFieldPtrsListList& interfaces,
    const dictionary& solverData
)
:
    coupledIterativeSolver(fieldName, matrix, bouCoeffs, intCoeffs, interfaces, solverData),
    solverId_(solverId)
{}


Foam::coupledSolverPerformance Foam::cudaCoupledLduSolverBase::solve
(
    FieldField<Field, scalar>& x,
    const FieldField<Field, scalar>& b,
    const direction cmpt
) const
{
    coupledSolverPerformance solverPerf(solverName(), fieldName());

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
    b200Binding::systemEntry& entry = b200Binding::system(matrices, ifaces);
    b200Binding::setCoeffs(entry, matrices, bou, inte);

    b200_solver_opts opts;
    opts.solver = solverId_;
    opts.precond = b200Binding::precondId(dict());
    opts.tolerance = tolerance();
    opts.relTol = relTolerance();
    opts.minIter = minIter();
    opts.maxIter = maxIter();

    const bool dumping = dict().found("b200Dump");
    PtrList<scalarField> x0(dumping ? nRows : 0);
    scalarField history(dumping ? 64 : 0, 0.0);

    List<double*> xp(nRows);
    List<const double*> bp(nRows);
    forAll (x, rowI)
    {
        if (dumping) x0.set(rowI, new scalarField(x[rowI]));
        xp[rowI] = x[rowI].begin();
        bp[rowI] = b[rowI].begin();
    }
    b200_perf perf;
    b200Binding::check
    (
        b200_solve(entry.sys, &opts, xp.begin(), bp.begin(), &perf, dumping ? history.begin() : NULL, history.size()),
        "cudaCoupledLduSolverBase::solve(FieldField<Field, scalar>&, const FieldField<Field, scalar>&, const direction) const"
    );

    if (dumping)
    {
        history.setSize(min(history.size(), perf.nIterations + 1));
        UPtrList<const scalarField> xs(nRows), bs(nRows);
        forAll (x, rowI)
        {
            xs.set(rowI, &x0[rowI]);
            bs.set(rowI, &b[rowI]);
        }
        const fileName dir(dict().lookup("b200Dump"));
        b200Binding::dump
        (
            dir/(fieldName() + "_proc" + name(Pstream::myProcNo()) + ".b200ldu"), matrices, ifaces, bou, inte, xs, bs,
            solverName(), b200Binding::precondName(dict()), tolerance(), relTolerance(), minIter(), maxIter(), history
        );
    }

    solverPerf.initialResidual() = perf.initialResidual;
    solverPerf.finalResidual() = perf.finalResidual;
    solverPerf.nIterations() = perf.nIterations;
    solverPerf.converged() = perf.converged;
    solverPerf.singular() = perf.singular;
    return solverPerf;
}
