/*---------------------------------------------------------------------------*\
  cudaLduSolver.C -- see cudaLduSolver.H.
\*---------------------------------------------------------------------------*/
#include "cudaLduSolver.H"
#include "b200Binding.H"
#include "addToRunTimeSelectionTable.H"
#include "Pstream.H"

namespace Foam
{
    defineTypeNameAndDebug(cudaPCG, 0);
    defineTypeNameAndDebug(cudaPBiCGStab, 0);
    defineTypeNameAndDebug(cudaPBiCG, 0);

    // constructor tables of lduMatrix::solver (lduMatrix.H: declareRunTimeSelectionTable ... symMatrix / asymMatrix)
    lduMatrix::solver::addsymMatrixConstructorToTable<cudaPCG> addcudaPCGSymMatrixConstructorToTable_;
    lduMatrix::solver::addsymMatrixConstructorToTable<cudaPBiCGStab> addcudaPBiCGStabSymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<cudaPBiCGStab> addcudaPBiCGStabAsymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<cudaPBiCG> addcudaPBiCGAsymMatrixConstructorToTable_;
}


Foam::cudaLduSolverBase::cudaLduSolverBase
(
    const int solverId,
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& coupleBouCoeffs,
    const FieldField<Field, scalar>& coupleIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const dictionary& dict
)
:
    lduMatrix::solver(fieldName, matrix, coupleBouCoeffs, coupleIntCoeffs, interfaces, dict),
    solverId_(solverId)
{
    readControls();
}


Foam::lduSolverPerformance Foam::cudaLduSolverBase::solve
(
    scalarField& x,
    const scalarField& b,
    const direction cmpt
) const
{
    lduSolverPerformance solverPerf(solverName(), fieldName());

    UPtrList<const lduMatrix> matrices(1);
    matrices.set(0, &matrix_);
    List<lduInterfaceFieldPtrsList> ifaces(1, interfaces_);
    b200Binding::systemEntry& entry = b200Binding::system(matrices, ifaces);   // cached on the addressing

    // fvMatrix::solve hands over the coefficients of component cmpt already (fvMatrixSolve.C: "interfaceBouCoeffs
    // ... .component(cmpt)"), so cmpt itself is not needed on the device
    List<const FieldField<Field, scalar>*> bou(1, &coupleBouCoeffs_), inte(1, &coupleIntCoeffs_);
    b200Binding::setCoeffs(entry, matrices, bou, inte);              // b200_sys_set_coeffs / _set_interface_coeffs

    b200_solver_opts opts;
    opts.solver = solverId_;
    opts.precond = b200Binding::precondId(dict());
    opts.tolerance = tolerance();
    opts.relTol = relTolerance();
    opts.minIter = minIter();
    opts.maxIter = maxIter();

    const bool dumping = dict().found("b200Dump");
    scalarField x0;
    scalarField history(dumping ? 64 : 0, 0.0);
    if (dumping) x0 = x;

    double* xp[1] = { x.begin() };
    const double* bp[1] = { b.begin() };
    b200_perf perf;
    b200Binding::check
    (
        b200_solve(entry.sys, &opts, xp, bp, &perf, dumping ? history.begin() : NULL, history.size()),
        "cudaLduSolverBase::solve(scalarField&, const scalarField&, const direction) const"
    );

    if (dumping)
    {
        history.setSize(min(history.size(), perf.nIterations + 1));
        UPtrList<const scalarField> xs(1), bs(1);
        xs.set(0, &x0);
        bs.set(0, &b);
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
