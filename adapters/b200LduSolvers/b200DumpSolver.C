/*---------------------------------------------------------------------------*\
  b200DumpSolver.C -- see b200DumpSolver.H.
\*---------------------------------------------------------------------------*/
#include "b200DumpSolver.H"
#include "b200Binding.H"
#include "addToRunTimeSelectionTable.H"
#include "Pstream.H"
#include "OSspecific.H"

namespace Foam
{
    defineTypeNameAndDebug(b200DumpLduSolver, 0);
    defineTypeNameAndDebug(b200DumpCoupledLduSolver, 0);

    lduMatrix::solver::addsymMatrixConstructorToTable<b200DumpLduSolver> addb200DumpLduSolverSymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<b200DumpLduSolver> addb200DumpLduSolverAsymMatrixConstructorToTable_;
    coupledLduSolver::addsymMatrixConstructorToTable<b200DumpCoupledLduSolver>
        addb200DumpCoupledLduSolverSymMatrixConstructorToTable_;
    coupledLduSolver::addasymMatrixConstructorToTable<b200DumpCoupledLduSolver>
        addb200DumpCoupledLduSolverAsymMatrixConstructorToTable_;

    // the user's dictionary with another solver word and fixed iteration controls
    static dictionary controlled(const dictionary& dict, const label nIter)
    {
        dictionary d(dict);
        d.remove("solver");
        d.add("solver", word(dict.lookup("dumpSolver")));
        if (nIter >= 0)
        {
            d.remove("tolerance");
            d.remove("relTol");
            d.remove("minIter");
            d.remove("maxIter");
            d.add("tolerance", scalar(0));
            d.add("relTol", scalar(0));
            d.add("minIter", nIter);
            d.add("maxIter", nIter);
        }
        return d;
    }
}


Foam::b200DumpLduSolver::b200DumpLduSolver
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


Foam::lduSolverPerformance Foam::b200DumpLduSolver::solve(scalarField& x, const scalarField& b, const direction cmpt) const
{
    const label nHist = dict().lookupOrDefault<label>("dumpIterations", 20);
    const scalarField x0(x);
    scalarField history(nHist + 1, 0.0);
    for (label k = 0; k <= nHist; k++)
    {
        scalarField xk(x0);
        const dictionary dk(controlled(dict(), k));
        const lduSolverPerformance pk =
            lduMatrix::solver::New(fieldName(), matrix_, coupleBouCoeffs_, coupleIntCoeffs_, interfaces_, dk)->solve(xk, b, cmpt);
        history[k] = (k == 0) ? pk.initialResidual() : pk.finalResidual();
        if (pk.nIterations() < k)
        {
            // singular / breakdown exit before k iterations: the history ends here
            history.setSize(k);
            break;
        }
    }

    UPtrList<const lduMatrix> matrices(1);
    matrices.set(0, &matrix_);
    List<lduInterfaceFieldPtrsList> ifaces(1, interfaces_);
    List<const FieldField<Field, scalar>*> bou(1, &coupleBouCoeffs_), inte(1, &coupleIntCoeffs_);
    UPtrList<const scalarField> xs(1), bs(1);
    xs.set(0, &x0);
    bs.set(0, &b);
    const fileName dir(dict().lookupOrDefault<fileName>("dumpDirectory", "b200dump"));
    mkDir(dir);
    b200Binding::dump
    (
        dir/(fieldName() + "_proc" + name(Pstream::myProcNo()) + ".b200ldu"), matrices, ifaces, bou, inte, xs, bs,
        word(dict().lookup("dumpSolver")), b200Binding::precondName(dict()), tolerance(), relTolerance(), minIter(), maxIter(),
        history
    );

    const dictionary du(controlled(dict(), -1));
    return lduMatrix::solver::New(fieldName(), matrix_, coupleBouCoeffs_, coupleIntCoeffs_, interfaces_, du)->solve(x, b, cmpt);
}


Foam::b200DumpCoupledLduSolver::b200DumpCoupledLduSolver
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


Foam::coupledSolverPerformance Foam::b200DumpCoupledLduSolver::solve
(
    FieldField<Field, scalar>& x,
    const FieldField<Field, scalar>& b,
    const direction cmpt
) const
{
    const label nHist = dict().lookupOrDefault<label>("dumpIterations", 20);
    const label nRows = matrix_.size();
    const FieldField<Field, scalar> x0(x);
    scalarField history(nHist + 1, 0.0);
    for (label k = 0; k <= nHist; k++)
    {
        FieldField<Field, scalar> xk(x0);
        const dictionary dk(controlled(dict(), k));
        const coupledSolverPerformance pk =
            coupledLduSolver::New(fieldName(), matrix_, bouCoeffs_, intCoeffs_, interfaces_, dk)->solve(xk, b, cmpt);
        history[k] = (k == 0) ? pk.initialResidual() : pk.finalResidual();
        if (pk.nIterations() < k)
        {
            history.setSize(k);
            break;
        }
    }

    UPtrList<const lduMatrix> matrices(nRows);
    List<lduInterfaceFieldPtrsList> ifaces(nRows);
    List<const FieldField<Field, scalar>*> bou(nRows), inte(nRows);
    UPtrList<const scalarField> xs(nRows), bs(nRows);
    forAll (matrix_, rowI)
    {
        matrices.set(rowI, &matrix_[rowI]);
        ifaces[rowI] = interfaces_[rowI];
        bou[rowI] = &bouCoeffs_[rowI];
        inte[rowI] = &intCoeffs_[rowI];
        xs.set(rowI, &x0[rowI]);
        bs.set(rowI, &b[rowI]);
    }
    const fileName dir(dict().lookupOrDefault<fileName>("dumpDirectory", "b200dump"));
    mkDir(dir);
    b200Binding::dump
    (
        dir/(fieldName() + "_proc" + name(Pstream::myProcNo()) + ".b200ldu"), matrices, ifaces, bou, inte, xs, bs,
        word(dict().lookup("dumpSolver")), b200Binding::precondName(dict()), tolerance(), relTolerance(), minIter(), maxIter(),
        history
    );

    const dictionary du(controlled(dict(), -1));
    return coupledLduSolver::New(fieldName(), matrix_, bouCoeffs_, intCoeffs_, interfaces_, du)->solve(x, b, cmpt);
}
