/*---------------------------------------------------------------------------*\
  cudaLduSmoother.C -- see cudaLduSmoother.H.
\*---------------------------------------------------------------------------*/
#include "cudaLduSmoother.H"
#include "addToRunTimeSelectionTable.H"

namespace Foam
{
    defineTypeNameAndDebug(cudaDICSmoother, 0);
    defineTypeNameAndDebug(cudaDILUSmoother, 0);
    defineTypeNameAndDebug(cudaDICGaussSeidel, 0);

    // lduMatrix::smoother tables (lduMatrix.H: symMatrix / asymMatrix)
    lduMatrix::smoother::addsymMatrixConstructorToTable<cudaDICSmoother> addcudaDICSmootherSymMatrixConstructorToTable_;
    lduMatrix::smoother::addasymMatrixConstructorToTable<cudaDILUSmoother> addcudaDILUSmootherAsymMatrixConstructorToTable_;
    lduMatrix::smoother::addsymMatrixConstructorToTable<cudaDICGaussSeidel> addcudaDICGaussSeidelSymMatrixConstructorToTable_;
}


Foam::cudaLduSmootherBase::cudaLduSmootherBase
(
    const int precondId,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& coupleBouCoeffs,
    const FieldField<Field, scalar>& coupleIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces
)
:
    lduMatrix::smoother(matrix, coupleBouCoeffs, coupleIntCoeffs, interfaces),
    precondId_(precondId),
    sys_(NULL)
{
    // one-row system on the cached device copy of this addressing; the coupled patches take part in the residual
    UPtrList<const lduMatrix> matrices(1);
    matrices.set(0, &matrix);
    List<lduInterfaceFieldPtrsList> ifaces(1, interfaces);
    b200Binding::systemEntry& entry = b200Binding::system(matrices, ifaces);
    List<const FieldField<Field, scalar>*> bou(1, &coupleBouCoeffs), inte(1, &coupleIntCoeffs);
    b200Binding::setCoeffs(entry, matrices, bou, inte);
    sys_ = entry.sys;
}


void Foam::cudaLduSmootherBase::smooth(scalarField& psi, const scalarField& source, const direction cmpt, const label nSweeps) const
{
    double* xp[1] = { psi.begin() };
    const double* bp[1] = { source.begin() };
    b200Binding::check
    (
        b200_smooth(sys_, precondId_, nSweeps, xp, bp),
        "cudaLduSmootherBase::smooth(scalarField&, const scalarField&, const direction, const label) const"
    );
}
