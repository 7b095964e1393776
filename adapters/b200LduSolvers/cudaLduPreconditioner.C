/*---------------------------------------------------------------------------*\
  cudaLduPreconditioner.C -- see cudaLduPreconditioner.H.
\*---------------------------------------------------------------------------*/
#include "cudaLduPreconditioner.H"
#include "addToRunTimeSelectionTable.H"

namespace Foam
{
    defineTypeNameAndDebug(cudaDIC, 0);
    defineTypeNameAndDebug(cudaDILU, 0);

    // lduMatrix::preconditioner tables (lduMatrix.H: symMatrix / asymMatrix, arguments
    // (matrix, coupleBouCoeffs, coupleIntCoeffs, interfaces, dict))
    lduMatrix::preconditioner::addsymMatrixConstructorToTable<cudaDIC> addcudaDICSymMatrixConstructorToTable_;
    lduMatrix::preconditioner::addasymMatrixConstructorToTable<cudaDILU> addcudaDILUAsymMatrixConstructorToTable_;
}


Foam::cudaLduPreconditionerBase::cudaLduPreconditionerBase
(
    const int precondId,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& coupleBouCoeffs,
    const FieldField<Field, scalar>& coupleIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces
)
:
    lduMatrix::preconditioner(matrix, coupleBouCoeffs, coupleIntCoeffs, interfaces),
    precondId_(precondId),
    sys_(NULL)
{
    // The preconditioner ignores the interfaces (DIC/DILU are sub-domain local), but the device system is the
    // cached one of this addressing, shared with the cuda* solvers, so the patches are described all the same.
    UPtrList<const lduMatrix> matrices(1);
    matrices.set(0, &matrix);
    List<lduInterfaceFieldPtrsList> ifaces(1, interfaces);
    b200Binding::systemEntry& entry = b200Binding::system(matrices, ifaces);
    List<const FieldField<Field, scalar>*> bou(1, &coupleBouCoeffs), inte(1, &coupleIntCoeffs);
    b200Binding::setCoeffs(entry, matrices, bou, inte);   // new coefficients: the library rebuilds rD at the next use
    sys_ = entry.sys;
}


void Foam::cudaLduPreconditionerBase::apply(scalarField& w, const scalarField& r, const int transpose) const
{
    const double* rp[1] = { r.begin() };
    double* wp[1] = { w.begin() };
    b200Binding::check
    (
        b200_precondition(sys_, precondId_, rp, wp, transpose),
        "cudaLduPreconditionerBase::apply(scalarField&, const scalarField&, const int) const"
    );
}
