/*---------------------------------------------------------------------------*\
  cudaGaussSeidelSmoother.C -- see cudaGaussSeidelSmoother.H.
\*---------------------------------------------------------------------------*/
#include "cudaGaussSeidelSmoother.H"
#include "addToRunTimeSelectionTable.H"
#include "HashTable.H"

namespace Foam
{
    defineTypeNameAndDebug(cudaGaussSeidel, 0);

    lduMatrix::smoother::addsymMatrixConstructorToTable<cudaGaussSeidel> addcudaGaussSeidelSymMatrixConstructorToTable_;
    lduMatrix::smoother::addasymMatrixConstructorToTable<cudaGaussSeidel> addcudaGaussSeidelAsymMatrixConstructorToTable_;

    // device matrices, cached on the addressing (the level schedule of a mesh is built once)
    static HashTable<b200_gs*, const void*, Hash<const void*> > gsCache_;
}


Foam::cudaGaussSeidel::cudaGaussSeidel
(
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& coupleBouCoeffs,
    const FieldField<Field, scalar>& coupleIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces
)
:
    lduMatrix::smoother(matrix, coupleBouCoeffs, coupleIntCoeffs, interfaces),
    gs_(NULL)
{
    const char* where = "cudaGaussSeidel::cudaGaussSeidel(...)";
    const lduAddressing& addr = matrix.lduAddr();
    const void* key = &addr;
    if (gsCache_.found(key))
    {
        gs_ = gsCache_[key];
    }
    else
    {
        b200Binding::check
        (
            b200_gs_create(b200Binding::context(), addr.size(), addr.lowerAddr().size(), addr.lowerAddr().begin(), addr.upperAddr().begin(), &gs_),
            where
        );
        gsCache_.insert(key, gs_);
    }
    b200Binding::check
    (
        b200_gs_set_coeffs(gs_, matrix.diag().begin(), matrix.upper().begin(), matrix.asymmetric() ? matrix.lower().begin() : NULL),
        where
    );
}


void Foam::cudaGaussSeidel::smooth(scalarField& psi, const scalarField& source, const direction cmpt, const label nSweeps) const
{
    scalarField bPrime(psi.size());
    for (label sweep = 0; sweep < nSweeps; sweep++)
    {
        // GaussSeidelSmoother.C: the coupled patches contribute to the right-hand side with the sign switched to the lhs
        bPrime = source;
        matrix_.initMatrixInterfaces(coupleBouCoeffs_, interfaces_, psi, bPrime, cmpt, true);
        matrix_.updateMatrixInterfaces(coupleBouCoeffs_, interfaces_, psi, bPrime, cmpt, true);
        b200Binding::check(b200_gs_sweep(gs_, psi.begin(), bPrime.begin()), "cudaGaussSeidel::smooth(...)");
    }
}
