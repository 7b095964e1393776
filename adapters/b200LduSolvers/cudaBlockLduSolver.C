/*---------------------------------------------------------------------------*\
  cudaBlockLduSolver.C -- see cudaBlockLduSolver.H.
\*---------------------------------------------------------------------------*/
#include "cudaBlockLduSolver.H"
#include "b200Binding.H"
#include "b200_blk.h"
#include "addToRunTimeSelectionTable.H"
#include "HashTable.H"

namespace Foam
{
    defineTypeNameAndDebug(cudaBlockCG, 0);
    defineTypeNameAndDebug(cudaBlockBiCGStab, 0);

    // BlockLduSolver<vector4> tables (BlockLduSolver.H: symMatrix / asymMatrix, arguments (fieldName, matrix, dict))
    BlockLduSolver<vector4>::addsymMatrixConstructorToTable<cudaBlockCG> addcudaBlockCGSymMatrixConstructorToTable_;
    BlockLduSolver<vector4>::addsymMatrixConstructorToTable<cudaBlockBiCGStab> addcudaBlockBiCGStabSymMatrixConstructorToTable_;
    BlockLduSolver<vector4>::addasymMatrixConstructorToTable<cudaBlockBiCGStab> addcudaBlockBiCGStabAsymMatrixConstructorToTable_;

    // device systems, cached on the addressing
    static HashTable<b200_blk*, const void*, Hash<const void*> > blkCache_;

    // active type of a coefficient array -> (kind, pointer to its doubles)
    static int coeffKind(const CoeffField<vector4>& c, const double*& p)
    {
        switch (c.activeType())
        {
            case blockCoeffBase::SCALAR:
                p = c.asScalar().begin();
                return B200_BLK_SCALAR;
            case blockCoeffBase::LINEAR:
                p = reinterpret_cast<const double*>(c.asLinear().begin());   // Field<vector4>: 4 contiguous doubles per entry
                return B200_BLK_LINEAR;
            case blockCoeffBase::SQUARE:
                p = reinterpret_cast<const double*>(c.asSquare().begin());   // Field<tensor4>: 16 doubles, row-major (i, j)
                return B200_BLK_SQUARE;
            default:
                FatalErrorIn("cudaBlockLduSolver: coeffKind(...)") << "unallocated coefficient array" << abort(FatalError);
        }
        return -1;
    }
}


Foam::cudaBlockLduSolverBase::cudaBlockLduSolverBase
(
    const int solverId,
    const word& fieldName,
    const BlockLduMatrix<vector4>& matrix,
    const dictionary& dict
)
:
    BlockLduSolver<vector4>(fieldName, matrix, dict),
    solverId_(solverId),
    tolerance_(dict.lookupOrDefault<scalar>("tolerance", 1e-6)),
    relTolerance_(dict.lookupOrDefault<scalar>("relTol", 0)),
    minIter_(dict.lookupOrDefault<label>("minIter", 0)),
    maxIter_(dict.lookupOrDefault<label>("maxIter", 1000))
{}


Foam::BlockSolverPerformance<Foam::vector4> Foam::cudaBlockLduSolverBase::solve(Field<vector4>& x, const Field<vector4>& b)
{
    const char* where = "cudaBlockLduSolverBase::solve(Field<vector4>&, const Field<vector4>&)";
    BlockSolverPerformance<vector4> solverPerf(solverName(), this->fieldName());
    const BlockLduMatrix<vector4>& m = this->matrix_;

    // coupled patches (coupleUpper / coupleLower, fvBlockMatrix.C:315-373): the block library of this version solves
    // one region on one rank; a matrix with a coupled or processor patch must not silently lose those coefficients
    forAll (m.interfaces(), patchI)
    {
        if (m.interfaces().set(patchI))
        {
            FatalErrorIn(where)
                << "the block-coupled device path has no coupled / processor patches yet (patch " << patchI << ")"
                << abort(FatalError);
        }
    }

    const lduAddressing& addr = m.lduAddr();
    const void* key = &addr;
    b200_blk* sys = NULL;
    if (blkCache_.found(key))
    {
        sys = blkCache_[key];
    }
    else
    {
        b200Binding::check
        (
            b200_blk_create(b200Binding::context(), addr.size(), addr.lowerAddr().size(), addr.lowerAddr().begin(), addr.upperAddr().begin(), &sys),
            where
        );
        blkCache_.insert(key, sys);
    }

    const double *dp = NULL, *up = NULL, *lp = NULL;
    const int dk = coeffKind(m.diag(), dp);
    const int uk = m.thereIsUpper() ? coeffKind(m.upper(), up) : B200_BLK_SCALAR;
    int lk = uk;
    if (m.asymmetric()) lk = coeffKind(m.lower(), lp);
    scalarField zeroUpper;
    if (!m.thereIsUpper())
    {
        zeroUpper.setSize(addr.lowerAddr().size(), 0.0);   // diagonal matrix
        up = zeroUpper.begin();
    }
    b200Binding::check(b200_blk_set_coeffs(sys, dk, dp, uk, up, lk, lp), where);

    b200_solver_opts opts;
    opts.solver = solverId_;
    opts.precond = b200Binding::precondId(b200Binding::precondName(this->dict()));
    opts.tolerance = tolerance_;
    opts.relTol = relTolerance_;
    opts.minIter = minIter_;
    opts.maxIter = maxIter_;

    b200_blk_perf perf;
    b200Binding::check
    (
        b200_blk_solve(sys, &opts, reinterpret_cast<double*>(x.begin()), reinterpret_cast<const double*>(b.begin()), &perf, NULL, 0),
        where
    );

    vector4 r0, r1;
    for (direction d = 0; d < 4; d++)
    {
        r0[d] = perf.initialResidual[d];
        r1[d] = perf.finalResidual[d];
    }
    solverPerf.initialResidual() = r0;
    solverPerf.finalResidual() = r1;
    solverPerf.nIterations() = perf.nIterations;
    solverPerf.converged() = perf.converged;
    solverPerf.singular() = perf.singular;
    return solverPerf;
}
