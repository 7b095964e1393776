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

    // coupled patches (coupleUpper / coupleLower, fvBlockMatrix.C:315-373): processor patches of a decomposed case go to
    // the library (b200_blk_add_interface); any other coupled patch type must not silently lose its coefficients
    DynamicList<label> procPatches, procRanks;
    forAll (m.interfaces(), patchI)
    {
        if (!m.interfaces().set(patchI)) continue;
        const processorLduInterfaceField* pp = dynamic_cast<const processorLduInterfaceField*>(&m.interfaces()[patchI]);
        if (!pp)
        {
            FatalErrorIn(where)
                << "the block-coupled device path serves processor patches only; patch " << patchI
                << " is another coupled type" << abort(FatalError);
        }
        procPatches.append(patchI);
        procRanks.append(pp->neighbProcNo());
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
        // the neighbour's index of the matching patch: the k-th processor patch of rank A towards rank B pairs with the
        // k-th of B towards A (decomposePar creates the two sides of a cut in the same order), b200Binding::describe pass 3
        List<labelList> table(Pstream::nProcs());
        table[Pstream::myProcNo()] = labelList(procRanks);
        if (Pstream::parRun())
        {
            Pstream::gatherList(table);
            Pstream::scatterList(table);
        }
        forAll (procPatches, i)
        {
            label kMine = 0;
            for (label j = 0; j < i; j++) if (procRanks[j] == procRanks[i]) kMine++;
            const labelList& theirs = table[procRanks[i]];
            label peerIface = -1, k = 0;
            forAll (theirs, e)
            {
                if (theirs[e] != Pstream::myProcNo()) continue;
                if (k == kMine) { peerIface = e; break; }
                k++;
            }
            if (peerIface < 0)
            {
                FatalErrorIn(where)
                    << "processor " << procRanks[i] << " has no processor patch that matches patch " << procPatches[i]
                    << abort(FatalError);
            }
            const unallocLabelList& fc = addr.patchAddr(procPatches[i]);
            int32_t index = -1;
            b200Binding::check(b200_blk_add_interface(sys, fc.size(), fc.begin(), procRanks[i], peerIface, &index), where);
        }
        blkCache_.insert(key, sys);
    }
    forAll (procPatches, i)
    {
        const double* cp = NULL;
        const int ck = coeffKind(m.coupleUpper()[procPatches[i]], cp);
        b200Binding::check(b200_blk_set_interface_coeffs(sys, i, ck, cp), where);
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
