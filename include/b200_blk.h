/* b200_blk.h -- C ABI of the block-coupled (vector4) solve path of libb200ldu.so.
 *
 * Drop-in boundary for
 *     fvBlockMatrix<vector4>::solve(const dictionary&)      /root/reference/filesToReplace/fvBlockMatrix.C:1360-1388
 *       -> BlockLduSolver<vector4>::New(psi.name(), *this, dict)->solve(psi.internalField(), source())
 * as reached from multiRegionSystem::assembleAndSolveEqns (src/multiRegionSystem/multiRegionSystem.C:293) for the
 * p-U coupled region type (src/regions/pUCoupledIcoFluid/pUCoupledIcoFluid.C:584-621); SURVEY.md 8 rows a18-a19.
 * A foam-extend adapter derives from BlockLduSolver<vector4> / BlockLduPrecon<vector4> and hands
 * the CoeffField<vector4> arrays of the matrix to these entry points as they lie in memory.
 *
 * Conventions (as include/b200_ldu.h): extern "C", plain pointers and sizes, host arrays owned by the caller and
 * copied by the call, int return 0 / negative B200_E* with text in b200_last_error(ctx).  No CPU fallback.
 *
 * Data: fields are Field<vector4> = [nCells][4] doubles.  A coefficient array has one ACTIVE TYPE for all its
 * entries, exactly like CoeffField<vector4> (fvBlockMatrix.C:84-126, 184-221): SCALAR (1 double per entry), LINEAR
 * (4, component-wise) or SQUARE (16, row-major (i,j)).  lower == NULL: symmetric matrix, the lower triangle is the
 * transposed upper coefficient.  Addressing is the lduAddressing of the mesh (upper-triangular face order).
 */
#ifndef B200_BLK_H
#define B200_BLK_H

#include "b200_ldu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_blk b200_blk;

#define B200_BLK_SCALAR 1
#define B200_BLK_LINEAR 4
#define B200_BLK_SQUARE 16

/* b200_solver_opts.solver for block systems (BlockLduSolver run-time selection names) */
#define B200_BLK_SOLVER_CG 0       /* cudaBlockCG       : BlockCGSolver       (symmetric) */
#define B200_BLK_SOLVER_BICGSTAB 1 /* cudaBlockBiCGStab : BlockBiCGStabSolver             */
/* b200_solver_opts.precond: B200_PRECOND_NONE (BlockNoPrecon), B200_PRECOND_DIAGONAL (BlockDiagonalPrecon),
 * B200_PRECOND_CHOLESKY (BlockCholeskyPrecon) */

/* BlockSolverPerformance<vector4>: the residuals are Type-valued (one per component), convergence is tested on
 * their cmptMax; normFactor is a scalar */
typedef struct b200_blk_perf
{
    double initialResidual[4];
    double finalResidual[4];
    int nIterations;
    int converged;
    int singular;
    double normFactor;
    double deviceMs; /* CUDA-event time of the solve on the compute stream (no H2D/D2H) */
} b200_blk_perf;

int b200_blk_create(b200_ctx* ctx, int32_t nCells, int32_t nFaces, const int32_t* lowerAddr, const int32_t* upperAddr,
                    b200_blk** out);
int b200_blk_destroy(b200_blk* sys);
/* coefficient kinds: B200_BLK_SCALAR / LINEAR / SQUARE; lower == NULL (lowerKind ignored): symmetric */
int b200_blk_set_coeffs(b200_blk* sys, int diagKind, const double* diag, int upperKind, const double* upper,
                        int lowerKind, const double* lower);

/* Coupled patches of a decomposed block matrix: BlockLduMatrix<vector4>::interfaces() entries that are
 * processorFvPatchField<vector4> (foam-extend BlockLduMatrixUpdateMatrixInterfaces.C; reached from the solvers' Amul
 * exactly like lduMatrix::updateMatrixInterfaces on the scalar path).  Add them after
 * b200_blk_create and before the first operation, in patch order; peerIface is the index the neighbour gives the
 * matching patch among ITS interfaces.  peerRank == own rank pairs two patches of this system (a cyclic-like pair; also
 * what the single-GPU tests use).  Amul then does  Ax[faceCells[f]] -= coupleUpper[f] * xNeighbour[f]  per patch in
 * patch order after the core product, and gSumProd / gSum / gAverage reduce over all ranks of the context.  Several ranks
 * need the peer-to-peer transport of the context (b200_ldu.h, B200_TRANSPORT); there is no NCCL leg on this path. */
int b200_blk_add_interface(b200_blk* sys, int32_t nFaces, const int32_t* faceCells, int32_t peerRank, int32_t peerIface,
                           int32_t* index);
/* coupleUpper of the patch, kind SCALAR / LINEAR / SQUARE, [nFaces][kind]; again after every assembly */
int b200_blk_set_interface_coeffs(b200_blk* sys, int32_t iface, int kind, const double* coupleUpper);

/* BlockLduMatrix<vector4>::Amul */
int b200_blk_amul(b200_blk* sys, const double* x, double* y);
/* BlockLduPrecon<vector4>::precondition(w, r) */
int b200_blk_precondition(b200_blk* sys, int precond, const double* r, double* w);
/* the (inverted) preconditioner diagonal, *kind doubles per cell (test / debug) */
int b200_blk_get_precon_diag(b200_blk* sys, int precond, double* out, int* kind);
/* gSumProd(a, b) and gSum(cmptMag(a)) of the device reductions (test hook): out5 = {sumProd, cmptMag[4]} */
int b200_blk_reduce(b200_blk* sys, const double* a, const double* b, double* out5);

/* BlockLduSolver<vector4>::solve(x, b): x in/out [nCells][4]; history (optional) [historyCap][4] residuals per
 * iteration, entry 0 = initial */
int b200_blk_solve(b200_blk* sys, const b200_solver_opts* opts, double* x, const double* b, b200_blk_perf* perf,
                   double* history, int historyCap);
/* the same in three steps, for device-resident timing */
int b200_blk_upload(b200_blk* sys, const double* x, const double* b);
int b200_blk_solve_resident(b200_blk* sys, const b200_solver_opts* opts, b200_blk_perf* perf, double* history,
                            int historyCap);
int b200_blk_download(b200_blk* sys, double* x);
/* keep / restore the uploaded initial guess on the device (repeated timed solves) */
int b200_blk_x_save(b200_blk* sys);
int b200_blk_x_restore(b200_blk* sys);
/* per kernel class CUDA-event times since the last reset: 0 Amul, 1 sweep fwd, 2 sweep bwd, 3 vector/reduce,
 * 4 preconditioner construction; enable with b200_blk_set_profiling */
int b200_blk_set_profiling(b200_blk* sys, int enable);
int b200_blk_get_kernel_times(b200_blk* sys, double* ms5, int64_t* launches5, int reset);

#ifdef __cplusplus
}
#endif
#endif
