/* b200_smooth.h -- C ABI of the Gauss-Seidel smoother / smoothSolver of libb200ldu.so (first piece).
 *
 * Drop-in boundary for the `lduMatrix::smoother` run-time selection table of foam-extend 4.1
 *     smoothSolver { smoother GaussSeidel; nSweeps n; }      (e.g. /root/reference/tutorials/fluidStructureInteraction/
 *                                                             HronTurekFsi3/system/fluid/fvSolution, `smoothSolver` entries)
 *     foam/matrices/lduMatrix/smoothers/GaussSeidel/GaussSeidelSmoother.C  (smooth)
 *     foam/matrices/lduMatrix/solvers/smoothSolver/smoothSolver.C          (solve)
 * for ONE lduMatrix (a segregated region equation) on one rank.  Coupled patches enter a Gauss-Seidel sweep only through
 * the right-hand side - the smoother starts every sweep from  bPrime = source;  updateMatrixInterfaces(-bouCoeffs, psi, bPrime)
 * - so an adapter (adapters/b200LduSolvers/cudaGaussSeidelSmoother.C) lets foam-extend build bPrime on the host, patches of
 * any kind included, and hands it to b200_gs_sweep; b200_gs_smooth / b200_gs_solve are the whole loop for a matrix without
 * coupled patches.  GAMG (agglomeration, restriction / prolongation) is NOT part of this library.
 *
 * Arithmetic of a sweep, per row c in ascending order (GaussSeidelSmoother.C):
 *     psi[c] = ((bPrime[c] - sum_{lower neighbours l, ascending} lower[f]*psiNew[l]) - sum_{owner faces f} upper[f]*psiOld[u[f]]) / diag[c]
 * On the device the rows run in wavefront-level order (a row waits for its lower neighbours' new values, published through
 * the output vector with a NaN sentinel); the order of a row's terms is the reference's, results are bit-identical to
 * oracle/ldu_oracle.c orc_gs_smooth.  Conventions as include/b200_ldu.h (host arrays, int return, b200_last_error).
 */
#ifndef B200_SMOOTH_H
#define B200_SMOOTH_H

#include "b200_ldu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_gs b200_gs;

int b200_gs_create(b200_ctx* ctx, int32_t nCells, int32_t nFaces, const int32_t* lowerAddr, const int32_t* upperAddr,
                   b200_gs** out);
int b200_gs_destroy(b200_gs* gs);
/* lower == NULL: symmetric matrix */
int b200_gs_set_coeffs(b200_gs* gs, const double* diag, const double* upper, const double* lower);
/* one sweep with the caller's bPrime (source with the coupled-patch contributions already added): psi in/out */
int b200_gs_sweep(b200_gs* gs, double* psi, const double* bPrime);
/* GaussSeidelSmoother::smooth for a matrix without coupled patches: nSweeps sweeps, bPrime = source */
int b200_gs_smooth(b200_gs* gs, double* psi, const double* source, int nSweeps);
/* smoothSolver::solve (nSweeps > 0) for a matrix without coupled patches; opts->solver / opts->precond are ignored;
 * history (optional): residual after every round of nSweeps sweeps, entry 0 = initial */
int b200_gs_solve(b200_gs* gs, const b200_solver_opts* opts, int nSweeps, double* psi, const double* source,
                  b200_perf* perf, double* history, int historyCap);

#ifdef __cplusplus
}
#endif
#endif
