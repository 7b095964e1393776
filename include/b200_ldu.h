/*
 * b200_ldu.h -- C ABI of libb200ldu.so: B200 (sm_100a) Krylov solve of (coupled) LDU systems.
 *
 * This is the drop-in boundary for the one hot path of multiRegionFoam this repository
 * accelerates.  Every entry point names the reference / foam-extend-4.1 interface it replaces;
 * INTEGRATION.md shows the foam-extend side binding (lduMatrix::solver / coupledLduSolver
 * adapters registered through addToRunTimeSelectionTable-style constructor tables).
 *
 *   caller in the reference                                    ->  entry points used
 *   coupledFvMatrix<scalar>::solve(dict)                           b200_sys_* + b200_solve
 *     (src/multiRegionSystem/multiRegionSystem.C:150-153)
 *   fvMatrix<Type>::solve()  (multiRegionSystem.C:293,             same, nRegions = 1, once per component
 *     src/regions/icoFluid/icoFluid.C:360,411-414, ...)
 *   monolithicCouplingFvPatchField::{init,update}InterfaceMatrix   b200_sys_add_interface (kind REGION_COUPLE):
 *     (src/fvPatchFields/.../monolithicCouplingFvPatchField.C:380-464)   static tables extracted once, applied on device
 *   processorFvPatchField::{init,update}InterfaceMatrix [FE]       b200_sys_add_interface (kind PROCESSOR): NCCL halo
 *   gSumProd / gSumMag / gAverage [FE Pstream]                     device reductions + ncclAllReduce inside b200_solve
 *   globalPolyPatch::patchFaceToGlobal/globalFaceToPatch +          b200_ggi_* (partitioned face transfer)
 *     ggiInterfaceToInterfaceMapping::transferFacesZoneToZone
 *     (src/numerics/globalPolyPatch/globalPolyPatchTemplates.C:140-235,
 *      src/numerics/interfaceToInterfaceMappings/.../ggiInterfaceToInterfaceMappingTemplates.C:37-77)
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function returns 0 on
 * success or a negative B200_E* code, with text in b200_last_error().  Host arrays are owned by
 * the caller and copied; no pointer is retained past the call.  One b200_ctx per process / MPI
 * rank bound to one GPU; not thread-safe; with nranks > 1 all ranks call b200_sys_finalize,
 * b200_solve*, b200_amul, b200_precondition collectively.  There is no CPU fallback: without a
 * usable CUDA device b200_ctx_create fails.
 *
 * scalar = IEEE double, label = int32 (reference build: -DWM_DP -DWM_LABEL_SIZE=32).
 */
#ifndef B200_LDU_H
#define B200_LDU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;
typedef struct b200_sys b200_sys;

/* error codes */
#define B200_OK 0
#define B200_EINVAL -1   /* bad argument / call order            */
#define B200_ECUDA -2    /* CUDA runtime failure                 */
#define B200_ENCCL -3    /* NCCL failure / NCCL not loadable     */
#define B200_ENOMEM -4
#define B200_ESTATE -5   /* e.g. solve before finalize/coeffs    */
#define B200_EUNSUPPORTED -6
#define B200_EDEVICE -7  /* device-side failure (sweep timeout)  */

/* interface kinds (lduInterfaceField flavours the adapter recognises) */
#define B200_IFACE_REGION_COUPLE 0 /* regionCouple / ggi (monolithicCouplingFvPatchField) */
#define B200_IFACE_PROCESSOR 1     /* processorFvPatchField                                */

/* solver / preconditioner selection (fvSolution words in comments) */
#define B200_SOLVER_PCG 0      /* cudaPCG      : PCG, CG (symmetric)            */
#define B200_SOLVER_BICGSTAB 1 /* cudaPBiCGStab: BiCGStab (lduSolvers, coupled) */
#define B200_SOLVER_PBICG 2    /* cudaPBiCG    : PBiCG, BiCG                    */

#define B200_PRECOND_NONE 0
#define B200_PRECOND_DIAGONAL 1
#define B200_PRECOND_DIC 2      /* cudaDIC  : DIC, FDIC                          */
#define B200_PRECOND_DILU 3     /* cudaDILU : DILU                               */
#define B200_PRECOND_CHOLESKY 4 /* Cholesky (per region: DIC if symmetric else DILU) */

typedef struct b200_solver_opts
{
    int solver;
    int precond;
    double tolerance; /* fvSolution default 1e-6 */
    double relTol;    /* 0                       */
    int minIter;      /* 0                       */
    int maxIter;      /* 1000                    */
} b200_solver_opts;

typedef struct b200_perf
{
    double initialResidual;
    double finalResidual;
    int nIterations;
    int converged;
    int singular;
    double normFactor;
    double deviceMs; /* CUDA-event time of the Krylov solve on the compute stream (no H2D/D2H) */
} b200_perf;

/* kernel classes for b200_get_kernel_times */
#define B200_K_AMUL 0
#define B200_K_IFACE 1
#define B200_K_SWEEP_FWD 2
#define B200_K_SWEEP_BWD 3
#define B200_K_VECTOR 4
#define B200_K_REDUCE 5
#define B200_K_PACK 6
#define B200_K_HALO 7
#define B200_K_NCLASSES 8

/* ---- context: one per process / rank ----------------------------------------------------- */
/* Fill 128 bytes with an NCCL unique id (rank 0 calls it and broadcasts the bytes). */
int b200_nccl_unique_id(void* out128);
/* ncclUniqueId may be NULL when nranks == 1. */
int b200_ctx_create(int device, int rank, int nranks, const void* ncclUniqueId, b200_ctx** out);
int b200_ctx_destroy(b200_ctx* ctx);
/* Text of the last error on this context (ctx may be NULL: last error of ctx_create). */
const char* b200_last_error(const b200_ctx* ctx);
int b200_version(void);

/* ---- system: the coupledLduMatrix of this rank (nRegions rows) ---------------------------- */
int b200_sys_create(b200_ctx* ctx, int nRegions, b200_sys** out);
int b200_sys_destroy(b200_sys* sys);
/* lduAddressing of region r: lowerAddr/upperAddr in OpenFOAM upper-triangular order. */
int b200_sys_set_region(b200_sys* sys, int r, int32_t nCells, int32_t nFaces,
                        const int32_t* lowerAddr, const int32_t* upperAddr);
/* One coupled patch of region r (call in patch-list order; processor patches last).
 * (peerRank, peerRegion, peerIface) name the shadow patch.  ggiOffsets == NULL: identity pairing
 * (conformal interface, processor patch); otherwise the CSR maps the shadow's nPeerFaces patch
 * values onto this patch: val[i] = sum_k peerVal[ggiAddr[k]] * ggiWeights[k].
 * Returns the interface index within region r (>= 0) or an error code. */
int b200_sys_add_interface(b200_sys* sys, int r, int kind, int32_t nFaces, const int32_t* faceCells,
                           int peerRank, int peerRegion, int peerIface, int32_t nPeerFaces,
                           const int32_t* ggiOffsets, const int32_t* ggiAddr, const double* ggiWeights);
/* Build device tables: row-packed Amul layout, interface/halo plans, sweep schedules.  Cached
 * until the addressing changes (topology change -> destroy and re-create). */
int b200_sys_finalize(b200_sys* sys);
/* lduMatrix coefficients of region r (host -> device).  lower == NULL: symmetric matrix. */
int b200_sys_set_coeffs(b200_sys* sys, int r, const double* diag, const double* upper, const double* lower);
/* boundaryCoeffs / internalCoeffs of interface iface of region r.  The transposed products (Tmul, PBiCG's shadow
 * system) apply intCoeffs, as lduMatrix::Tmul passes interfaceIntCoeffs to updateMatrixInterfaces.  intCoeffs == NULL
 * means "the same as bouCoeffs" (symmetric coupling); the CPU oracle follows the same rule (orc_add_iface). */
int b200_sys_set_interface_coeffs(b200_sys* sys, int r, int iface, const double* bouCoeffs, const double* intCoeffs);
/* regionInterfaceType::attach() / detach() (src/regionInterfaces/regionInterface/regionInterfaceType.C:543-627) flip the
 * regionCouple patches of an interface.  A detached patch must not take part in a coupled matrix-vector product:
 * monolithicCouplingFvPatchField::initInterfaceMatrixUpdate is fatal for it (monolithicCouplingFvPatchField.C:406-413);
 * while an interface of the system is detached, b200_solve / b200_amul return B200_ESTATE (the adapter turns that into
 * the reference's FatalError).  Only kind B200_IFACE_REGION_COUPLE. */
int b200_sys_set_interface_attached(b200_sys* sys, int r, int iface, int attached);
/* attach() forces the interpolation weights to be re-computed (regionInterfaceType.C:551-558; the FSI cases rebuild the
 * interpolator every interpolatorUpdateFrequency steps, :483-511): replace the GGI addressing / weights of interface
 * iface of region r (arguments as in b200_sys_add_interface; ggiOffsets == NULL: identity).  Allowed before and after
 * finalize; the cached device interface tables are rebuilt at the next use, the sweep / Amul layouts are kept.  The
 * patch itself (nFaces, faceCells) cannot change: a topology change means destroy and re-create. */
int b200_sys_set_interface_ggi(b200_sys* sys, int r, int iface, int32_t nPeerFaces, const int32_t* ggiOffsets,
                               const int32_t* ggiAddr, const double* ggiWeights);
/* A regionCouple patch whose SHADOW patch decomposePar spread over several ranks (each region is decomposed on its own:
 * tutorials/conjugateHeatTransfer/flowOverHeatedPlate/system/fluid/decomposeParDict:17-27, n (2 1 2)).  foam-extend
 * interpolates such a pair on the global face zones: the shadow's patch-internal field is expanded to the zone (all ranks
 * contribute their faces at zoneAddressing()), interpolated with zone-level addressing, and filtered back to the local
 * faces (the path monolithicCouplingFvPatchField.C:392-405 takes through regionCouplePatch().interpolate() when the
 * patch is not localParallel()).  Here: add the interface with nPeerFaces = size of the shadow ZONE and GGI tables whose
 * addresses are zone face labels (rows = the local faces), then name the pieces of the shadow zone: piece k is interface
 * pieceIface[k] of region pieceRegion[k] on rank pieceRank[k] (own rank allowed) and holds the zone faces
 * pieceZoneAddr[pieceOffsets[k] .. pieceOffsets[k+1]) in its patch face order.  Every rank that holds a piece of either
 * zone must list ALL non-empty pieces of the opposite zone (the halo plan sends a patch to the ranks it reads from).
 * Before b200_sys_finalize.  peerRank / peerRegion / peerIface of b200_sys_add_interface are ignored for such an interface. */
int b200_sys_set_interface_pieces(b200_sys* sys, int r, int iface, int nPieces, const int32_t* pieceRank,
                                  const int32_t* pieceRegion, const int32_t* pieceIface, const int32_t* pieceOffsets,
                                  const int32_t* pieceZoneAddr);
/* ---- device-side coefficient refresh of the temperature equations --------------------------- */
/* The two T regions re-assemble their fvScalarMatrix every time step
 *   B200_TEQN_CONDUCT    fvm::ddt(rho*cv, T) == fvm::laplacian(kappa, T)
 *                        (src/regions/conductTemperature/conductTemperature.C:135-142)
 *   B200_TEQN_TRANSPORT  rho*cp*(fvm::ddt(T) + fvm::div(phi, T)) == fvm::laplacian(kappa, T)
 *                        (src/regions/transportTemperature/transportTemperature.C:129-140)
 * (Euler ddt, Gauss upwind div, Gauss linear uncorrected laplacian, as in the tutorials' fvSchemes).  Instead of the
 * host assembling diag / upper / lower / source and b200_sys_set_coeffs + b200_upload copying them (the whole matrix
 * per step), the adapter hands over the static geometry once and the matrix is produced on the device from it and
 * from the resident field; per step only what changed (phi, kappa_f) crosses the bus.
 *
 * b200_sys_set_fv_geometry (after finalize; region r in upper-triangular face order): V = mesh.V(), magSf and
 * deltaCoeffs of the internal faces, and the boundary faces of the region in patch order (bCells = faceCells, their
 * internalCoeffs and boundary-source contributions as fvMatrix::addBoundaryDiag / addBoundarySource would add them,
 * coupled patches included). */
#define B200_TEQN_CONDUCT 0
#define B200_TEQN_TRANSPORT 1
int b200_sys_set_fv_geometry(b200_sys* sys, int r, const double* V, const double* magSf, const double* deltaCoeffs,
                             int32_t nBoundary, const int32_t* bCells, const double* bIntCoeffs, const double* bSrcCoeffs);
/* Assemble region r on the device: coefficients as b200_sys_set_coeffs would have set them, and the region's part of
 * the resident right-hand side b, with T.oldTime() = the resident x (b200_upload, or the result of the previous
 * b200_solve_resident).  rhoC = rho*cv | rho*cp, rDeltaT = 1/deltaT, kappa uniform unless kappaFace (host, nFaces:
 * the face-interpolated conductivity) is given; phi (host, nFaces) is the face flux of the transport form.  kappaFace
 * / phi == NULL: keep the table of the previous call.  Interface coefficients still come through
 * b200_sys_set_interface_coeffs. */
int b200_sys_assemble_T(b200_sys* sys, int r, int form, double rhoC, double rDeltaT, double kappa,
                        const double* kappaFace, const double* phi);
int64_t b200_sys_num_cells(const b200_sys* sys);
int64_t b200_sys_num_faces(const b200_sys* sys);

/* ---- solve -------------------------------------------------------------------------------- */
/* coupledLduSolver::solve / lduMatrix::solver::solve: x[r] (in: initial guess, out: solution) and
 * b[r] are host arrays of region r.  history (may be NULL) receives the normalised residual after
 * k iterations at history[k], k < historyCap. */
int b200_solve(b200_sys* sys, const b200_solver_opts* opts, double* const* x, const double* const* b,
               b200_perf* perf, double* history, int historyCap);
/* Device-resident variant: x and b already uploaded with b200_upload; result stays on device. */
int b200_upload(b200_sys* sys, const double* const* x, const double* const* b);
int b200_solve_resident(b200_sys* sys, const b200_solver_opts* opts, b200_perf* perf, double* history,
                        int historyCap);
int b200_download(b200_sys* sys, double* const* x);
/* Device-side copy of the resident field: save x -> x0 / restore x0 -> x, for repeated solves from one
 * initial guess (PICARD / DNA sub-iterations of assembleAndSolveEqns, multiRegionSystem.C:280-306). */
int b200_x_save(b200_sys* sys);
int b200_x_restore(b200_sys* sys);
/* Page-lock / unlock a caller-owned host array (an OpenFOAM Field's storage) so that the copies inside
 * b200_solve / b200_sys_set_coeffs run at full PCIe rate.  Optional. */
int b200_host_register(b200_ctx* ctx, void* p, uint64_t bytes);
int b200_host_unregister(b200_ctx* ctx, void* p);

/* ---- test hooks (single operations through the same kernels) -------------------------------- */
/* y = A x (transpose != 0: y = A^T x with internalCoeffs on interfaces) */
int b200_amul(b200_sys* sys, const double* const* x, double* const* y, int transpose);
/* lduMatrix::residual (foam/matrices/lduMatrix/lduMatrix/lduMatrixATmul.C; coupled rows: coupledLduMatrix): r = b - A x,
 * every row rounded as the reference's loops do (rA = source - diag*psi, then -= lower*psi[l] / upper*psi[u] in face
 * order, interfaces applied in the switchToLhs sense, monolithicCouplingFvPatchField.C:441-447). */
int b200_residual(b200_sys* sys, const double* const* x, const double* const* b, double* const* r);
/* lduMatrix::sumA: sumA[c] = diag[c] + the row's lower / upper coefficients, minus the boundaryCoeffs of the coupled
 * patches at their faceCells (same file).  Uses the coefficients of the last b200_sys_set_coeffs. */
int b200_sum_a(b200_sys* sys, double* const* sumA);
/* lduMatrix::smoother of the DIC / DILU family (foam-extend DICSmoother.C / DILUSmoother.C; with B200_PRECOND_CHOLESKY the "ILU"
 * smoother some tutorials select): nSweeps times  rA = residual(psi, source); rA = M^-1 rA; psi += rA  - the preconditioner's
 * own sweeps applied to the residual, so it runs at the speed of the DIC / DILU sweeps, on all regions, interfaces and ranks
 * of the system.  x in/out.  (The Gauss-Seidel smoother is include/b200_smooth.h.) */
int b200_smooth(b200_sys* sys, int precond, int nSweeps, double* const* x, const double* const* b);
/* w = M^-1 r with the given preconditioner (transpose != 0: preconditionT) */
int b200_precondition(b200_sys* sys, int precond, const double* const* r, double* const* w, int transpose);
/* reciprocal preconditioned diagonal of the last preconditioner setup */
int b200_get_rD(b200_sys* sys, int precond, double* const* rD);
/* global reductions as used by the solver: out[0] = sum a*b, out[1] = sum |a| */
int b200_reduce(b200_sys* sys, const double* const* a, const double* const* b, double* out2);

/* ---- profiling ------------------------------------------------------------------------------ */
/* enable != 0: bracket every kernel launch with CUDA events (adds launch overhead; use a separate run) */
int b200_set_profiling(b200_sys* sys, int enable);
/* accumulated device ms and launch counts per kernel class since the last reset */
int b200_get_kernel_times(b200_sys* sys, double* msPerClass, int64_t* launchesPerClass, int reset);
/* Debug: per-group counters of the sweep kernel of one direction (dir > 0 forward): 16 int64 per group
 * {consumer cycles, consumer wait cycles, start ns, end ns, producer polls, nT, general blocks, blocks,
 *  producer-0 cycles, its stage-wait / value-wait / spin cycles, 0...}.  enable != 0 arms the
 * counters for the following sweeps; out (may be NULL) receives the counters of the previous ones.
 * Returns the number of groups. */
int b200_debug_sweep_stats(b200_sys* sys, int dir, int enable, long long* out, int cap);
/* number of kernels launched by this library on this context since creation */
int64_t b200_launch_count(const b200_ctx* ctx);

/* ---- partitioned-coupling face transfer -------------------------------- */
/* result[i*nComp+d] = sum_k ff[addr[k]*nComp+d]*weights[k] on the device; host in/out. */
int b200_ggi_interpolate(b200_ctx* ctx, int32_t nTo, int32_t nFrom, const int32_t* offsets,
                         const int32_t* addr, const double* weights, const double* ff, int nComp,
                         double* result);
/* GGI weight construction: what GGIInterpolation<standAlonePatch, standAlonePatch>(zoneA, zoneB, ...,
 * SMALL, SMALL, rescale = true, BB_OCTREE) computes for the reference
 * (src/numerics/interfaceToInterfaceMappings/ggiInterfaceToInterfaceMapping/ggiInterfaceToInterfaceMapping.C:62-77; rebuilt
 * when the interface moves, src/regionInterfaces/regionInterface/regionInterfaceType.C:483-511, 551-558): for every master
 * face the slave faces it overlaps and the weights (intersection area / master face area in the master plane, entries
 * below nonOverlapTol dropped, rows rescaled to sum to one when rescale != 0).  Patches as standAlonePatch holds them:
 * faces = CSR of point labels (3..8 points), points = xyz triples.  Candidate pairs come from a host bounding-box hash
 * grid; clipping and rescaling run on the device.  Returns the number of addressing entries (>= 0) or an error code; the
 * result (the arguments b200_ggi_interpolate / b200_sys_set_interface_ggi take, master = the receiving side) is kept in
 * the context until the next build and copied out by b200_ggi_fetch.  For the other direction swap the patches. */
int b200_ggi_build(b200_ctx* ctx, int32_t nMaster, const int32_t* mFaceOffsets, const int32_t* mFaceLabels,
                   int32_t nMasterPoints, const double* mPoints, int32_t nSlave, const int32_t* sFaceOffsets,
                   const int32_t* sFaceLabels, int32_t nSlavePoints, const double* sPoints, double nonOverlapTol, int rescale);
/* offsets[nMaster+1], addr[nnz], weights[nnz] of the last b200_ggi_build */
int b200_ggi_fetch(b200_ctx* ctx, int32_t nMaster, int32_t* offsets, int32_t* addr, double* weights);
/* Conformal interfaces, directMapInterfaceToInterfaceMapping (directMap variant).
 * b200_direct_map_build = calcZone{A,B}ToZone{B,A}{Face,Point}Map
 * (src/numerics/interfaceToInterfaceMappings/directMapInterfaceToInterfaceMapping/directMapInterfaceToInterfaceMapping.C:
 * 147-168, 268-289, 389-410, 510-531): map[i] = the FIRST location j of the other zone with mag(to[i] - from[j]) < tol,
 * -1 if there is none; locations are xyz triples (face centres or points), tol = relTol_ (0.001, :52) * min edge length
 * of zone A (:153, :560-597), computed by the caller.  The reference's N^2 loop on one core becomes one thread per
 * receiving location.  Returns the number of unmatched locations (>= 0; > 0 is the reference's "interfaces are not
 * conformal" FatalError, :170-181) or an error code.
 * b200_direct_map_transfer = transferFacesZoneToZone / transferPointsZoneToZone
 * (.../directMapInterfaceToInterfaceMappingTemplates.C:37-86, 89-140): to[i] = from[map[i]]. */
int b200_direct_map_build(b200_ctx* ctx, int32_t nTo, const double* toXyz, int32_t nFrom, const double* fromXyz, double tol,
                          int32_t* map);
int b200_direct_map_transfer(b200_ctx* ctx, int32_t nTo, const int32_t* map, int32_t nFrom, const double* from, int nComp,
                             double* to);
/* globalPolyPatch::patchFaceToGlobal: zone field = all-reduce(sum) of the zero-padded scatter of the
 * local patch values through faceToGlobalAddr (collective over the context's ranks). */
int b200_patch_face_to_global(b200_ctx* ctx, int32_t nLocal, const int32_t* faceToGlobalAddr,
                              const double* pField, int nComp, int32_t nZoneFaces, double* gField);
/* globalPolyPatch::globalFaceToPatch: pField[i] = gField[faceToGlobalAddr[i]] */
int b200_global_face_to_patch(b200_ctx* ctx, int32_t nLocal, const int32_t* faceToGlobalAddr,
                              const double* gField, int nComp, double* pField);

#ifdef __cplusplus
}
#endif
#endif /* B200_LDU_H */
