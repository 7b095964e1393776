// kernels.cuh -- hand-written sm_100a kernels of the coupled LDU Krylov solve.
// All FP64, all HBM-bound; compiled with -fmad=false so that every product/sum rounds exactly
// like the reference build (g++ -O3, x86-64 baseline, no FMA contraction).
#pragma once

#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "schedule.hpp" // term codes (kCodeNone / kCodeOwn / kCodeShfl)

namespace b200
{

// value-as-flag sentinel of the sweep kernels: a quiet NaN payload no arithmetic produces
#define B200_SENTINEL_BITS 0xFFF8DEADBEEF0B20ull
#define B200_GREAT 1.0e+20
#define B200_SMALL 1.0e-20
#define B200_VSMALL 1.0e-300
#define B200_SMALL_ 1.0e-15

constexpr int kMaxDots = 4;
constexpr int kHistOnDevice = 1 << 16;

// scalar state of one solve, resident on the device (no host round trip inside an iteration)
struct DevScalars
{
    double red[8];
    double rho, rhoOld, alpha, omega, beta, betaOmega;
    double wArA, wArAold;
    double normFactor, initialResidual, finalResidual, xRef;
    double tolerance, relTol;
    double nGlobalCells;
    int minIter, maxIter;
    int nIter, done, converged, singular, restart, first, error;
    int histCap;
};

enum ScalarOp
{
    OP_NONE = 0,
    OP_XREF,
    OP_NORM_INIT_BICGSTAB,
    OP_NORM_INIT_PCG,
    OP_BICG_ALPHA,
    OP_BICG_OMEGA,
    OP_BICG_RESIDUAL,
    OP_PCG_RHO,
    OP_PCG_ALPHA,
    OP_PCG_RESIDUAL,
    OP_PBICG_RHO,
    OP_STORE_RED, // test hook: keep red[] only
};

__device__ __forceinline__ double sentinel() { return __longlong_as_double((long long)B200_SENTINEL_BITS); }
// L2 residency hints.  The sweeps exchange cross-group values through their output vector: a group polls slots
// that were pre-filled with the sentinel by the vector kernel BEFORE the sweep, and the sweep then streams ~300 MB
// (coefficient streams, input vectors) through the 126 MB L2.  Without hints the sentinel lines are the oldest
// lines in L2 and are evicted first, so that the first touch of a not-yet-written slot goes to DRAM - on the
// dependent path of every group that runs right behind its predecessor.  The fills are therefore stored evict_last
// and the streams are loaded evict_first.
#ifndef B200_L2_HINTS
#define B200_L2_HINTS 1
#endif
__device__ __forceinline__ unsigned long long l2_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_keep(double* p, double v, unsigned long long pol)
{
#if B200_L2_HINTS == 1
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ bool is_sentinel(double v)
{
    return (unsigned long long)__double_as_longlong(v) == B200_SENTINEL_BITS;
}
__device__ __forceinline__ double ld_volatile(const double* p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile(double* p, double v)
{
    asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ------------------------------------------------------------------------------ reductions
// block-level sum of up to kMaxDots values: warp shuffles, then one smem round
template <int ND>
__device__ __forceinline__ void block_reduce_store(double (&v)[ND], double* partials, int pstride, int slot)
{
    __shared__ double sm[ND][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < ND; k++)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) sm[k][wid] = v[k];
    }
    __syncthreads();
    if (wid == 0)
    {
        const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int k = 0; k < ND; k++)
        {
            double t = lane < nw ? sm[k][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) partials[(size_t)k * pstride + slot] = t;
        }
    }
}

__device__ __forceinline__ void stop_check(DevScalars* sc, volatile int* hostFlags)
{
    // lduMatrix::solver::stop + lduSolverPerformance::checkConvergence
    if (sc->nIter >= sc->minIter)
    {
        const bool conv = sc->finalResidual < sc->tolerance ||
                          (sc->relTol > B200_SMALL_ && sc->finalResidual <= sc->relTol * sc->initialResidual);
        sc->converged = conv ? 1 : 0;
        if (sc->nIter >= sc->maxIter || conv) sc->done = 1;
    }
    if (hostFlags)
    {
        hostFlags[1] = sc->nIter;
        if (sc->done)
        {
            // [2]: index of the solver-loop iteration that set done, + 2 (head of the solve: -1), in ONE word: with
            // several ranks every host leaves its loop at the same iteration derived from it (b200_ldu.cu, solve loops)
            hostFlags[2] = sc->nIter + 1;
            hostFlags[0] = 1;
        }
    }
}

__device__ __forceinline__ void bicg_beta(DevScalars* sc, double rwr)
{
    // bicgStabSolver::solve loop head: rhoOld = rho; rho = gSumProd(rw, r); beta = rho/rhoOld*(alpha/omega)
    sc->rhoOld = sc->rho;
    sc->rho = rwr;
    sc->beta = sc->rho / sc->rhoOld * (sc->alpha / sc->omega);
    sc->restart = 0;
    if (sc->rho == 0)
    {
        // restart if breakdown occurs: rw = r (done by k_bicg_p), rho = (r, r) (picked up in OP_BICG_ALPHA)
        sc->restart = 1;
        sc->alpha = 0;
        sc->omega = 0;
        sc->beta = 0;
    }
    sc->betaOmega = sc->beta * sc->omega;
}

__device__ void scalar_op(DevScalars* sc, int op, double* history, volatile int* hostFlags)
{
    switch (op)
    {
        case OP_XREF: sc->xRef = sc->red[0] / sc->nGlobalCells; break; // gAverage
        case OP_NORM_INIT_BICGSTAB:
        case OP_NORM_INIT_PCG:
            sc->normFactor = sc->red[0] + B200_SMALL;
            sc->initialResidual = sc->red[1] / sc->normFactor;
            sc->finalResidual = sc->initialResidual;
            sc->nIter = 0;
            sc->done = 0;
            sc->converged = 0;
            sc->singular = 0;
            sc->first = 1;
            if (history && sc->histCap > 0) history[0] = sc->initialResidual;
            stop_check(sc, hostFlags);
            if (op == OP_NORM_INIT_BICGSTAB)
            {
                sc->rho = B200_GREAT;
                sc->alpha = 0;
                sc->omega = B200_GREAT;
                bicg_beta(sc, sc->red[2]);
            }
            else
            {
                sc->wArA = B200_GREAT;
                sc->wArAold = B200_GREAT;
            }
            break;
        case OP_BICG_ALPHA:
            if (sc->restart) sc->rho = sc->red[1];
            sc->alpha = sc->rho / sc->red[0];
            break;
        case OP_BICG_OMEGA: sc->omega = sc->red[0] / sc->red[1]; break;
        case OP_BICG_RESIDUAL:
            sc->finalResidual = sc->red[0] / sc->normFactor;
            sc->nIter++;
            if (history && sc->nIter < sc->histCap) history[sc->nIter] = sc->finalResidual;
            stop_check(sc, hostFlags);
            bicg_beta(sc, sc->red[1]);
            break;
        case OP_PCG_RHO:
        case OP_PBICG_RHO:
            sc->wArAold = sc->wArA;
            sc->wArA = sc->red[0];
            sc->first = (sc->nIter == 0) ? 1 : 0;
            sc->beta = sc->wArA / sc->wArAold;
            break;
        case OP_PCG_ALPHA:
        {
            const double wApA = sc->red[0];
            if (!(fabs(wApA) / sc->normFactor > B200_VSMALL))
            {
                sc->singular = 1; // checkSingularity -> break before the update
                sc->done = 1;
                if (hostFlags)
                {
                    hostFlags[2] = sc->nIter + 2; // set inside iteration nIter (stop_check runs after nIter++)
                    hostFlags[0] = 1;
                }
            }
            else
                sc->alpha = sc->wArA / wApA;
            break;
        }
        case OP_PCG_RESIDUAL:
            sc->finalResidual = sc->red[0] / sc->normFactor;
            sc->nIter++;
            if (history && sc->nIter < sc->histCap) history[sc->nIter] = sc->finalResidual;
            stop_check(sc, hostFlags);
            break;
        default: break;
    }
}

// Sum the per-block partials of nd quantities in a fixed order (run-to-run deterministic), then
// either apply the scalar update (single rank) or leave red[] for the NCCL all-reduce.
struct PartCounts
{
    int n[kMaxDots];
};
// Peer-to-peer all-reduce of the <= kMaxDots sums of a reduction phase, fused into k_finalize (the global sums of
// gSumProd / gSumMag / gAverage).  Every rank owns a mailbox in its HBM, [2 parities][nranks][8 doubles],
// mapped into the address space of every other rank (cudaIpc over NVLink, or the same device for ranks that share
// one).  A rank stores its partial sums into slot [parity][its rank] of EVERY mailbox, fences, stores the sequence
// number into the slot's last word, then waits until its own mailbox holds the sequence number in all nranks slots
// and adds the slots up in rank order: every rank gets bit-identical sums, in one kernel, for the latency of one
// NVLink store instead of a kernel + ncclAllReduce + a kernel.  Two parities: a rank can be one all-reduce ahead of a
// peer that has not read its slots yet, never two (it needs that peer's contribution to finish the one in between).
struct PeerAR
{
    double* const* peerMbox; // [nranks] device-visible mailbox of every rank (own one included)
    double* mbox;            // this rank's mailbox
    int rank, nranks;        // nranks <= 1: no exchange
    unsigned long long seq;  // sequence number of this all-reduce (>= 1, the same on all ranks)
};
constexpr int kMboxSlot = 8; // doubles per mailbox slot: values 0 .. kMaxDots-1, sequence number in the last one

__device__ __forceinline__ void peer_allreduce(const PeerAR& ar, int nd, double* red, int* err)
{
    __syncthreads();
    const int par = (int)(ar.seq & 1ull);
    if ((int)threadIdx.x < ar.nranks)
    {
        const int peer = threadIdx.x;
        volatile double* dst = ar.peerMbox[peer] + ((size_t)par * ar.nranks + ar.rank) * kMboxSlot;
        for (int k = 0; k < nd; k++) dst[k] = red[k];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(dst + kMboxSlot - 1) = ar.seq;
        const volatile unsigned long long* f =
            reinterpret_cast<const volatile unsigned long long*>(ar.mbox + ((size_t)par * ar.nranks + peer) * kMboxSlot + kMboxSlot - 1);
        long long tries = 0;
        while (*f != ar.seq)
        {
            if (++tries > 64) __nanosleep(tries > 100000 ? 1000 : 40);
            if (tries > (1ll << 27))
            { // ~2 minutes: a peer is gone
                atomicExch(err, 2);
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const volatile double* m = ar.mbox + (size_t)par * ar.nranks * kMboxSlot;
        for (int k = 0; k < nd; k++)
        {
            double t = 0.0;
            for (int r = 0; r < ar.nranks; r++) t += m[(size_t)r * kMboxSlot + k];
            red[k] = t;
        }
    }
    __syncthreads();
}

// all-reduce of nd <= kMaxDots doubles that are already on the device (set-up time: the global cell count)
__global__ void k_peer_allreduce(double* v, int nd, PeerAR ar, int* err) { peer_allreduce(ar, nd, v, err); }

__global__ void __launch_bounds__(1024) k_finalize(const double* __restrict__ partials, int pstride, PartCounts cnt,
                                                    int nd, DevScalars* sc, int op, int applyOp, int force,
                                                    double* history, volatile int* hostFlags, PeerAR ar, int* err)
{
    if (sc->done && !force) return;
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < nd; k++)
    {
        double t = 0.0;
        const int nParts = cnt.n[k];
        for (int i = threadIdx.x; i < nParts; i += blockDim.x) t += partials[(size_t)k * pstride + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sm[wid] = t;
        __syncthreads();
        if (wid == 0)
        {
            double u = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
            if (lane == 0) sc->red[k] = u;
        }
        __syncthreads();
    }
    if (ar.nranks > 1) peer_allreduce(ar, nd, sc->red, err);
    if (applyOp && threadIdx.x == 0) scalar_op(sc, op, history, hostFlags);
}

__global__ void k_scalar_op(DevScalars* sc, int op, int force, double* history, volatile int* hostFlags)
{
    if (sc->done && !force) return;
    scalar_op(sc, op, history, hostFlags);
}

// ------------------------------------------------------------------------------ Amul
// Row-packed LDU product.  One thread per row; a warp walks one 32-row slice so that the
// coefficient and column streams are read as contiguous 256 B / 128 B segments.  Row c
// accumulates diag*x, then its lower neighbours by ascending column, then its upper neighbours
// by ascending column: the order of lduMatrix::Amul's face loop.
// ND fused dot products of the result: ND=1: (y,d0); ND=2: (y,d0),(y,y).  Rows touched by an
// interface are excluded from the dots here (k_iface adds them once their value is final).
// OP 0: y = A x.  OP 1: lduMatrix::residual, y = b - A x row by row as the reference's loops round it
// (rA = source - diag*psi; rA[u] -= lower*psi[l]; rA[l] -= upper*psi[u]); b arrives through d0.  OP 2: lduMatrix::sumA,
// y = diag + the row's off-diagonal coefficients (x is not read).
template <int ND, int OP = 0>
__global__ void __launch_bounds__(256) k_amul(int nRows, int nSlices, const double* __restrict__ diag,
                                               const int* __restrict__ sliceOff, const int* __restrict__ col,
                                               const double* __restrict__ val, const double* __restrict__ x,
                                               double* __restrict__ y, const double* __restrict__ d0,
                                               const unsigned* __restrict__ ifaceMask, double* partials, int pstride,
                                               const DevScalars* sc, int force)
{
    if (sc->done && !force) return;
    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    double dots[ND > 0 ? ND : 1];
#pragma unroll
    for (int k = 0; k < (ND > 0 ? ND : 1); k++) dots[k] = 0.0;
    for (int s = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); s < nSlices; s += gridDim.x * warpsPerBlock)
    {
        const int row = s * 32 + lane; // slot; padding slots have diag 0 and no entries
        const int o0 = sliceOff[s], o1 = sliceOff[s + 1];
        const size_t base = (size_t)o0 * 32 + lane;
        const int width = o1 - o0;
        if (row < nRows)
        {
            double acc = OP == 2 ? diag[row] : (OP == 1 ? d0[row] - diag[row] * x[row] : diag[row] * x[row]);
            int j = 0;
            if (OP != 0)
            {
                for (; j < width; j++)
                {
                    const int c0 = col[base + (size_t)j * 32];
                    const double v0 = val[base + (size_t)j * 32];
                    if (c0 >= 0) acc = OP == 2 ? acc + v0 : acc - v0 * x[c0];
                }
            }
            for (; j + 4 <= width; j += 4)
            {
                int c0 = col[base + (size_t)(j + 0) * 32], c1 = col[base + (size_t)(j + 1) * 32];
                int c2 = col[base + (size_t)(j + 2) * 32], c3 = col[base + (size_t)(j + 3) * 32];
                double v0 = val[base + (size_t)(j + 0) * 32], v1 = val[base + (size_t)(j + 1) * 32];
                double v2 = val[base + (size_t)(j + 2) * 32], v3 = val[base + (size_t)(j + 3) * 32];
                double x0 = c0 >= 0 ? x[c0] : 0.0, x1 = c1 >= 0 ? x[c1] : 0.0;
                double x2 = c2 >= 0 ? x[c2] : 0.0, x3 = c3 >= 0 ? x[c3] : 0.0;
                if (c0 >= 0) acc += v0 * x0;
                if (c1 >= 0) acc += v1 * x1;
                if (c2 >= 0) acc += v2 * x2;
                if (c3 >= 0) acc += v3 * x3;
            }
            for (; j < width; j++)
            {
                int c0 = col[base + (size_t)j * 32];
                double v0 = val[base + (size_t)j * 32];
                if (c0 >= 0) acc += v0 * x[c0];
            }
            y[row] = acc;
            if (ND > 0)
            {
                const bool touched = ifaceMask && ((ifaceMask[s] >> lane) & 1u);
                if (!touched)
                {
                    dots[0] += acc * d0[row];
                    if (ND > 1) dots[1] += acc * acc;
                }
            }
        }
    }
    if (ND > 0) block_reduce_store<(ND > 0 ? ND : 1)>(dots, partials, pstride, blockIdx.x);
}

// Interface update (monolithicCouplingFvPatchField::updateInterfaceMatrix,
// processorFvPatchField::updateInterfaceMatrix): one thread per touched row, entries in
// (non-processor, processor) x patch-list order:  y[row] -= coeff * pnf, where pnf is the
// shadow side's patchInternalField (GGI-weighted when non-conformal) gathered on the fly from x
// (same rank) or from the halo receive buffer (other rank).
// OP 1 (residual): the update is applied in the switchToLhs sense, y[row] += coeff * pnf
// (monolithicCouplingFvPatchField.C:441-447).  OP 2 (sumA): y[row] -= coeff (lduMatrix::sumA's interface loop).
template <int ND, int OP = 0>
__global__ void __launch_bounds__(128) k_iface(int nTouched, const int* __restrict__ rows,
                                                const int* __restrict__ rowStart, const int* __restrict__ entCoef,
                                                const int* __restrict__ entSrc, const int* __restrict__ entCnt,
                                                const int* __restrict__ gSrc, const double* __restrict__ gW,
                                                const double* __restrict__ coef, const double* __restrict__ x,
                                                const double* __restrict__ recv, double* __restrict__ y,
                                                const double* __restrict__ d0, double* partials, int pstride,
                                                int slotBase, const DevScalars* sc, int force,
                                                const unsigned long long* haloFlags, int nHaloPeers, unsigned long long haloSeq)
{
    if (sc->done && !force) return;
    if (nHaloPeers > 0)
    { // peer-to-peer halo: the neighbours' k_halo_push has published this exchange's sequence number when the data is here
        if ((int)threadIdx.x < nHaloPeers)
        {
            const volatile unsigned long long* f = haloFlags + threadIdx.x;
            long long tries = 0;
            while (*f < haloSeq)
            {
                if (++tries > 16) __nanosleep(tries > 100000 ? 1000 : 30);
                if (tries > (1ll << 27)) break; // a peer is gone: the solve's error word is set by the all-reduce that follows
            }
            __threadfence_system();
        }
        __syncthreads();
    }
    double dots[ND > 0 ? ND : 1];
#pragma unroll
    for (int k = 0; k < (ND > 0 ? ND : 1); k++) dots[k] = 0.0;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nTouched)
    {
        const int row = rows[t];
        double acc = y[row];
        for (int e = rowStart[t]; e < rowStart[t + 1]; e++)
        {
            const int cnt = entCnt[e];
            double pnf;
            if (OP == 2)
            {
                acc -= coef[entCoef[e]];
                continue;
            }
            if (cnt == 0)
            {
                const int s = entSrc[e];
                pnf = s >= 0 ? x[s] : __ldcg(recv + (-1 - s));
            }
            else
            {
                pnf = 0.0; // GGIInterpolation::interpolate: zero-initialised, accumulated in list order
                const int g0 = entSrc[e];
                for (int k = 0; k < cnt; k++)
                {
                    const int s = gSrc[g0 + k];
                    const double f = s >= 0 ? x[s] : __ldcg(recv + (-1 - s));
                    pnf += f * gW[g0 + k];
                }
            }
            if (OP == 1)
                acc += coef[entCoef[e]] * pnf;
            else
                acc -= coef[entCoef[e]] * pnf;
        }
        y[row] = acc;
        if (ND > 0)
        {
            dots[0] = acc * d0[row];
            if (ND > 1) dots[1] = acc * acc;
        }
    }
    if (ND > 0) block_reduce_store<(ND > 0 ? ND : 1)>(dots, partials, pstride, slotBase + blockIdx.x);
}

__global__ void k_halo_pack(int n, const int* __restrict__ cells, const double* __restrict__ x, double* __restrict__ sendbuf,
                            const DevScalars* sc, int force)
{
    if (sc->done && !force) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sendbuf[i] = x[cells[i]];
}

// Peer-to-peer halo (processorFvPatchField::initInterfaceMatrixUpdate): the patchInternalField of one
// processor patch is stored straight into the neighbour rank's receive buffer (mapped peer memory), and the last block
// of the launch - every block fences and counts itself in - publishes the sequence number of this exchange in the
// neighbour's flag word.  The neighbour's k_iface waits for that word; the interior rows of its Amul run meanwhile.
__global__ void k_halo_push(int n, const int* __restrict__ cells, const double* __restrict__ x, double* __restrict__ remote,
                            unsigned long long* remoteFlag, unsigned long long seq, unsigned* counter, const DevScalars* sc, int force)
{
    if (sc->done && !force) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) remote[i] = x[cells[i]];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned prev = atomicAdd(counter, 1u);
        if (prev == gridDim.x - 1)
        {
            *counter = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(remoteFlag) = seq;
        }
    }
}

// zone all-reduce over the peer mailboxes' exchange area (globalPolyPatch::patchFaceToGlobal) for contexts
// without an NCCL communicator: chunk c of the zone array is pushed into slot [parity][rank] of every rank's exchange
// buffer, then summed in rank order
__global__ void k_xchg_push(int n, const double* __restrict__ src, double* const* peerX, int rank, int nranks, int cap, int par,
                            unsigned long long seq, unsigned* counter)
{
    for (int peer = 0; peer < nranks; peer++)
    {
        double* dst = peerX[peer] + ((size_t)par * nranks + rank) * (size_t)(cap + 8);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned prev = atomicAdd(counter, 1u);
        if (prev == gridDim.x - 1)
        {
            *counter = 0u;
            __threadfence_system();
            for (int peer = 0; peer < nranks; peer++)
                *reinterpret_cast<volatile unsigned long long*>(peerX[peer] + ((size_t)par * nranks + rank) * (size_t)(cap + 8) + cap) = seq;
        }
    }
}
__global__ void k_xchg_sum(int n, double* __restrict__ dst, const double* xbuf, int nranks, int cap, int par, unsigned long long seq, int* err)
{
    if (threadIdx.x < nranks)
    {
        const volatile unsigned long long* f =
            reinterpret_cast<const volatile unsigned long long*>(xbuf + ((size_t)par * nranks + threadIdx.x) * (size_t)(cap + 8) + cap);
        long long tries = 0;
        while (*f != seq)
        {
            if (++tries > 64) __nanosleep(tries > 100000 ? 1000 : 40);
            if (tries > (1ll << 27))
            {
                atomicExch(err, 2);
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        double t = 0.0;
        for (int r = 0; r < nranks; r++) t += __ldcv(xbuf + ((size_t)par * nranks + r) * (size_t)(cap + 8) + i);
        dst[i] = t;
    }
}

__global__ void k_pack_sell(size_t nSlots, const int* __restrict__ src, const double* __restrict__ coef,
                            double* __restrict__ val, int transposeShift)
{
    // transposeShift = F: swap the roles of upper and lower (Tmul)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nSlots; i += (size_t)gridDim.x * blockDim.x)
    {
        int s = src[i];
        if (s >= 0 && transposeShift) s = s < transposeShift ? s + transposeShift : s - transposeShift;
        val[i] = s >= 0 ? coef[s] : 0.0;
    }
}

// ------------------------------------------------------------------------------ sweeps
// Pipelined DIC/DILU sweeps over the slot-ordered system (see schedule.hpp).  One warp per group.
// The per-step records (pre-multiplied coefficients + term codes) and the input vectors are
// streamed into a shared-memory ring by bulk asynchronous copies (TMA, cp.async.bulk + mbarrier);
// in-warp dependencies travel through one warp shuffle per time step, the own-lane dependency
// stays in a register, dependencies on other warps are read from the output vector, which
// doubles as its own ready flag (sentinel until written), and are prefetched a block of steps ahead.
struct PipeDev
{
    int nGroups;
    int dir;      // +1 forward, -1 backward (time reversed)
    int nStages;  // shared-memory ring stages
    int stageBytes;
    const int* gBase;
    const int* gNT;
    const int* gW;
    const int* gCH;
    const int* gShflMask; // bit 31: group must take the generic single-warp path
    const long long* gTermOff;
    const int* order; // ticket -> group (nullptr: identity)
    unsigned char* stream; // generic groups: 12*gTermOff[g]; step record = coef[W][32] f64 | code[W][32] i32
    const int* face;       // face of each term of the unified stream (pack time)
    // split streams of the fast path (schedule.hpp, build_split)
    const unsigned char* gFast;
    const int* gLg;
    const int* gRg;
    const int* gKg;
    const long long* gPOff;
    const long long* gCOff;
    const long long* gPFaceOff;
    const long long* gCFaceOff;
    unsigned char* pStream;
    unsigned char* cStream;
    const int* pFace;
    const int* cFace;
    int l2Ahead;      // blocks the L2 prefetch runs ahead of the stage fill
    int debugFlags;   // bit 0: consumer always takes the general (descriptor-driven) path (debug); bit 2: fault injection, the
                      // first group does nothing, so that every group that depends on it runs into the polling limit (tests)
    int spinLimit;    // polling rounds before a group gives up and the call returns B200_EDEVICE (B200_SWEEP_SPIN_LIMIT)
    long long* stats; // optional [8 * nGroups]: consumer cycles, wait cycles, start ns, end ns, producer polls, nT, general blocks, blocks (debug)
};


// MODE 0: forward  : acc = a[slot]*b[slot]; acc -= c * w[nbr]   (a = rD, b = rA, c = rD[row]*lower[f])
// MODE 1: backward : acc = a[slot];         acc -= c * w[nbr]   (a = forward result, c = rD[row]*upper[f])
// MODE 2: calcReciprocalD: acc = a[slot];   acc -= c / w[nbr]   (a = diag, c = upper[f]*lower[f])
template <int MODE>
__device__ __forceinline__ double sweep_apply(double acc, double c, double v)
{
    return MODE == 2 ? acc - c / v : acc - c * v;
}

__device__ __forceinline__ double ld_relaxed(const double* p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_relaxed(double* p, double v)
{
    asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v));
}

#ifndef B200_SWEEP_SPIN_LIMIT
#define B200_SWEEP_SPIN_LIMIT (1 << 24)
#endif
constexpr int kSweepSpinLimit = B200_SWEEP_SPIN_LIMIT; // polling rounds before a group gives up (B200_EDEVICE)
__device__ __noinline__ double sweep_spin(const double* p, int* err, int spinLimit)
{
    for (long long tries = 0; tries < (long long)spinLimit; tries++)
    {
        const double v = ld_relaxed(p);
        if (!is_sentinel(v)) return v;
        if (tries >= 64) __nanosleep(tries > 4096 ? 400 : 64); // later: a busy-spinning warp steals issue slots from the consumer warps
        if ((tries & 4095) == 4095 && *(volatile int*)err) break;
    }
    atomicExch(err, 1);
    return ld_relaxed(p);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// streamed once: evict_first (see B200_L2_HINTS)
__device__ __forceinline__ void bulk_g2s_stream(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol)
{
#if B200_L2_HINTS
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
#else
    bulk_g2s(dst, src, bytes, bar);
#endif
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
}

// Shared set-up of one group's shared-memory ring.
struct SweepRing
{
    int W, CHg, nT, base, nChunks, recBytes, schedBytes, NS, dir, stageBytes;
    unsigned shflMask, vecBytes, chunkBytes;
    unsigned long long* bars;
    unsigned char* stages;
    const unsigned char* gstream;
};

template <int MODE>
__device__ __forceinline__ void ring_issue(const SweepRing& R, const double* a, const double* b, int c, int st)
{
    unsigned char* dst = R.stages + (size_t)st * R.stageBytes;
    mbar_expect_tx(&R.bars[st], R.chunkBytes);
    bulk_g2s(dst, R.gstream + (size_t)c * R.schedBytes, (unsigned)R.schedBytes, &R.bars[st]);
    const long long s0 =
        R.dir > 0 ? ((long long)R.base + (long long)c * R.CHg) * 32 : ((long long)R.base + R.nT - (long long)(c + 1) * R.CHg) * 32;
    bulk_g2s(dst + R.schedBytes, a + s0, R.vecBytes, &R.bars[st]);
    if (MODE == 0) bulk_g2s(dst + R.schedBytes + R.vecBytes, b + s0, R.vecBytes, &R.bars[st]);
}

template <int MODE>
__device__ __forceinline__ SweepRing ring_setup(const PipeDev& S, int g, int lane, const double* a, const double* b, unsigned char* smem)
{
    SweepRing R;
    R.W = S.gW[g];
    R.CHg = S.gCH[g];
    R.nT = S.gNT[g];
    R.base = S.gBase[g];
    R.shflMask = (unsigned)S.gShflMask[g] & 0x7fffffffu;
    R.nChunks = R.nT / R.CHg;
    R.recBytes = R.W * 384;
    R.schedBytes = R.CHg * R.recBytes;
    R.NS = S.nStages;
    R.dir = S.dir;
    R.stageBytes = S.stageBytes;
    R.bars = reinterpret_cast<unsigned long long*>(smem);
    R.stages = smem + 128;
    R.gstream = S.stream + 12ll * S.gTermOff[g];
    R.vecBytes = (unsigned)(R.CHg * 256);
    R.chunkBytes = (unsigned)R.schedBytes + R.vecBytes * (MODE == 0 ? 2u : 1u);
    if (lane == 0)
    {
        for (int st = 0; st < R.NS; st++) mbar_init(&R.bars[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int c = 0; c < R.NS && c < R.nChunks; c++) ring_issue<MODE>(R, a, b, c, c);
    }
    __syncwarp();
    mbar_wait(&R.bars[0], 0u);
    return R;
}

// ================================================================================================
// Fast path: warp-specialised CTA over the SPLIT streams (schedule.hpp, build_split).
// The recurrence of a group is serial, so what sits on the dependent chain of a time step decides the
// speed of the whole sweep (measured on B200: DADD 8, DMUL 8, 64-bit SHFL 31, LDS 29, mbarrier try_wait 90
// cycles).  Everything static was decided on the host, and the CTA is organised so that the consumer
// warp's chain per step is ONE multiply and ONE subtract:
//   * lanes are skewed by kSkew steps (schedule.hpp).  With kSkew = 2 the neighbour-lane value a step needs is two
//     steps old and its shuffle and product are issued a step ahead, off the chain; with kSkew = 1 (the setting since
//     the producers, not the consumer, were found to set the pace) the shuffle is on the chain again, but a group is
//     31 steps shorter, a hop in j costs 31 instead of 62 steps of lag and a block seam spreads over 5 instead of 9 blocks;
//   * a loader thread streams the records and the input vectors of a block of kNH = 8 steps into one of
//     NS shared-memory stages with bulk asynchronous copies (TMA: cp.async.bulk + mbarrier);
//   * kNH producer warps (producer h prepares step h of every block) fetch the cross-group values
//     (prefetched one block ahead, verified against the sentinel), apply the row's LEADING terms in
//     reference order and leave {acc0, cval[]} in the hdr part of the stage; then they bump the stage's
//     counter - bits 8.. of the increment say whether the step is canonical;
//   * the consumer warp polls that counter with a plain load (no mbarrier wait, no fence on its path),
//     loads the 8 steps' operands at once and runs  acc = (acc0 - c0*shfl(prev2)) - c1*prev1 ; one store
//     per step.  Blocks with a non-canonical step (block seams) and calcReciprocalD (division) interpret
//     the descriptor bytes instead (general path).  After a block it resets the counter and publishes its
//     progress, which is what the loader polls before it refills the stage.
constexpr int kNH = kSweepBlock; // producer warps = steps per block (8)
// Producer organisation (B200_PROD_SETS > 0): kProdSets sets of producer warps work on kProdSets consecutive blocks at
// the same time (set s takes the blocks blk = s mod kProdSets); inside a set a producer prepares kProdM consecutive
// steps of the block.  A producer's step is a chain of latencies (stage wait, code loads, the L2 round trip of the
// cross-group values, the check, the hand-over) that no amount of tuning brought under ~850 cycles per block; with
// one set that chain IS the block time of the group (106 cycles per step against the consumer's 61).  With several
// sets the chains of consecutive blocks overlap, so the loads of the cross-group values can be issued when they are
// needed (no prefetch a block ahead, no second register set) and a group's pace is the consumer's.
#ifndef B200_PROD_SETS
#define B200_PROD_SETS 0
#endif
#ifndef B200_PROD_M
#define B200_PROD_M 4
#endif
// B200_PROD_DIRECT (needs sets): the producers read THEIR operands - the P-records (coefficients and codes of the leading
// cross-group terms) and the input vectors a, b - straight from global memory (coalesced 8-byte-per-lane loads of lines
// the loader has pulled into L2), and only the consumer's operands (C-block) travel through the bulk-copy ring.  A stage
// shrinks from P + C + a + b + hdr (26.6 KB on C2) to C + hdr (14.3 KB): eight stages instead of four fit next to a second
// CTA on the SM.  Measured on B200 a refill lands ~1 900 cycles after its issue even on an otherwise idle GPU and
// ~2 100 - 2 800 cycles under load; with four stages (three blocks of look-ahead, ~560 cycles of consumer work each) the
// consumer of even the FIRST group - which depends on nobody - waited for its ring about a third of the time.
#ifndef B200_PROD_DIRECT
#define B200_PROD_DIRECT 1
#endif
#ifndef B200_PROD_PIPE
#define B200_PROD_PIPE 1
#endif
constexpr bool kProdDirect = B200_PROD_SETS > 0 && B200_PROD_DIRECT;
// B200_STORE_SMEM: the consumer leaves its results in shared memory (in the slot of the hdr plane it has just read acc0
// from) and the loader warp writes a finished block to the output vector - 8 coalesced stores per lane - before it
// refills the stage.  A global store costs the consumer ~16 cycles of issue per step (scripts/micro/consumer.cu: 43 ->
// 59 cycles per step), a shared-memory store 3; followers need the whole block anyway, so they see it ~200 cycles later.
#ifndef B200_STORE_SMEM
#define B200_STORE_SMEM 0
#endif
constexpr bool kStoreSmem = B200_STORE_SMEM != 0;
// B200_CONS_PIPE: canonical blocks are worked in two halves whose operand loads overlap the other half's recurrence
// (the operands of steps 4-7 are loaded while steps 0-3 run, those of the next block's steps 0-3 - when it has been
// delivered - while steps 4-7 run), instead of 24 loads in front of every block.
#ifndef B200_CONS_PIPE
#define B200_CONS_PIPE 0
#endif
#ifndef B200_PROBE
#define B200_PROBE 0 // timing probes of the consumer loop (wrong results), see split_consumer
#endif
constexpr bool kConsPipe = B200_CONS_PIPE != 0;
constexpr int kProdSets = B200_PROD_SETS;
constexpr int kProdM = B200_PROD_SETS > 0 ? B200_PROD_M : 1;
constexpr int kProdPerSet = kNH / kProdM;
constexpr int kProducers = B200_PROD_SETS > 0 ? kProdSets * kProdPerSet : kNH;
static_assert(kNH % kProdM == 0, "a producer takes a whole number of steps of a block");
constexpr int kTraceBlocks = 142; // debug: per-block time stamps (clock64 of the group's SM) of the first blocks
constexpr int kStatsStride = 16 + 8 * kTraceBlocks; // debug counters per group: 16 totals, then per block 8 stamps:
// consumer {ready seen, done}, loader issue, producer 0 {step start, next stage landed, values checked, delivered},
// producer 7 delivered
constexpr int kSweepSmemHeader = 512; // bytes in front of the stages (barriers, counters)
constexpr int kSweepMaxStages = 16;
constexpr int kL2Ahead = 4;      // default number of blocks the L2 prefetch runs ahead of the stage fill (PipeDev::l2Ahead)

// Flag words in shared memory.  A formal release/acquire pair costs a MEMBAR.ALL.CTA on the releasing side, which
// also waits for the warp's outstanding GLOBAL accesses (the consumer's st.cg results, the producers' prefetches
// of cross-group values): hundreds of cycles per block on the dependent path.  The handshakes below therefore use
// relaxed accesses and rely on the shared-memory pipeline of an SM processing the accesses of a warp in program
// order (STS data ... ATOMS flag on one side, LDS flag ... LDS data on the other); "memory" clobbers keep the
// compiler from moving anything across them.
__device__ __forceinline__ unsigned ld_flag_smem(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag_smem(unsigned* p, unsigned v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_flag_smem(unsigned* p, unsigned v)
{
    asm volatile("red.relaxed.cta.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

struct SplitCtx
{
    int nT, base, nBlocks, NS, stageBytes;
    int Lg, Rg, Kg, pRec, cRec, hdrStep;
    int offHdr, offP, offA; // byte offsets inside a stage: C-block at 0, then hdr, P-records, a (b)
    unsigned pBytes, cBytes, rawBytes;
    unsigned long long* rawBar; // [NS] TMA completion
    unsigned* cnt;              // [NS] producer arrivals of the block in the stage (+ 256 per non-canonical step)
    unsigned* done;             // blocks the consumer has finished
    unsigned char* stages;
    const unsigned char *pStream, *cStream;
    unsigned long long polStream; // L2 evict_first policy of the streamed operands
};

__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes, unsigned long long pol)
{
#if B200_L2_HINTS
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
#endif
}

// Pull the records and input vectors of block blk into L2 well ahead of the stage fill, so that the fill itself
// sees L2 latency only: the few stages that fit in shared memory would not cover HBM latency at one block per
// ~500 cycles.
template <int MODE>
__device__ __forceinline__ void split_prefetch(const SplitCtx& C, const double* a, const double* b, int blk)
{
    constexpr int dir = MODE == 1 ? -1 : 1;
    if (blk >= C.nBlocks) return;
    if (C.pBytes) bulk_prefetch_l2(C.pStream + (size_t)blk * C.pBytes, C.pBytes, C.polStream);
    bulk_prefetch_l2(C.cStream + (size_t)blk * C.cBytes, C.cBytes, C.polStream);
    const long long s0 = dir > 0 ? ((long long)C.base + (long long)blk * kNH) * 32 : ((long long)C.base + C.nT - (long long)(blk + 1) * kNH) * 32;
    bulk_prefetch_l2(a + s0, kNH * 256, C.polStream);
    if (MODE == 0) bulk_prefetch_l2(b + s0, kNH * 256, C.polStream);
}

template <int MODE>
__device__ __forceinline__ void split_issue(const SplitCtx& C, const double* a, const double* b, int blk, int st)
{
    constexpr int dir = MODE == 1 ? -1 : 1;
    unsigned char* dst = C.stages + (size_t)st * C.stageBytes;
    mbar_expect_tx(&C.rawBar[st], C.rawBytes);
    bulk_g2s_stream(dst, C.cStream + (size_t)blk * C.cBytes, C.cBytes, &C.rawBar[st], C.polStream);
    if (kProdDirect) return; // the producers fetch their operands themselves
    if (C.pBytes) bulk_g2s_stream(dst + C.offP, C.pStream + (size_t)blk * C.pBytes, C.pBytes, &C.rawBar[st], C.polStream);
    const long long s0 = dir > 0 ? ((long long)C.base + (long long)blk * kNH) * 32 : ((long long)C.base + C.nT - (long long)(blk + 1) * kNH) * 32;
    bulk_g2s_stream(dst + C.offA, a + s0, kNH * 256, &C.rawBar[st], C.polStream);
    if (MODE == 0) bulk_g2s_stream(dst + C.offA + kNH * 256, b + s0, kNH * 256, &C.rawBar[st], C.polStream);
}

__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s32(unsigned a)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds_u8(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// operands streamed once from global memory by the producers (direct mode): L2 only
__device__ __forceinline__ double ldg_stream_f64(const void* p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_stream_s32(const void* p)
{
    int v;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void mbar_wait_u32(unsigned bar, unsigned parity)
{
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
}

// ------------------------------------------------------------------------------------------ consumer
// The consumer warp executes every instruction of the group's critical path itself, so the block loop is kept
// minimal: the C-block and the hdr part of a stage are plane-major (operands of step q of a plane at q * 256 from
// the plane's base), which turns the 24 operand loads of a canonical block into loads at immediate offsets from
// two base registers.
template <int MODE, int RG, int STATS>
__device__ __forceinline__ void split_consumer(const PipeDev& S, const SplitCtx& C, const int g, const int lane, double* out)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int dir = MODE == 1 ? -1 : 1;
    constexpr long long outStride = dir > 0 ? 32 : -32;
    constexpr int PL = kNH * 256; // bytes of one plane of a block
    const int nT = C.nT;
    double* outPtr = out + ((long long)C.base + (dir > 0 ? 0 : nT - 1)) * 32 + lane;
    const int srcLane = (lane - dir) & 31; // the linked neighbour lane of a canonical step
    // cheap accumulators (two clock reads per block, one write at the end) also run in the product instantiation when the
    // caller armed the counters with debug flag 2; the per-block time stamps are in the STATS instantiation only
    constexpr bool statsOn = STATS != 0; // STATS: 0 product (no instrumentation at all), 1 per-block time stamps + counters, 2 counters only
    const bool forceGeneral = MODE == 2 || (S.debugFlags & 1);
    const int Kg = C.Kg;
    double h[kSkew]; // h[k]: the value this lane produced k+1 steps ago
#pragma unroll
    for (int k = 0; k < kSkew; k++) h[k] = 0.0;
    auto push = [&](double v) {
#pragma unroll
        for (int k = kSkew - 1; k > 0; k--) h[k] = h[k - 1];
        h[0] = v;
    };
    const unsigned char* const stage0 = C.stages + lane * 8;
    const unsigned char* const stageEnd = stage0 + (size_t)C.NS * C.stageBytes;
    const unsigned char* sb = stage0; // C-block of the current stage (this lane's column)
    unsigned* cntp = C.cnt;
    // result of step q of the current block: to the output vector, or (kStoreSmem) into the acc0 slot of the hdr plane,
    // which this lane has read before it gets here; the loader warp stores the block (sweep_group_split)
    auto put = [&](const unsigned char* hdq, double* gp, double v) {
        if (kStoreSmem)
            sts_f64(smem_u32(hdq), v);
        else
            st_relaxed(gp, v);
    };
    auto wait_block = [&](unsigned* cp) -> unsigned {
        unsigned cc;
        do cc = ld_flag_smem(cp);
        while ((cc & 0xffu) != (unsigned)kNH);
        return cc;
    };
    constexpr int HB = kNH / 2;
    double A0[HB], A1[HB], A2[HB]; // kConsPipe: a0, c0, c1 of steps 0 .. HB-1 of the block the loop is about to work on
    auto loadA = [&](const unsigned char* s) {
#pragma unroll
        for (int q = 0; q < HB; q++)
        {
            A0[q] = *reinterpret_cast<const double*>(s + C.offHdr + q * 256);
            A1[q] = *reinterpret_cast<const double*>(s + PL + q * 256);
            A2[q] = *reinterpret_cast<const double*>(s + 2 * PL + q * 256);
        }
    };
    unsigned cPre = 0u;
    bool havePre = false; // kConsPipe: the next block's counter has been read (cPre) and its first half is in A
    long long tWait = 0, t0 = 0, g0 = 0, nGeneral = 0, tCanon = 0, tOther = 0, nCanon = 0;
    if (statsOn)
    {
        t0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    }
    for (int blk = 0; blk < C.nBlocks; blk++)
    {
        unsigned c;
        if (kConsPipe && havePre)
            c = cPre;
        else if (statsOn)
        {
            const long long w0 = clock64();
            c = wait_block(cntp);
            tWait += clock64() - w0;
        }
        else
            c = wait_block(cntp);
        if (kConsPipe && !havePre) loadA(sb);
        havePre = false;
        // the stage after this one (kConsPipe looks ahead into it)
        const unsigned char* sbN = sb + C.stageBytes;
        unsigned* cntN = cntp + 1;
        if (sbN == stageEnd)
        {
            sbN = stage0;
            cntN = C.cnt;
        }
        if (STATS == 1 && lane == 0 && blk < kTraceBlocks) S.stats[(long long)kStatsStride * g + 16 + blk * 8 + 0] = clock64();
        const unsigned char* hd = sb + C.offHdr;
        const long long b0 = statsOn ? clock64() : 0;
        const bool canon = c == (unsigned)kNH && !forceGeneral;
        if (c == (unsigned)kNH && !forceGeneral && kConsPipe)
        {
            double B0[HB], B1[HB], B2[HB];
#pragma unroll
            for (int q = 0; q < HB; q++)
            {
                B0[q] = *reinterpret_cast<const double*>(hd + (HB + q) * 256);
                B1[q] = *reinterpret_cast<const double*>(sb + PL + (HB + q) * 256);
                B2[q] = *reinterpret_cast<const double*>(sb + 2 * PL + (HB + q) * 256);
            }
#pragma unroll
            for (int q = 0; q < HB; q++)
            {
                const double sh = __shfl_sync(FULL, h[kSkew - 1], srcLane);
                const double pre = A0[q] - A1[q] * sh;
                const double acc = pre - A2[q] * h[0];
                put(hd + q * 256, outPtr + q * outStride, acc);
                push(acc);
            }
            if (blk + 1 < C.nBlocks)
            { // the next block, if it has been delivered: its first half while the second half of this one runs
                cPre = ld_flag_smem(cntN);
                if ((cPre & 0xffu) == (unsigned)kNH)
                {
                    havePre = true;
                    loadA(sbN);
                }
            }
#pragma unroll
            for (int q = 0; q < HB; q++)
            {
                const double sh = __shfl_sync(FULL, h[kSkew - 1], srcLane);
                const double pre = B0[q] - B1[q] * sh;
                const double acc = pre - B2[q] * h[0];
                put(hd + (HB + q) * 256, outPtr + (HB + q) * outStride, acc);
                push(acc);
            }
        }
        else if (c == (unsigned)kNH && !forceGeneral)
        {
            double a0[kNH], c0[kNH], c1[kNH];
#pragma unroll
            for (int q = 0; q < kNH; q++)
            {
#if B200_PROBE == 4 // timing probe (wrong results): no operand loads
                a0[q] = 1.0 + q;
                c0[q] = 0.25;
                c1[q] = 0.125;
#else
                a0[q] = *reinterpret_cast<const double*>(hd + q * 256);
                c0[q] = *reinterpret_cast<const double*>(sb + PL + q * 256);
                c1[q] = *reinterpret_cast<const double*>(sb + 2 * PL + q * 256);
#endif
            }
#pragma unroll
            for (int q = 0; q < kNH; q++)
            {
#if B200_PROBE == 3 // timing probe (wrong results): no shuffle
                const double sh = h[kSkew - 1];
#else
                const double sh = __shfl_sync(FULL, h[kSkew - 1], srcLane);
#endif
                const double pre = a0[q] - c0[q] * sh;
                const double acc = pre - c1[q] * h[0];
#if B200_PROBE == 2 // timing probe (wrong results): no store
                if (acc == 1.2345e300) put(hd + q * 256, outPtr + q * outStride, acc);
#else
                put(hd + q * 256, outPtr + q * outStride, acc);
#endif
                push(acc);
            }
        }
        else if ((c & 0xff00u) == 0u && !forceGeneral)
        {
            // Blocks with DUAL steps (schedule.hpp): every lane is in canonical form or in seam form (own-lane term first,
            // then cval 0, then a shuffled value or a cval).  The canonical operands of the 8 steps are loaded at once as
            // above; the arrival counter carries one bit per dual step, and only those steps (with lanes skewed by 2 steps,
            // every other step of a seam crossing; every step with kSkew = 1) load the flags and the extra operands and evaluate the seam form next
            // to the canonical one - one multiply and three subtractions on the dependent chain.
            nGeneral++;
            const unsigned dualMask = c >> 16;
            double a0[kNH], c0[kNH], c1[kNH];
#pragma unroll
            for (int q = 0; q < kNH; q++)
            {
                a0[q] = *reinterpret_cast<const double*>(hd + q * 256);
                c0[q] = *reinterpret_cast<const double*>(sb + PL + q * 256);
                c1[q] = *reinterpret_cast<const double*>(sb + 2 * PL + q * 256);
            }
#pragma unroll
            for (int q = 0; q < kNH; q++)
            {
                const double sh = __shfl_sync(FULL, h[kSkew - 1], srcLane);
                const double pSh = c0[q] * sh;
                const double pOwn = c1[q] * h[0];
                double acc = (a0[q] - pSh) - pOwn;
                if (dualMask & (1u << q))
                {
                    const unsigned fl = *reinterpret_cast<const unsigned*>(sb + q * 256 + 4) >> 26; // meta bits 58..: seam form, selB
                    const double c2 = RG > 2 ? *reinterpret_cast<const double*>(sb + 3 * PL + q * 256) : 0.0;
                    const double cv0 = Kg > 0 ? *reinterpret_cast<const double*>(hd + PL + q * 256) : 0.0;
                    const double cv1 = Kg > 1 ? *reinterpret_cast<const double*>(hd + 2 * PL + q * 256) : 0.0;
                    const unsigned selB = (fl >> 1) & 3u;
                    const double vB = selB == 2u ? cv1 : (selB == 1u ? cv0 : sh);
                    const double sm = ((a0[q] - pOwn) - c2 * cv0) - c0[q] * vB;
                    if (fl & 1u) acc = sm;
                }
                put(hd + q * 256, outPtr + q * outStride, acc);
                push(acc);
            }
        }
        else
        {
            // Blocks with a non-canonical step (block seams: the own-lane term comes first in the reference order and
            // the skew of the lanes spreads one seam over 62 steps, i.e. 9 blocks per seam of every group): the terms
            // stay in reference order and are selected by their descriptor bytes, but the operands of HW steps are
            // loaded at once and - when every shuffled term of the step comes from the linked neighbour lane (meta
            // bit 57, set on the host) - a single shuffle is issued ahead of the dependent chain, as in the canonical path.
            nGeneral++;
            constexpr int HW = RG <= 3 ? 4 : 2;
#pragma unroll 1
            for (int q0 = 0; q0 < kNH; q0 += HW)
            {
                unsigned long long meta[HW];
                double a0[HW], cv0[HW], cv1[HW], cf[HW][RG];
#pragma unroll
                for (int u = 0; u < HW; u++)
                {
                    const int q = q0 + u;
                    meta[u] = *reinterpret_cast<const unsigned long long*>(sb + q * 256);
                    a0[u] = *reinterpret_cast<const double*>(hd + q * 256);
                    cv0[u] = Kg > 0 ? *reinterpret_cast<const double*>(hd + PL + q * 256) : 0.0;
                    cv1[u] = Kg > 1 ? *reinterpret_cast<const double*>(hd + 2 * PL + q * 256) : 0.0;
#pragma unroll
                    for (int r = 0; r < RG; r++) cf[u][r] = *reinterpret_cast<const double*>(sb + (1 + r) * PL + q * 256);
                }
#pragma unroll
                for (int u = 0; u < HW; u++)
                {
                    double acc = a0[u];
                    // the linked-lane shuffle is issued by all lanes before any branch (a dual step splits the lanes)
                    const double sh = __shfl_sync(FULL, h[kSkew - 1], srcLane);
                    if ((meta[u] >> 58) & 1ull)
                    { // seam-form lane of a dual step (only calcReciprocalD and the debug mode get here): planes 1, 2, 0
                        const unsigned selB = (unsigned)(meta[u] >> 59) & 3u;
                        acc = sweep_apply<MODE>(acc, cf[u][1], h[0]);
                        if (RG > 2 && !((meta[u] >> 16) & kMetaPad)) acc = sweep_apply<MODE>(acc, cf[u][RG > 2 ? 2 : 0], cv0[u]);
                        if (!(meta[u] & kMetaPad)) acc = sweep_apply<MODE>(acc, cf[u][0], selB == 2u ? cv1[u] : (selB == 1u ? cv0[u] : sh));
                    }
                    else if ((meta[u] >> 57) & 1ull)
                    { // all shuffled terms of this step read the linked lane (dual steps always do)
#pragma unroll
                        for (int r = 0; r < RG; r++)
                        {
                            const unsigned byte = (unsigned)(meta[u] >> (8 * r)) & 0xffu;
                            const double v = (byte & kMetaConst) ? ((byte & 0x20u) ? cv1[u] : cv0[u]) : ((byte & kMetaOwn) ? h[0] : sh);
                            // padding terms are skipped: the reference has no such term, and calcReciprocalD's division is
                            // ~100 instructions that a plane without a real term in any lane then never issues
                            if (!(byte & kMetaPad)) acc = sweep_apply<MODE>(acc, cf[u][r], v);
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int r = 0; r < RG; r++)
                        {
                            const unsigned byte = (unsigned)(meta[u] >> (8 * r)) & 0xffu;
                            const double shr = __shfl_sync(FULL, h[kSkew - 1], (int)(byte & kMetaLane));
                            const double v = (byte & kMetaConst) ? ((byte & 0x20u) ? cv1[u] : cv0[u]) : ((byte & kMetaOwn) ? h[0] : shr);
                            if (!(byte & kMetaPad)) acc = sweep_apply<MODE>(acc, cf[u][r], v);
                        }
                    }
                    __syncwarp();
                    put(hd + (q0 + u) * 256, outPtr + (q0 + u) * outStride, acc);
                    push(acc);
                }
            }
        }
        outPtr += kNH * outStride;
        __syncwarp();
        if (statsOn)
        {
            const long long d = clock64() - b0;
            if (canon)
            {
                tCanon += d;
                nCanon++;
            }
            else
                tOther += d;
        }
        if (lane == 0)
        { // block done: the stage may be refilled
            st_flag_smem(cntp, 0u);
            st_flag_smem(C.done, (unsigned)(blk + 1));
            if (STATS == 1 && blk < kTraceBlocks) S.stats[(long long)kStatsStride * g + 16 + blk * 8 + 1] = clock64();
        }
        sb += C.stageBytes;
        cntp++;
        if (sb == stageEnd)
        {
            sb = stage0;
            cntp = C.cnt;
        }
    }
    if (statsOn && lane == 0)
    {
        long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        long long* sp = S.stats + (long long)kStatsStride * g;
        sp[0] = clock64() - t0;
        sp[1] = tWait;
        sp[2] = g0;
        sp[3] = g1;
        sp[5] = nT;
        sp[6] = nGeneral;
        sp[7] = C.nBlocks;
        if (STATS == 2)
        { // the per-block trace area is unused in this instantiation: body cycles of canonical / other blocks
            sp[16] = tCanon;
            sp[17] = nCanon;
            sp[18] = tOther;
        }
    }
}

// ------------------------------------------------------------------------------------------ producers
// A block is ready when its slowest producer has delivered, and a producer's step is one long chain of dependent
// shared-memory loads, checks and flag operations executed by a single warp: measured (per-block time stamps), that
// chain - not the memory system and not the consumer - set the pace of a group.  The loop is therefore written for a
// short chain: 32-bit shared-memory addresses stepped from stage to stage (no 64-bit address arithmetic), the
// operand loads of the current block issued before the wait for the next stage so that their latency overlaps it,
// no instrumentation in the product instantiation (STATS = false).

template <int MODE, int LG, int STATS>
__device__ __forceinline__ void split_producer(const PipeDev& S, const SplitCtx& C, const int g, const int h, const int lane, double* out,
                                               int* err)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int LGA = LG > 0 ? LG : 1;
    constexpr int dir = MODE == 1 ? -1 : 1;
    const int vecIdx = (dir > 0 ? h : kNH - 1 - h) * 32 + lane; // this producer's step inside the a / b chunk
    const double neutral = MODE == 2 ? 1.0 : 0.0;
    const int Kg = C.Kg;
    // byte offsets of this producer's (step h) and this lane's operands inside a stage
    const unsigned offCoef = (unsigned)(C.offP + h * C.pRec + lane * 8);              // coef i at + i * 256
    const unsigned offCode = (unsigned)(C.offP + h * C.pRec + LG * 256 + lane * 4);   // code i at + i * 128
    const unsigned offConst = (unsigned)(C.offP + h * C.pRec + LG * 384 + lane * 4);  // const code k at + k * 128
    const unsigned offA = (unsigned)(C.offA + vecIdx * 8);                            // a; b at + kNH * 256
    const unsigned offGen = (unsigned)(h * 256 + lane * 8 + 7);                       // top byte of the step's meta word
    const unsigned offHd = (unsigned)(C.offHdr + h * 256 + lane * 8);                 // hdr planes acc0 | cval 0 | cval 1
    const unsigned stage0 = smem_u32(C.stages), stageBytes = (unsigned)C.stageBytes;
    const unsigned stageEnd = stage0 + (unsigned)C.NS * stageBytes;
    const unsigned bar0 = smem_u32(C.rawBar), barEnd = bar0 + 8u * (unsigned)C.NS;
    const unsigned cnt0 = smem_u32(C.cnt);
    // two register sets for the prefetched codes / cross-group values: the block loop is unrolled by two and the
    // sets swap roles, so that no register copy (which would wait for the prefetch to land) sits in the loop
    int codesA[LGA], codesB[LGA], ccA[2], ccB[2];
    double mvA[LGA], mvB[LGA], mcA[2], mcB[2];
    // codes of a step's cross-group terms (from the landed stage), and the loads of their values: issued separately,
    // see the order of a step below
    auto fetch_codes = [&](const unsigned sg, int* cd, int* kc) {
#pragma unroll
        for (int i = 0; i < LG; i++) cd[i] = lds_s32(sg + offCode + i * 128);
#pragma unroll
        for (int k = 0; k < 2; k++) kc[k] = (k < Kg) ? lds_s32(sg + offConst + k * 128) : -1;
    };
    auto fetch_values = [&](const int* cd, double* vals, const int* kc, double* kv) {
#pragma unroll
        for (int i = 0; i < LG; i++)
        {
            vals[i] = neutral;
            if (cd[i] >= 0) vals[i] = ld_relaxed(out + cd[i]);
        }
#pragma unroll
        for (int k = 0; k < 2; k++)
        {
            kv[k] = 0.0;
            if (kc[k] >= 0) kv[k] = ld_relaxed(out + kc[k]);
        }
    };
    unsigned sg = stage0, bar = bar0, cnt = cnt0, par = 0u;
    const bool timed = (STATS != 0) && h == 0;
    long long tStage = 0, tVal = 0, tSpin = 0, tAll = timed ? clock64() : 0;
    auto step = [&](const int blk, int* codes, double* mv, int* cc, double* mc, int* codesN, double* mvN, int* ccN, double* mcN) {
        long long* const tr = (STATS == 1 && lane == 0 && blk < kTraceBlocks) ? S.stats + (long long)kStatsStride * g + 16 + blk * 8 : nullptr;
        if (STATS == 1 && tr && h == 0) tr[3] = clock64();
        // ---- operands of the current step (its stage landed: waited for one block ago)
        double acc = lds_f64(sg + offA);
        double bb = 0.0;
        if (MODE == 0) bb = lds_f64(sg + offA + kNH * 256);
        double cf[LGA];
#pragma unroll
        for (int i = 0; i < LG; i++) cf[i] = lds_f64(sg + offCoef + i * 256);
        const unsigned b7 = lds_u8(sg + offGen); // bit 0: descriptor-driven step, bit 5: dual step (schedule.hpp)
        // what the consumer learns with the arrival: bits 8-15 count the descriptor-driven steps, bit 16 + h says that
        // step h is a dual step
        const unsigned stepKind = MODE == 2 ? 0x100u : (((b7 & 1u) << 8) | (((b7 >> 5) & 1u) << (16 + h)));
        // ---- prefetch the codes / cross-group values of this producer's step in the next block
        unsigned sgN = sg + stageBytes, barN = bar + 8u, parN = par;
        if (sgN == stageEnd)
        {
            sgN = stage0;
            barN = bar0;
            parN ^= 1u;
        }
        const bool haveNext = blk + 1 < C.nBlocks;
        if (haveNext)
        {
            const long long w0 = timed ? clock64() : 0;
            mbar_wait_u32(barN, parN);
            if (timed) tStage += clock64() - w0;
            if (STATS == 1 && tr && h == 0) tr[4] = clock64();
            fetch_codes(sgN, codesN, ccN);
        }
        if (MODE == 0) acc *= bb;
        // ---- the (possibly late) cross-group values
        const long long v0 = timed ? clock64() : 0;
        bool bad = false;
#pragma unroll
        for (int i = 0; i < LG; i++) bad |= codes[i] >= 0 && is_sentinel(mv[i]);
#pragma unroll
        for (int k = 0; k < 2; k++) bad |= cc[k] >= 0 && is_sentinel(mc[k]);
        const bool anyBad = __any_sync(FULL, bad);
        const long long v1 = timed ? clock64() : 0;
        if (STATS == 1 && tr && h == 0) tr[5] = v1;
        if (anyBad)
        { // a value had not arrived when it was prefetched: poll for it.  A group that runs right behind the group
          // it depends on gets here in every block; a polling round re-reads, in one batch of independent loads, whatever
          // is still missing of this block.
            if (STATS == 1 && lane == 0) atomicAdd((unsigned long long*)(S.stats + (long long)kStatsStride * g + 4), 1ull);
            int tries = 0;
            bool still;
            do
            {
#pragma unroll
                for (int i = 0; i < LG; i++)
                    if (codes[i] >= 0 && is_sentinel(mv[i])) mv[i] = ld_relaxed(out + codes[i]);
#pragma unroll
                for (int k = 0; k < 2; k++)
                    if (cc[k] >= 0 && is_sentinel(mc[k])) mc[k] = ld_relaxed(out + cc[k]);
                still = false;
#pragma unroll
                for (int i = 0; i < LG; i++) still |= codes[i] >= 0 && is_sentinel(mv[i]);
#pragma unroll
                for (int k = 0; k < 2; k++) still |= cc[k] >= 0 && is_sentinel(mc[k]);
                if (++tries >= 64)
                {
                    __nanosleep(tries > 4096 ? 400 : 64);
                    if ((tries & 4095) == 4095 && *(volatile int*)err) break;
                    if (tries >= S.spinLimit)
                    {
                        atomicExch(err, 1);
                        break;
                    }
                }
            } while (__any_sync(FULL, still));
        }
        if (timed)
        {
            tVal += v1 - v0;
            tSpin += clock64() - v1;
            if (blk == C.nBlocks / 2 && lane == 0)
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(S.stats[(long long)kStatsStride * g + 14]));
        }
        // ---- only now the loads of the next block's cross-group values.  ptxas tracks all of these loads with ONE
        // scoreboard (checked in the SASS), so a check of older values also waits for whatever was issued since: issued
        // here - after the check, before the hand-over - the loads are the only ones outstanding at the next check, and
        // their flight overlaps the hand-over, the loop back, the operand loads and the wait for the stage after next.
        if (haveNext) fetch_values(codesN, mvN, ccN, mcN);
#pragma unroll
        for (int i = 0; i < LG; i++) acc = sweep_apply<MODE>(acc, cf[i], mv[i]); // padding: coefficient 0, neutral value
        // ---- hand over: the hdr part of the stage is free because the stage was refilled after the consumer released it
        sts_f64(sg + offHd, acc);
        if (Kg > 0) sts_f64(sg + offHd + kNH * 256, mc[0]);
        if (Kg > 1) sts_f64(sg + offHd + 2 * kNH * 256, mc[1]);
        __syncwarp();
        if (lane == 0) asm volatile("red.relaxed.cta.shared.add.u32 [%0], %1;" ::"r"(cnt), "r"(1u + stepKind) : "memory");
        if (STATS == 1 && tr && h == 0) tr[6] = clock64();
        if (STATS == 1 && tr && h == kNH - 1) tr[7] = clock64();
        cnt = (sgN == stage0) ? cnt0 : cnt + 4u;
        sg = sgN;
        bar = barN;
        par = parN;
    };
    (void)barEnd;
    mbar_wait_u32(bar0, 0u);
    fetch_codes(stage0, codesA, ccA);
    fetch_values(codesA, mvA, ccA, mcA);
    for (int blk = 0; blk < C.nBlocks; blk += 2)
    {
        step(blk, codesA, mvA, ccA, mcA, codesB, mvB, ccB, mcB);
        if (blk + 1 < C.nBlocks) step(blk + 1, codesB, mvB, ccB, mcB, codesA, mvA, ccA, mcA);
    }
    if (timed && lane == 0)
    {
        long long* sp = S.stats + (long long)kStatsStride * g;
        sp[8] = clock64() - tAll;
        sp[9] = tStage;
        sp[10] = tVal;
        sp[11] = tSpin;
    }
}

// Producer of the set organisation: this warp prepares steps h0 .. h0 + M - 1 of the blocks set, set + nSets, ...
template <int MODE, int LG, int M, int STATS>
__device__ __forceinline__ void split_producer_sets(const PipeDev& S, const SplitCtx& C, const int g, const int set, const int nSets, const int h0,
                                                    const int lane, double* out, int* err, const double* __restrict__ aVec,
                                                    const double* __restrict__ bVec)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int LGA = LG > 0 ? LG : 1;
    constexpr int dir = MODE == 1 ? -1 : 1;
    const double neutral = MODE == 2 ? 1.0 : 0.0;
    const int Kg = C.Kg;
    // byte offsets of step h0's and this lane's operands inside a stage; step h0 + u at + u * (the step's stride)
    const unsigned offCoef = (unsigned)(C.offP + h0 * C.pRec + lane * 8);             // coef i at + i * 256
    const unsigned offCode = (unsigned)(C.offP + h0 * C.pRec + LG * 256 + lane * 4);  // code i at + i * 128
    const unsigned offConst = (unsigned)(C.offP + h0 * C.pRec + LG * 384 + lane * 4); // const code k at + k * 128
    const unsigned pRec = (unsigned)C.pRec;
    const int vec0 = (dir > 0 ? h0 : kNH - 1 - h0) * 32 + lane; // step h0 inside the a / b chunk; step h0 + u at + u * dir * 32
    const unsigned offA = (unsigned)(C.offA + vec0 * 8);
    const int vecStep = dir * 256;
    const unsigned offGen = (unsigned)(h0 * 256 + lane * 8 + 7); // top byte of the step's meta word
    const unsigned offHd = (unsigned)(C.offHdr + h0 * 256 + lane * 8); // hdr planes acc0 | cval 0 | cval 1
    const unsigned stage0 = smem_u32(C.stages), stageBytes = (unsigned)C.stageBytes;
    const unsigned bar0 = smem_u32(C.rawBar), cnt0 = smem_u32(C.cnt);
    const int NS = C.NS;
    int st = set % NS;
    unsigned par = (unsigned)(set / NS) & 1u;
    // direct mode: this lane's operands of step h0 of block 0 in global memory
    const unsigned char* const pG0 = C.pStream + (size_t)h0 * C.pRec;
    const double* const aG0 = aVec + ((long long)C.base + (dir > 0 ? h0 : C.nT - 1 - h0)) * 32 + lane;
    const double* const bG0 = bVec ? bVec + ((long long)C.base + (dir > 0 ? h0 : C.nT - 1 - h0)) * 32 + lane : nullptr;
    const bool timed = (STATS != 0) && set == 0 && h0 == 0;
    long long tStage = 0, tVal = 0, tSpin = 0, tAll = timed ? clock64() : 0;
    for (int blk = set; blk < C.nBlocks; blk += nSets)
    {
        const unsigned sg = stage0 + (unsigned)st * stageBytes;
        long long* const tr = (STATS == 1 && lane == 0 && blk < kTraceBlocks) ? S.stats + (long long)kStatsStride * g + 16 + blk * 8 : nullptr;
        if (STATS == 1 && tr && h0 == 0) tr[3] = clock64();
        // ---- direct mode: the operands of this block from global memory, issued before anything is waited for
        int cd[M][LGA], kc[M][2];
        double mv[M][LGA], mc[M][2];
        double acc[M], bbv[M], cfv[M][LGA];
        if (kProdDirect)
        {
            const unsigned char* pG = pG0 + (size_t)blk * C.pBytes;
            const long long vOff = (long long)blk * kNH * 32 * dir;
#pragma unroll
            for (int u = 0; u < M; u++)
            {
#pragma unroll
                for (int i = 0; i < LG; i++) cd[u][i] = ldg_stream_s32(pG + u * pRec + LG * 256 + i * 128 + lane * 4);
#pragma unroll
                for (int k = 0; k < 2; k++) kc[u][k] = (k < Kg) ? ldg_stream_s32(pG + u * pRec + LG * 384 + k * 128 + lane * 4) : -1;
            }
#pragma unroll
            for (int u = 0; u < M; u++)
            {
                acc[u] = ldg_stream_f64(aG0 + vOff + u * dir * 32);
                bbv[u] = MODE == 0 ? ldg_stream_f64(bG0 + vOff + u * dir * 32) : 0.0;
#pragma unroll
                for (int i = 0; i < LG; i++) cfv[u][i] = ldg_stream_f64(pG + u * pRec + i * 256 + lane * 8);
            }
            // the cross-group values as soon as their codes are here
#pragma unroll
            for (int u = 0; u < M; u++)
            {
#pragma unroll
                for (int i = 0; i < LG; i++)
                {
                    mv[u][i] = neutral;
                    if (cd[u][i] >= 0) mv[u][i] = ld_relaxed(out + cd[u][i]);
                }
#pragma unroll
                for (int k = 0; k < 2; k++)
                {
                    mc[u][k] = 0.0;
                    if (kc[u][k] >= 0) mc[u][k] = ld_relaxed(out + kc[u][k]);
                }
            }
        }
        const long long w0 = timed ? clock64() : 0;
        // A parity wait only tells the last two phases of a barrier apart, and bulk copies of different stages complete
        // out of order.  The host therefore makes the stage count a multiple of the set count (upload_pipe_dir): every
        // fill of a stage is then waited for by the same set, one after the other, and "phase par complete" cannot be
        // answered for an older fill.
        mbar_wait_u32(bar0 + 8u * (unsigned)st, par);
        if (timed) tStage += clock64() - w0;
        if (STATS == 1 && tr && h0 == 0) tr[4] = clock64();
        // ---- ring mode: codes of the cross-group terms, then their loads: the only global accesses of the step
        if (!kProdDirect)
        {
#pragma unroll
            for (int u = 0; u < M; u++)
            {
#pragma unroll
                for (int i = 0; i < LG; i++) cd[u][i] = lds_s32(sg + offCode + u * pRec + i * 128);
#pragma unroll
                for (int k = 0; k < 2; k++) kc[u][k] = (k < Kg) ? lds_s32(sg + offConst + u * pRec + k * 128) : -1;
            }
#pragma unroll
            for (int u = 0; u < M; u++)
            {
#pragma unroll
                for (int i = 0; i < LG; i++)
                {
                    mv[u][i] = neutral;
                    if (cd[u][i] >= 0) mv[u][i] = ld_relaxed(out + cd[u][i]);
                }
#pragma unroll
                for (int k = 0; k < 2; k++)
                {
                    mc[u][k] = 0.0;
                    if (kc[u][k] >= 0) mc[u][k] = ld_relaxed(out + kc[u][k]);
                }
            }
        }
        // ---- operands from the stage while the loads fly
        unsigned kind = 0u;
#pragma unroll
        for (int u = 0; u < M; u++)
        {
            if (!kProdDirect)
            {
                acc[u] = lds_f64(offA + sg + u * vecStep);
                bbv[u] = MODE == 0 ? lds_f64(offA + sg + u * vecStep + kNH * 256) : 0.0;
            }
            if (MODE == 0) acc[u] *= bbv[u];
            const unsigned b7 = lds_u8(sg + offGen + u * 256); // bit 0: descriptor-driven step, bit 5: dual step (schedule.hpp)
            kind += MODE == 2 ? 0x100u : (((b7 & 1u) << 8) | (((b7 >> 5) & 1u) << (16 + h0 + u)));
        }
        // ---- the cross-group values: poll for whatever had not been written yet
        const long long v0 = timed ? clock64() : 0;
        bool bad = false;
#pragma unroll
        for (int u = 0; u < M; u++)
        {
#pragma unroll
            for (int i = 0; i < LG; i++) bad |= cd[u][i] >= 0 && is_sentinel(mv[u][i]);
#pragma unroll
            for (int k = 0; k < 2; k++) bad |= kc[u][k] >= 0 && is_sentinel(mc[u][k]);
        }
        const bool anyBad = __any_sync(FULL, bad);
        const long long v1 = timed ? clock64() : 0;
        if (STATS == 1 && tr && h0 == 0) tr[5] = v1;
        if (anyBad)
        {
            if (STATS == 1 && lane == 0) atomicAdd((unsigned long long*)(S.stats + (long long)kStatsStride * g + 4), 1ull);
            int tries = 0;
            bool still;
            do
            {
#pragma unroll
                for (int u = 0; u < M; u++)
                {
#pragma unroll
                    for (int i = 0; i < LG; i++)
                        if (cd[u][i] >= 0 && is_sentinel(mv[u][i])) mv[u][i] = ld_relaxed(out + cd[u][i]);
#pragma unroll
                    for (int k = 0; k < 2; k++)
                        if (kc[u][k] >= 0 && is_sentinel(mc[u][k])) mc[u][k] = ld_relaxed(out + kc[u][k]);
                }
                still = false;
#pragma unroll
                for (int u = 0; u < M; u++)
                {
#pragma unroll
                    for (int i = 0; i < LG; i++) still |= cd[u][i] >= 0 && is_sentinel(mv[u][i]);
#pragma unroll
                    for (int k = 0; k < 2; k++) still |= kc[u][k] >= 0 && is_sentinel(mc[u][k]);
                }
                if (++tries >= 64)
                {
                    __nanosleep(tries > 4096 ? 400 : 64);
                    if ((tries & 4095) == 4095 && *(volatile int*)err) break;
                    if (tries >= S.spinLimit)
                    {
                        atomicExch(err, 1);
                        break;
                    }
                }
            } while (__any_sync(FULL, still));
        }
        if (timed)
        {
            tVal += v1 - v0;
            tSpin += clock64() - v1;
            if (blk == (C.nBlocks / 2 / nSets) * nSets && lane == 0)
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(S.stats[(long long)kStatsStride * g + 14]));
        }
        // ---- leading terms in reference order, hand-over
#pragma unroll
        for (int u = 0; u < M; u++)
        {
#pragma unroll
            for (int i = 0; i < LG; i++)
                acc[u] = sweep_apply<MODE>(acc[u], kProdDirect ? cfv[u][i] : lds_f64(sg + offCoef + u * pRec + i * 256), mv[u][i]); // padding: coefficient 0, neutral value
            sts_f64(sg + offHd + u * 256, acc[u]);
            if (Kg > 0) sts_f64(sg + offHd + u * 256 + kNH * 256, mc[u][0]);
            if (Kg > 1) sts_f64(sg + offHd + u * 256 + 2 * kNH * 256, mc[u][1]);
        }
        __syncwarp();
        if (lane == 0) asm volatile("red.relaxed.cta.shared.add.u32 [%0], %1;" ::"r"(cnt0 + 4u * (unsigned)st), "r"((unsigned)M + kind) : "memory");
        if (STATS == 1 && tr && h0 == 0) tr[6] = clock64();
        if (STATS == 1 && tr && h0 + M == kNH) tr[7] = clock64();
        st += nSets;
        while (st >= NS)
        {
            st -= NS;
            par ^= 1u;
        }
    }
    if (timed && lane == 0)
    {
        long long* sp = S.stats + (long long)kStatsStride * g;
        sp[8] = clock64() - tAll;
        sp[9] = tStage;
        sp[10] = tVal;
        sp[11] = tSpin;
    }
}

// Software-pipelined producer of the direct mode (B200_PROD_PIPE): the static operands of a block (codes, coefficients,
// a, b - always available) are requested one producer iteration ahead, so that the loads of the cross-group values - the
// only ones that depend on another group's progress - go out as soon as this warp is done with its previous block, as in
// the single-set producer (split_producer), and the two global round trips of a block do not add up on the path between
// "the value exists" and "the block is handed to the consumer".
template <int MODE, int LG, int M, int STATS>
__device__ __forceinline__ void split_producer_pipe(const PipeDev& S, const SplitCtx& C, const int g, const int set, const int nSets, const int h0,
                                                    const int lane, double* out, int* err, const double* __restrict__ aVec,
                                                    const double* __restrict__ bVec)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int LGA = LG > 0 ? LG : 1;
    constexpr int dir = MODE == 1 ? -1 : 1;
    const double neutral = MODE == 2 ? 1.0 : 0.0;
    const int Kg = C.Kg;
    const unsigned pRec = (unsigned)C.pRec;
    const unsigned offGen = (unsigned)(h0 * 256 + lane * 8 + 7);       // top byte of the step's meta word
    const unsigned offHd = (unsigned)(C.offHdr + h0 * 256 + lane * 8); // hdr planes acc0 | cval 0 | cval 1
    const unsigned stage0 = smem_u32(C.stages), stageBytes = (unsigned)C.stageBytes;
    const unsigned bar0 = smem_u32(C.rawBar), cnt0 = smem_u32(C.cnt);
    const int NS = C.NS, nBlocks = C.nBlocks;
    const unsigned char* const pG0 = C.pStream + (size_t)h0 * C.pRec;
    const double* const aG0 = aVec + ((long long)C.base + (dir > 0 ? h0 : C.nT - 1 - h0)) * 32 + lane;
    const double* const bG0 = bVec ? bVec + ((long long)C.base + (dir > 0 ? h0 : C.nT - 1 - h0)) * 32 + lane : nullptr;
    const bool timed = (STATS != 0) && set == 0 && h0 == 0;
    long long tStage = 0, tVal = 0, tSpin = 0, tAll = timed ? clock64() : 0;
    struct Ops
    {
        int cd[M][LGA], kc[M][2];
        double acc[M], bb[M], cf[M][LGA];
    };
    double mv[M][LGA], mc[M][2]; // cross-group values of the block in work, then of the next one
    auto load_static = [&](const int blk, Ops& o) {
        if (blk >= nBlocks) return;
        const unsigned char* pG = pG0 + (size_t)blk * C.pBytes;
        const long long vOff = (long long)blk * kNH * 32 * dir;
#pragma unroll
        for (int u = 0; u < M; u++)
        {
#pragma unroll
            for (int i = 0; i < LG; i++) o.cd[u][i] = ldg_stream_s32(pG + u * pRec + LG * 256 + i * 128 + lane * 4);
#pragma unroll
            for (int k = 0; k < 2; k++) o.kc[u][k] = (k < Kg) ? ldg_stream_s32(pG + u * pRec + LG * 384 + k * 128 + lane * 4) : -1;
            o.acc[u] = ldg_stream_f64(aG0 + vOff + u * dir * 32);
            o.bb[u] = MODE == 0 ? ldg_stream_f64(bG0 + vOff + u * dir * 32) : 0.0;
#pragma unroll
            for (int i = 0; i < LG; i++) o.cf[u][i] = ldg_stream_f64(pG + u * pRec + i * 256 + lane * 8);
        }
    };
    auto load_values = [&](const Ops& o) {
#pragma unroll
        for (int u = 0; u < M; u++)
        {
#pragma unroll
            for (int i = 0; i < LG; i++)
            {
                mv[u][i] = neutral;
                if (o.cd[u][i] >= 0) mv[u][i] = ld_relaxed(out + o.cd[u][i]);
            }
#pragma unroll
            for (int k = 0; k < 2; k++)
            {
                mc[u][k] = 0.0;
                if (o.kc[u][k] >= 0) mc[u][k] = ld_relaxed(out + o.kc[u][k]);
            }
        }
    };
    int st = set % NS;
    unsigned par = (unsigned)(set / NS) & 1u;
    auto work = [&](const int blk, Ops& cur, const Ops& nxt) {
        const unsigned sg = stage0 + (unsigned)st * stageBytes;
        const long long w0 = timed ? clock64() : 0;
        mbar_wait_u32(bar0 + 8u * (unsigned)st, par); // the stage of this block (its hdr part is free, its meta words are here)
        if (timed) tStage += clock64() - w0;
        unsigned kind = 0u;
#pragma unroll
        for (int u = 0; u < M; u++)
        {
            const unsigned b7 = lds_u8(sg + offGen + u * 256); // bit 0: descriptor-driven step, bit 5: dual step (schedule.hpp)
            kind += MODE == 2 ? 0x100u : (((b7 & 1u) << 8) | (((b7 >> 5) & 1u) << (16 + h0 + u)));
            if (MODE == 0) cur.acc[u] *= cur.bb[u];
        }
        // ---- the cross-group values (requested when the previous block of this warp was handed over)
        const long long v0 = timed ? clock64() : 0;
        bool bad = false;
#pragma unroll
        for (int u = 0; u < M; u++)
        {
#pragma unroll
            for (int i = 0; i < LG; i++) bad |= cur.cd[u][i] >= 0 && is_sentinel(mv[u][i]);
#pragma unroll
            for (int k = 0; k < 2; k++) bad |= cur.kc[u][k] >= 0 && is_sentinel(mc[u][k]);
        }
        const bool anyBad = __any_sync(FULL, bad);
        const long long v1 = timed ? clock64() : 0;
        if (anyBad)
        {
            if (STATS == 1 && lane == 0) atomicAdd((unsigned long long*)(S.stats + (long long)kStatsStride * g + 4), 1ull);
            int tries = 0;
            bool still;
            do
            {
#pragma unroll
                for (int u = 0; u < M; u++)
                {
#pragma unroll
                    for (int i = 0; i < LG; i++)
                        if (cur.cd[u][i] >= 0 && is_sentinel(mv[u][i])) mv[u][i] = ld_relaxed(out + cur.cd[u][i]);
#pragma unroll
                    for (int k = 0; k < 2; k++)
                        if (cur.kc[u][k] >= 0 && is_sentinel(mc[u][k])) mc[u][k] = ld_relaxed(out + cur.kc[u][k]);
                }
                still = false;
#pragma unroll
                for (int u = 0; u < M; u++)
                {
#pragma unroll
                    for (int i = 0; i < LG; i++) still |= cur.cd[u][i] >= 0 && is_sentinel(mv[u][i]);
#pragma unroll
                    for (int k = 0; k < 2; k++) still |= cur.kc[u][k] >= 0 && is_sentinel(mc[u][k]);
                }
                if (++tries >= 64)
                {
                    __nanosleep(tries > 4096 ? 400 : 64);
                    if ((tries & 4095) == 4095 && *(volatile int*)err) break;
                    if (tries >= S.spinLimit)
                    {
                        atomicExch(err, 1);
                        break;
                    }
                }
            } while (__any_sync(FULL, still));
        }
        if (timed)
        {
            tVal += v1 - v0;
            tSpin += clock64() - v1;
        }
        // ---- leading terms in reference order, hand-over
#pragma unroll
        for (int u = 0; u < M; u++)
        {
#pragma unroll
            for (int i = 0; i < LG; i++) cur.acc[u] = sweep_apply<MODE>(cur.acc[u], cur.cf[u][i], mv[u][i]); // padding: coefficient 0, neutral value
            sts_f64(sg + offHd + u * 256, cur.acc[u]);
            if (Kg > 0) sts_f64(sg + offHd + u * 256 + kNH * 256, mc[u][0]);
            if (Kg > 1) sts_f64(sg + offHd + u * 256 + 2 * kNH * 256, mc[u][1]);
        }
        __syncwarp();
        if (lane == 0) asm volatile("red.relaxed.cta.shared.add.u32 [%0], %1;" ::"r"(cnt0 + 4u * (unsigned)st), "r"((unsigned)M + kind) : "memory");
        // ---- the next block of this warp: its values now (its codes came in during this block), its successor's static
        // operands into the registers that have just become free
        if (blk + nSets < nBlocks) load_values(nxt);
        load_static(blk + 2 * nSets, cur);
        st += nSets;
        while (st >= NS)
        {
            st -= NS;
            par ^= 1u;
        }
    };
    Ops X, Y;
    load_static(set, X);
    load_static(set + nSets, Y);
    if (set < nBlocks) load_values(X);
    for (int blk = set; blk < nBlocks; blk += 2 * nSets)
    {
        work(blk, X, Y);
        if (blk + nSets < nBlocks) work(blk + nSets, Y, X);
    }
    if (timed && lane == 0)
    {
        long long* sp = S.stats + (long long)kStatsStride * g;
        sp[8] = clock64() - tAll;
        sp[9] = tStage;
        sp[10] = tVal;
        sp[11] = tSpin;
    }
}

// Warp roles: consumer, kNH producers, loader.  (Measured on B200: giving the consumer a scheduler sub-partition of its
// own - wid % 4, 11 warps - or the highest warp id of the CTA changes the sweep time by < 2 %: the consumer is not
// short of issue slots.)
constexpr int kRoleConsumer = -1, kRoleLoader = -2;
constexpr int kSweepWarps = 2 + kProducers;
__device__ __forceinline__ int sweep_role(int warp)
{
    if (warp == 0) return kRoleConsumer;
    if (warp == 1 + kProducers) return kRoleLoader;
    return warp - 1;
}

template <int MODE, int STATS>
__device__ __forceinline__ void sweep_group_split(const PipeDev& S, const int g, const int warp, const int lane, const double* __restrict__ a,
                                                  const double* __restrict__ b, double* out, int* err, unsigned char* smem)
{
    SplitCtx C;
    C.nT = S.gNT[g];
    C.base = S.gBase[g];
    C.nBlocks = C.nT / kNH;
    C.NS = S.nStages;
    C.stageBytes = S.stageBytes;
    C.Lg = S.gLg[g];
    C.Rg = S.gRg[g];
    C.Kg = S.gKg[g];
    C.pRec = C.Lg * 384 + C.Kg * 128;
    C.cRec = 256 + C.Rg * 256;
    C.hdrStep = 256 * (1 + C.Kg);
    C.pBytes = (unsigned)(kNH * C.pRec);
    C.cBytes = (unsigned)(kNH * C.cRec);
    C.offHdr = (int)C.cBytes;
    C.offP = C.offHdr + kNH * C.hdrStep;
    C.offA = C.offP + (int)C.pBytes;
    C.rawBytes = kProdDirect ? C.cBytes : C.pBytes + C.cBytes + (unsigned)(kNH * 256 * (MODE == 0 ? 2 : 1));
    // header of kSweepSmemHeader bytes: rawBar[16] | cnt[16] | done | (debug) issueClk[16]
    C.rawBar = reinterpret_cast<unsigned long long*>(smem);
    C.cnt = reinterpret_cast<unsigned*>(smem + 128);
    C.done = reinterpret_cast<unsigned*>(smem + 192);
    C.stages = smem + kSweepSmemHeader;
    C.pStream = S.pStream + S.gPOff[g];
    C.cStream = S.cStream + S.gCOff[g];
    C.polStream = l2_evict_first();
    if (threadIdx.x == 0)
    {
        for (int st = 0; st < C.NS; st++)
        {
            mbar_init(&C.rawBar[st], 1);
            C.cnt[st] = 0u;
        }
        *C.done = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int c = 0; c < C.NS && c < C.nBlocks; c++) split_issue<MODE>(C, a, b, c, c);
        for (int c = C.NS; c < C.NS + S.l2Ahead; c++) split_prefetch<MODE>(C, a, b, c);
    }
    __syncthreads();
    const int role = sweep_role(warp);
    if (role == kRoleConsumer)
    {
        switch (C.Rg)
        {
            case 2: split_consumer<MODE, 2, STATS>(S, C, g, lane, out); break;
            case 3: split_consumer<MODE, 3, STATS>(S, C, g, lane, out); break;
            case 4: split_consumer<MODE, 4, STATS>(S, C, g, lane, out); break;
            case 5: split_consumer<MODE, 5, STATS>(S, C, g, lane, out); break;
            default: split_consumer<MODE, 6, STATS>(S, C, g, lane, out); break;
        }
    }
    else if (role == kRoleLoader)
    {
        if (kStoreSmem)
        {
            // loader + storer: the whole warp writes a finished block from the hdr plane of its stage to the output vector
            // (one coalesced store per step), then lane 0 refills the stage
            constexpr int dir = MODE == 1 ? -1 : 1;
            int st = 0;
            double* op = out + ((long long)C.base + (dir > 0 ? 0 : C.nT - 1)) * 32 + lane;
            for (int blk = 0; blk < C.nBlocks; blk++)
            {
                const unsigned need = (unsigned)(blk + 1);
                while (ld_flag_smem(C.done) < need) __nanosleep(20);
                const unsigned hdA = smem_u32(C.stages) + (unsigned)st * (unsigned)C.stageBytes + (unsigned)C.offHdr + (unsigned)lane * 8u;
                double v[kNH];
#pragma unroll
                for (int q = 0; q < kNH; q++) v[q] = lds_f64(hdA + q * 256);
#pragma unroll
                for (int q = 0; q < kNH; q++) st_relaxed(op + (long long)q * dir * 32, v[q]);
                op += (long long)kNH * dir * 32;
                __syncwarp();
                if (lane == 0 && blk + C.NS < C.nBlocks)
                {
                    if (STATS == 1 && blk + C.NS < kTraceBlocks) S.stats[(long long)kStatsStride * g + 16 + (blk + C.NS) * 8 + 2] = clock64();
                    split_issue<MODE>(C, a, b, blk + C.NS, st);
                    split_prefetch<MODE>(C, a, b, blk + C.NS + S.l2Ahead);
                }
                if (++st == C.NS) st = 0;
            }
            return;
        }
        // loader: one thread refills a stage as soon as the consumer has released the block it held
        const bool probe = STATS != 0 && S.stats != nullptr; // armed counters: lane 1 measures how long a refill takes to land
        volatile long long* issueClk = reinterpret_cast<volatile long long*>(smem + 256); // [NS], debug part of the header
        if (lane == 0)
        {
            int st = 0;
            for (int blk = C.NS; blk < C.nBlocks; blk++)
            {
                const unsigned need = (unsigned)(blk - C.NS + 1);
                while (ld_flag_smem(C.done) < need) __nanosleep(20);
                if (STATS == 1 && blk < kTraceBlocks) S.stats[(long long)kStatsStride * g + 16 + blk * 8 + 2] = clock64();
                if (probe) issueClk[st] = clock64();
                split_issue<MODE>(C, a, b, blk, st);
                split_prefetch<MODE>(C, a, b, blk + S.l2Ahead);
                if (++st == C.NS) st = 0;
            }
        }
        else if (lane == 1 && probe)
        {
            long long sum = 0, mx = 0;
            int n = 0, st = 0;
            unsigned par = 0u;
            for (int blk = 0; blk < C.nBlocks; blk++)
            {
                mbar_wait(&C.rawBar[st], par);
                if (blk >= C.NS)
                {
                    const long long d = clock64() - issueClk[st];
                    sum += d;
                    mx = d > mx ? d : mx;
                    n++;
                }
                if (++st == C.NS)
                {
                    st = 0;
                    par ^= 1u;
                }
            }
            long long* sp = S.stats + (long long)kStatsStride * g;
            sp[12] = sum;
            sp[13] = n;
            sp[15] = mx;
        }
    }
#if B200_PROD_SETS > 0
#if B200_PROD_DIRECT && B200_PROD_PIPE
#define PRODUCER_FN split_producer_pipe
#else
#define PRODUCER_FN split_producer_sets
#endif
    else if (role >= 0)
    {
        const int set = role / kProdPerSet, h0 = (role % kProdPerSet) * kProdM;
        switch (C.Lg)
        {
            case 0: PRODUCER_FN<MODE, 0, kProdM, STATS>(S, C, g, set, kProdSets, h0, lane, out, err, a, b); break;
            case 1: PRODUCER_FN<MODE, 1, kProdM, STATS>(S, C, g, set, kProdSets, h0, lane, out, err, a, b); break;
            case 2: PRODUCER_FN<MODE, 2, kProdM, STATS>(S, C, g, set, kProdSets, h0, lane, out, err, a, b); break;
            case 3: PRODUCER_FN<MODE, 3, kProdM, STATS>(S, C, g, set, kProdSets, h0, lane, out, err, a, b); break;
            case 4: PRODUCER_FN<MODE, 4, kProdM, STATS>(S, C, g, set, kProdSets, h0, lane, out, err, a, b); break;
            case 5: PRODUCER_FN<MODE, 5, kProdM, STATS>(S, C, g, set, kProdSets, h0, lane, out, err, a, b); break;
            default: PRODUCER_FN<MODE, 6, kProdM, STATS>(S, C, g, set, kProdSets, h0, lane, out, err, a, b); break;
        }
    }
#else
    else if (role >= 0)
    {
        const int h = role;
        switch (C.Lg)
        {
            case 0: split_producer<MODE, 0, STATS>(S, C, g, h, lane, out, err); break;
            case 1: split_producer<MODE, 1, STATS>(S, C, g, h, lane, out, err); break;
            case 2: split_producer<MODE, 2, STATS>(S, C, g, h, lane, out, err); break;
            case 3: split_producer<MODE, 3, STATS>(S, C, g, h, lane, out, err); break;
            case 4: split_producer<MODE, 4, STATS>(S, C, g, h, lane, out, err); break;
            case 5: split_producer<MODE, 5, STATS>(S, C, g, h, lane, out, err); break;
            default: split_producer<MODE, 6, STATS>(S, C, g, h, lane, out, err); break;
        }
    }
#endif
}

// Generic path (any W): plain loops, cross-warp values read when consumed.
template <int MODE>
__device__ __noinline__ void sweep_group_generic(const PipeDev& S, const int g, const int lane, const double* __restrict__ a,
                                                 const double* __restrict__ b, double* out, int* err, unsigned char* smem)
{
    constexpr unsigned FULL = 0xffffffffu;
    const SweepRing R = ring_setup<MODE>(S, g, lane, a, b, smem);
    const int W = R.W, CHg = R.CHg, recBytes = R.recBytes, NS = R.NS, dir = R.dir;
    const unsigned shflMask = R.shflMask;
    const int codeOff = W * 256 + lane * 4;
    const int coefOff = lane * 8;
    const long long outStride = dir > 0 ? 32 : -32;
    double* outPtr = out + ((long long)R.base + (dir > 0 ? 0 : R.nT - 1)) * 32 + lane;
    const int vecStart = dir > 0 ? lane : (CHg - 1) * 32 + lane;
    const int vecStride = dir > 0 ? 32 : -32;
    double hist[kSkew]; // hist[k]: the value this lane produced k+1 steps ago
#pragma unroll
    for (int k = 0; k < kSkew; k++) hist[k] = 0.0;
    int stCur = 0;
    unsigned par = 0u;
    for (int c = 0; c < R.nChunks; c++)
    {
        const unsigned char* sb = R.stages + (size_t)stCur * R.stageBytes;
        const unsigned char* recPtr = sb;
        const double* aP = reinterpret_cast<const double*>(sb + R.schedBytes) + vecStart;
        const double* bP = aP + CHg * 32;
        const int stNext = (stCur + 1 == NS) ? 0 : stCur + 1;
        const unsigned parNext = (stNext == 0) ? (par ^ 1u) : par;
        for (int q = 0; q < CHg; q++)
        {
            double acc = *aP;
            if (MODE == 0) acc *= *bP;
            for (int j = 0; j < W; j++)
            {
                const int code = *reinterpret_cast<const int*>(recPtr + codeOff + j * 128);
                const double cfj = *reinterpret_cast<const double*>(recPtr + coefOff + j * 256);
                double v = hist[0];
                if ((shflMask >> j) & 1u)
                {
                    const double sh = __shfl_sync(FULL, hist[kSkew - 1], code <= kCodeShfl ? kCodeShfl - code : lane);
                    if (code <= kCodeShfl) v = sh;
                }
                if (code >= 0)
                {
                    v = ld_relaxed(out + code);
                    if (is_sentinel(v)) v = sweep_spin(out + code, err, S.spinLimit);
                }
                if (code != kCodeNone) acc = sweep_apply<MODE>(acc, cfj, v);
            }
            st_relaxed(outPtr, acc);
#pragma unroll
            for (int k = kSkew - 1; k > 0; k--) hist[k] = hist[k - 1];
            hist[0] = acc;
            recPtr += recBytes;
            aP += vecStride;
            bP += vecStride;
            outPtr += outStride;
        }
        if (c + 1 < R.nChunks) mbar_wait(&R.bars[stNext], parNext);
        __syncwarp();
        if (lane == 0 && c + NS < R.nChunks) ring_issue<MODE>(R, a, b, c + NS, stCur);
        stCur = stNext;
        par = parNext;
    }
}

// One CTA (1 consumer + kNH producer warps + 1 loader warp) per group; CTAs take tickets so that groups start in a
// topological order of the group graph: a group only ever waits for groups that are already running.
constexpr int kSweepThreads = 32 * kSweepWarps;
template <int MODE, int STATS>
__global__ void __launch_bounds__(kSweepThreads, 2) k_sweep(PipeDev S, const double* __restrict__ a, const double* __restrict__ b,
                                                          double* out, unsigned* ticket, unsigned ticketBase, int* err,
                                                          const DevScalars* sc, int force)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned sTicket;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sTicket = atomicAdd(ticket, 1u) - ticketBase;
    __syncthreads();
    const unsigned t = sTicket;
    if (sc->done && !force) return;
    if ((int)t >= S.nGroups) return;
    const int g = S.order ? S.order[t] : (int)t;
    if ((S.debugFlags & 4) && t == 0) return; // fault injection (tests): this group's results never appear
    const bool stamp = STATS != 1 && S.stats != nullptr && threadIdx.x == 0; // start / end of the group
    if (stamp) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(S.stats[(long long)kStatsStride * g + 2]));
    if (S.gFast[g])
        sweep_group_split<MODE, STATS>(S, g, warp, lane, a, b, out, err, smem);
    else if (sweep_role(warp) == kRoleConsumer)
        sweep_group_generic<MODE>(S, g, lane, a, b, out, err, smem);
    if (stamp)
    { // thread 0 is the consumer's lane 0: the group's last result has been stored
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(S.stats[(long long)kStatsStride * g + 3]));
        S.stats[(long long)kStatsStride * g + 5] = S.gNT[g];
    }
}

// Fill the coefficient part of a sweep stream from the face coefficients (once per solve).
// prodMode 1: c = c1[f]*c2[f]  (calcReciprocalD);  prodMode 0: c = rD[slot]*c1[f]
__global__ void __launch_bounds__(256) k_pack_stream(PipeDev S, const double* __restrict__ c1, const double* __restrict__ c2,
                                                      const double* __restrict__ rD, int prodMode)
{
    const int g = blockIdx.x;
    if (g >= S.nGroups) return;
    const int nT = S.gNT[g], base = S.gBase[g];
    auto value = [&](int f, int step, int lane) -> double {
        if (f < 0) return 0.0;
        if (prodMode) return c1[f] * c2[f];
        const long long slot = ((long long)base + (S.dir > 0 ? step : nT - 1 - step)) * 32 + lane;
        return rD[slot] * c1[f];
    };
    const int tid = blockIdx.y * blockDim.x + threadIdx.x, nth = gridDim.y * blockDim.x;
    if (!S.gFast[g])
    {
        const int W = S.gW[g];
        const long long off = S.gTermOff[g];
        const int perStep = W * 32;
        const long long n = (long long)nT * perStep;
        unsigned char* gs = S.stream + 12ll * off;
        for (long long e = tid; e < n; e += nth)
        {
            const int step = (int)(e / perStep);
            const int rem = (int)(e - (long long)step * perStep);
            reinterpret_cast<double*>(gs + (size_t)step * perStep * 12)[rem] = value(S.face[off + e], step, rem & 31);
        }
        return;
    }
    const int Lg = S.gLg[g], Rg = S.gRg[g], Kg = S.gKg[g];
    const int pRec = Lg * 384 + Kg * 128, cRec = 256 + Rg * 256;
    if (Lg > 0)
    {
        const long long n = (long long)nT * Lg * 32;
        const int* face = S.pFace + S.gPFaceOff[g];
        unsigned char* ps = S.pStream + S.gPOff[g];
        for (long long e = tid; e < n; e += nth)
        {
            const int step = (int)(e / (Lg * 32));
            const int rem = (int)(e - (long long)step * (Lg * 32));
            reinterpret_cast<double*>(ps + (size_t)step * pRec)[rem] = value(face[e], step, rem & 31);
        }
    }
    if (Rg > 0)
    {
        const long long n = (long long)nT * Rg * 32;
        const int* face = S.cFace + S.gCFaceOff[g];
        unsigned char* cs = S.cStream + S.gCOff[g];
        for (long long e = tid; e < n; e += nth)
        {
            const int step = (int)(e / (Rg * 32));
            const int rem = (int)(e - (long long)step * (Rg * 32));
            const int r = rem >> 5; // plane-major inside a block of kNH steps (PipeSchedule::c_coef_off)
            reinterpret_cast<double*>(cs + (size_t)(step / kNH) * kNH * cRec + (size_t)kNH * 256 * (1 + r) + (size_t)(step % kNH) * 256)[rem & 31] =
                value(face[e], step, rem & 31);
        }
    }
}

// slot <-> cell permutation of vectors
__global__ void k_to_slots(size_t nSlots, const int* __restrict__ cellOfSlot, const double* __restrict__ in, double* __restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nSlots; i += (size_t)gridDim.x * blockDim.x)
    {
        const int c = cellOfSlot[i];
        out[i] = c >= 0 ? in[c] : 0.0;
    }
}
__global__ void k_from_slots(size_t nCells, const int* __restrict__ slotOfCell, const double* __restrict__ in, double* __restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nCells; i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[slotOfCell[i]];
}

// ------------------------------------------------------------------------------ vector kernels
#define B200_GRID_STRIDE(i, n) \
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)(n); i += (size_t)gridDim.x * blockDim.x)

__global__ void k_fill(size_t n, double* __restrict__ a, double v)
{
    B200_GRID_STRIDE(i, n) a[i] = v;
}
__global__ void k_fill2_sentinel(size_t n, double* __restrict__ a, double* __restrict__ b, const DevScalars* sc, int force)
{
    if (sc->done && !force) return;
    const double s = sentinel();
    const unsigned long long keep = l2_evict_last();
    B200_GRID_STRIDE(i, n)
    {
        st_keep(a + i, s, keep);
        if (b) st_keep(b + i, s, keep);
    }
}
__global__ void k_fill_xref(size_t n, double* __restrict__ a, const DevScalars* sc)
{
    const double v = sc->xRef;
    B200_GRID_STRIDE(i, n) a[i] = v;
}
__global__ void k_invert(size_t n, const double* __restrict__ in, double* __restrict__ out, const int* __restrict__ cellOfSlot)
{
    B200_GRID_STRIDE(i, n) out[i] = cellOfSlot[i] >= 0 ? 1.0 / in[i] : 0.0; // padding slots stay zero
}
__global__ void k_mul(size_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out,
                      const DevScalars* sc, int force)
{
    if (sc->done && !force) return;
    B200_GRID_STRIDE(i, n) out[i] = a[i] * b[i];
}
__global__ void k_copy(size_t n, const double* __restrict__ a, double* __restrict__ out, const DevScalars* sc, int force)
{
    if (sc->done && !force) return;
    B200_GRID_STRIDE(i, n) out[i] = a[i];
}

// sum x (for gAverage)
__global__ void __launch_bounds__(256) k_sum(size_t n, const double* __restrict__ x, double* partials, int pstride)
{
    double d[1] = {0.0};
    B200_GRID_STRIDE(i, n) d[0] += x[i];
    block_reduce_store<1>(d, partials, pstride, blockIdx.x);
}

// test hook / generic: sum a*b, sum |a|
__global__ void __launch_bounds__(256) k_dot_mag(size_t n, const double* __restrict__ a, const double* __restrict__ b,
                                                  double* partials, int pstride)
{
    double d[2] = {0.0, 0.0};
    B200_GRID_STRIDE(i, n)
    {
        d[0] += a[i] * b[i];
        d[1] += fabs(a[i]);
    }
    block_reduce_store<2>(d, partials, pstride, blockIdx.x);
}

// r = b - Ax; normFactor terms; bicg: rw = r, p = v = 0
// partials: [0] sum(|Ax - tmp| + |b - tmp|), [1] sum |r|, [2] sum r*r
__global__ void __launch_bounds__(256) k_init_residual(size_t n, const double* __restrict__ bsrc,
                                                        const double* __restrict__ Ax, const double* __restrict__ tmp,
                                                        double* __restrict__ r, double* __restrict__ rw,
                                                        double* __restrict__ zero1, double* __restrict__ zero2,
                                                        double* partials, int pstride)
{
    double d[3] = {0.0, 0.0, 0.0};
    B200_GRID_STRIDE(i, n)
    {
        const double bi = bsrc[i], ax = Ax[i], t = tmp[i];
        const double ri = bi - ax;
        r[i] = ri;
        if (rw) rw[i] = ri;
        if (zero1) zero1[i] = 0.0;
        if (zero2) zero2[i] = 0.0;
        d[0] += fabs(ax - t) + fabs(bi - t);
        d[1] += fabs(ri);
        d[2] += ri * ri;
    }
    block_reduce_store<3>(d, partials, pstride, blockIdx.x);
}

// p = r + beta*p - beta*omega*v ; sentinel-fill the two sweep outputs of the following precondition.
// partial quantity 1 = sum r*r; on restart (rho == 0) also rw = r.
__global__ void __launch_bounds__(256) k_bicg_p(size_t n, const double* __restrict__ r, double* __restrict__ p,
                                                 const double* __restrict__ v, double* __restrict__ rw,
                                                 double* __restrict__ fillA, double* __restrict__ fillB,
                                                 double* partials, int pstride, const DevScalars* sc)
{
    if (sc->done) return;
    const double beta = sc->beta, bo = sc->betaOmega;
    const int restart = sc->restart;
    const double sen = sentinel();
    const unsigned long long keep = l2_evict_last();
    double d[1] = {0.0};
    B200_GRID_STRIDE(i, n)
    {
        const double ri = r[i];
        p[i] = ri + beta * p[i] - bo * v[i];
        if (fillA) st_keep(fillA + i, sen, keep);
        if (fillB) st_keep(fillB + i, sen, keep);
        d[0] += ri * ri;
        if (restart) rw[i] = ri;
    }
    // (r, r) is always reduced (r is read anyway); OP_BICG_ALPHA uses it as rho only on restart.
    block_reduce_store<1>(d, partials + pstride, pstride, blockIdx.x); // quantity 1 -> red[1]
}

// s = r - alpha*v ; sentinel fills
__global__ void __launch_bounds__(256) k_bicg_s(size_t n, const double* __restrict__ r, const double* __restrict__ v,
                                                 double* __restrict__ s, double* __restrict__ fillA,
                                                 double* __restrict__ fillB, const DevScalars* sc)
{
    if (sc->done) return;
    const double alpha = sc->alpha;
    const double sen = sentinel();
    const unsigned long long keep = l2_evict_last();
    B200_GRID_STRIDE(i, n)
    {
        s[i] = r[i] - alpha * v[i];
        if (fillA) st_keep(fillA + i, sen, keep);
        if (fillB) st_keep(fillB + i, sen, keep);
    }
}

// x = x + alpha*ph + omega*sh ; r = s - omega*t ; partials: [0] sum |r|, [1] sum rw*r
__global__ void __launch_bounds__(256) k_bicg_xr(size_t n, double* __restrict__ x, const double* __restrict__ ph,
                                                  const double* __restrict__ sh, const double* __restrict__ s,
                                                  const double* __restrict__ t, double* __restrict__ r,
                                                  const double* __restrict__ rw, double* partials, int pstride,
                                                  const DevScalars* sc)
{
    if (sc->done) return;
    const double alpha = sc->alpha, omega = sc->omega;
    double d[2] = {0.0, 0.0};
    B200_GRID_STRIDE(i, n)
    {
        x[i] = x[i] + alpha * ph[i] + omega * sh[i];
        const double ri = s[i] - omega * t[i];
        r[i] = ri;
        d[0] += fabs(ri);
        d[1] += rw[i] * ri;
    }
    block_reduce_store<2>(d, partials, pstride, blockIdx.x);
}

// DIC / DILU smoother: psi += M^-1 (source - A psi)
__global__ void __launch_bounds__(256) k_add_inplace(size_t n, double* __restrict__ x, const double* __restrict__ w)
{
    B200_GRID_STRIDE(i, n) x[i] = x[i] + w[i];
}

// (a,b) with sentinel fill of two arrays
__global__ void __launch_bounds__(256) k_dot(size_t n, const double* __restrict__ a, const double* __restrict__ b,
                                              double* partials, int pstride, const DevScalars* sc)
{
    if (sc->done) return;
    double d[1] = {0.0};
    B200_GRID_STRIDE(i, n) d[0] += a[i] * b[i];
    block_reduce_store<1>(d, partials, pstride, blockIdx.x);
}

// PCG: pA = first ? wA : wA + beta*pA
__global__ void __launch_bounds__(256) k_pcg_p(size_t n, const double* __restrict__ wA, double* __restrict__ pA,
                                                const DevScalars* sc)
{
    if (sc->done) return;
    const int first = sc->first;
    const double beta = sc->beta;
    B200_GRID_STRIDE(i, n) pA[i] = first ? wA[i] : wA[i] + beta * pA[i];
}

// PCG: x += alpha*pA ; rA -= alpha*wA ; partial [0] sum |rA| ; sentinel fills for the next precondition
// PBiCG (foam/matrices/lduMatrix/solvers/PBiCG/PBiCG.C): the shadow system runs next to the primal one
__global__ void __launch_bounds__(256) k_sub(size_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out)
{
    B200_GRID_STRIDE(i, n) out[i] = a[i] - b[i];
}
__global__ void __launch_bounds__(256) k_pbicg_p(size_t n, const double* __restrict__ wA, double* __restrict__ pA,
                                                  const double* __restrict__ wT, double* __restrict__ pT, const DevScalars* sc)
{
    if (sc->done) return;
    const int first = sc->first;
    const double beta = sc->beta;
    B200_GRID_STRIDE(i, n)
    {
        pA[i] = first ? wA[i] : wA[i] + beta * pA[i];
        pT[i] = first ? wT[i] : wT[i] + beta * pT[i];
    }
}
__global__ void __launch_bounds__(256) k_pbicg_xr(size_t n, double* __restrict__ x, const double* __restrict__ pA,
                                                   double* __restrict__ rA, const double* __restrict__ wA,
                                                   double* __restrict__ rT, const double* __restrict__ wT, double* partials,
                                                   int pstride, const DevScalars* sc)
{
    if (sc->done) return;
    const double alpha = sc->alpha;
    double d[1] = {0.0};
    B200_GRID_STRIDE(i, n)
    {
        x[i] += alpha * pA[i];
        const double ri = rA[i] - alpha * wA[i];
        rA[i] = ri;
        rT[i] -= alpha * wT[i];
        d[0] += fabs(ri);
    }
    block_reduce_store<1>(d, partials, pstride, blockIdx.x);
}
__global__ void __launch_bounds__(256) k_pcg_xr(size_t n, double* __restrict__ x, const double* __restrict__ pA,
                                                 double* __restrict__ rA, const double* __restrict__ wA,
                                                 double* __restrict__ fillA, double* __restrict__ fillB,
                                                 double* partials, int pstride, const DevScalars* sc)
{
    if (sc->done) return;
    const double alpha = sc->alpha;
    const double sen = sentinel();
    const unsigned long long keep = l2_evict_last();
    double d[1] = {0.0};
    B200_GRID_STRIDE(i, n)
    {
        x[i] += alpha * pA[i];
        const double ri = rA[i] - alpha * wA[i];
        rA[i] = ri;
        d[0] += fabs(ri);
        if (fillA) st_keep(fillA + i, sen, keep);
        if (fillB) st_keep(fillB + i, sen, keep);
    }
    block_reduce_store<1>(d, partials, pstride, blockIdx.x);
}

// ------------------------------------------------------------------------------ GGI face transfer
__global__ void k_ggi_interpolate(int nTo, const int* __restrict__ offsets, const int* __restrict__ addr,
                                  const double* __restrict__ weights, const double* __restrict__ ff, int nComp,
                                  double* __restrict__ result)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTo * nComp) return;
    const int i = t / nComp, d = t - i * nComp;
    double acc = 0.0;
    for (int k = offsets[i]; k < offsets[i + 1]; k++) acc += ff[(size_t)addr[k] * nComp + d] * weights[k];
    result[t] = acc;
}

__global__ void k_scatter_zone(int nLocal, const int* __restrict__ addr, const double* __restrict__ pField, int nComp,
                               double* __restrict__ gField)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nLocal * nComp) return;
    const int i = t / nComp, d = t - i * nComp;
    gField[(size_t)addr[i] * nComp + d] = pField[t];
}

__global__ void k_gather_zone(int nLocal, const int* __restrict__ addr, const double* __restrict__ gField, int nComp,
                              double* __restrict__ pField)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nLocal * nComp) return;
    const int i = t / nComp, d = t - i * nComp;
    pField[t] = gField[(size_t)addr[i] * nComp + d];
}

} // namespace b200
