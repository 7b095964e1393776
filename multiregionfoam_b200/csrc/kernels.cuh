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
        if (sc->done) hostFlags[0] = 1;
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
                if (hostFlags) hostFlags[0] = 1;
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
__global__ void __launch_bounds__(1024) k_finalize(const double* __restrict__ partials, int pstride, PartCounts cnt,
                                                    int nd, DevScalars* sc, int op, int applyOp, int force,
                                                    double* history, volatile int* hostFlags)
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
template <int ND>
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
            double acc = diag[row] * x[row];
            int j = 0;
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
template <int ND>
__global__ void __launch_bounds__(128) k_iface(int nTouched, const int* __restrict__ rows,
                                                const int* __restrict__ rowStart, const int* __restrict__ entCoef,
                                                const int* __restrict__ entSrc, const int* __restrict__ entCnt,
                                                const int* __restrict__ gSrc, const double* __restrict__ gW,
                                                const double* __restrict__ coef, const double* __restrict__ x,
                                                const double* __restrict__ recv, double* __restrict__ y,
                                                const double* __restrict__ d0, double* partials, int pstride,
                                                int slotBase, const DevScalars* sc, int force)
{
    if (sc->done && !force) return;
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
            if (cnt == 0)
            {
                const int s = entSrc[e];
                pnf = s >= 0 ? x[s] : recv[-1 - s];
            }
            else
            {
                pnf = 0.0; // GGIInterpolation::interpolate: zero-initialised, accumulated in list order
                const int g0 = entSrc[e];
                for (int k = 0; k < cnt; k++)
                {
                    const int s = gSrc[g0 + k];
                    const double f = s >= 0 ? x[s] : recv[-1 - s];
                    pnf += f * gW[g0 + k];
                }
            }
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
// doubles as its own ready flag (sentinel until written), and are prefetched kPF steps ahead.
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
    int cBlockBytes;  // bytes of one block of the consumer ring (max over the fast groups)
    int debugFlags;   // bit 0: consumer always takes the select path (debug)
    long long* stats; // optional [8 * nGroups]: consumer cycles, wait cycles, start ns, end ns, producer polls, nT (debug)
};

constexpr int kPF = 4; // prefetch distance (time steps) of cross-warp values

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

__device__ __noinline__ double sweep_spin(const double* p, int* err)
{
    for (long long tries = 0; tries < (1ll << 24); tries++)
    {
        const double v = ld_relaxed(p);
        if (!is_sentinel(v)) return v;
        if (tries >= 16) __nanosleep(tries > 4096 ? 400 : 64); // later: a busy-spinning warp steals issue slots from the consumer warps
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
// A lone warp issues one instruction every ~5 cycles and the recurrence is serial, so the consumer
// warp must execute as little as possible per time step.  Everything static (which terms lead, which
// lane a term shuffles from, how many terms a step has) was decided on the host:
//   * kNH = 8 producer warps (producer h prepares step 8b+h of block b): read the P-stream record and
//     the input vectors from the TMA-fed raw ring, fetch the cross-group values (prefetched one block
//     ahead, verified against the sentinel), apply the row's LEADING terms in reference order and
//     hand {acc0, cval[]} to the consumer through the hdr part of the consumer ring;
//   * the C-stream records {meta, remaining coefficients} go by TMA straight into the consumer ring;
//   * the consumer warp runs only the recurrence: per remaining term  v = shfl(prev, lane from meta);
//     acc -= coef * v;  then one store.  Steps whose remaining terms need a cross-group value (block
//     seams) and calcReciprocalD (division) take a select path.
constexpr int kNH = kSweepBlock; // producer warps = steps per block (8)
constexpr int kRB = 3;           // blocks in the consumer ring

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct SplitCtx
{
    int nT, base, dir, nBlocks, NS, stageBytes;
    int Lg, Rg, Kg, pRec, cRec, hdrStep, cBlockBytes;
    unsigned rawChunkBytes;
    unsigned long long *rawBar, *hdrFull, *empty, *cTma;
    unsigned char *rawRing, *cRing;
    const unsigned char *pStream, *cStream;
};

template <int MODE>
__device__ __forceinline__ void split_issue_raw(const SplitCtx& C, const double* a, const double* b, int blk, int st)
{
    unsigned char* dst = C.rawRing + (size_t)st * C.stageBytes;
    mbar_expect_tx(&C.rawBar[st], C.rawChunkBytes);
    const unsigned pBytes = (unsigned)(kNH * C.pRec);
    bulk_g2s(dst, C.pStream + (size_t)blk * pBytes, pBytes, &C.rawBar[st]);
    const long long s0 = C.dir > 0 ? ((long long)C.base + (long long)blk * kNH) * 32 : ((long long)C.base + C.nT - (long long)(blk + 1) * kNH) * 32;
    bulk_g2s(dst + pBytes, a + s0, kNH * 256, &C.rawBar[st]);
    if (MODE == 0) bulk_g2s(dst + pBytes + kNH * 256, b + s0, kNH * 256, &C.rawBar[st]);
}

__device__ __forceinline__ void split_issue_c(const SplitCtx& C, int blk, int slot)
{
    const unsigned bytes = (unsigned)(kNH * C.cRec);
    mbar_expect_tx(&C.cTma[slot], bytes);
    bulk_g2s(C.cRing + (size_t)slot * C.cBlockBytes, C.cStream + (size_t)blk * bytes, bytes, &C.cTma[slot]);
}

// ------------------------------------------------------------------------------------------ consumer
template <int MODE, int RG>
__device__ __forceinline__ void split_consumer(const PipeDev& S, const SplitCtx& C, const int g, const int lane, const double* a,
                                               const double* b, double* out)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int dir = C.dir, nT = C.nT;
    double* outPtr = out + ((long long)C.base + (dir > 0 ? 0 : nT - 1)) * 32 + lane;
    const long long outStride = dir > 0 ? 32 : -32;
    const int hdrOff = kNH * C.cRec; // hdr part of a ring block follows its C-records
    double prev = 0.0;
    int slot = 0;
    unsigned par = 0u;
    long long tWait = 0, t0 = 0, g0 = 0, tTma = 0, tTail = 0;
    if (S.stats)
    {
        t0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    }
    for (int blk = 0; blk < C.nBlocks; blk++)
    {
        const unsigned char* blkBase = C.cRing + (size_t)slot * C.cBlockBytes;
        if (S.stats)
        {
            const long long w0 = clock64();
            mbar_wait(&C.cTma[slot], par);
            tTma += clock64() - w0;
        }
        else
            mbar_wait(&C.cTma[slot], par);
#pragma unroll
        for (int hb = 0; hb < kNH; hb += 4)
        {
            if (S.stats)
            {
                const long long w0 = clock64();
                mbar_wait(&C.hdrFull[slot * 2 + (hb >> 2)], par);
                tWait += clock64() - w0;
            }
            else
                mbar_wait(&C.hdrFull[slot * 2 + (hb >> 2)], par);
            double acc0[4], cf[4][RG];
            unsigned long long meta[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const unsigned char* rec = blkBase + (size_t)(hb + q) * C.cRec;
                acc0[q] = *reinterpret_cast<const double*>(blkBase + hdrOff + (size_t)(hb + q) * C.hdrStep + lane * 8);
                meta[q] = *reinterpret_cast<const unsigned long long*>(rec + lane * 8);
#pragma unroll
                for (int r = 0; r < RG; r++) cf[q][r] = *reinterpret_cast<const double*>(rec + 256 + r * 256 + lane * 8);
            }
            unsigned flags = 0, rmax = 0;
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                flags |= (unsigned)(meta[q] >> 56) & 1u;
                rmax = max(rmax, (unsigned)(meta[q] >> 48) & 0xffu);
            }
            if (MODE != 2 && !flags && !(S.debugFlags & 1))
            {
                // R = number of terms to run for these 4 steps (uniform); terms beyond a lane's own are padding
                auto run = [&](auto RR) {
                    constexpr int R = decltype(RR)::value;
#pragma unroll
                    for (int q = 0; q < 4; q++)
                    {
                        double sh[R > 0 ? R : 1];
#pragma unroll
                        for (int r = 0; r < R; r++) sh[r] = __shfl_sync(FULL, prev, (int)(meta[q] >> (8 * r)) & 31);
                        double acc = acc0[q];
#pragma unroll
                        for (int r = 0; r < R; r++) acc -= cf[q][r] * sh[r];
                        st_relaxed(outPtr, acc);
                        prev = acc;
                        outPtr += outStride;
                    }
                };
                if (RG >= 2 && rmax == RG - 1)
                    run(std::integral_constant<int, (RG >= 2 ? RG - 1 : RG)>());
                else if (RG >= 3 && rmax == RG - 2)
                    run(std::integral_constant<int, (RG >= 3 ? RG - 2 : RG)>());
                else
                    run(std::integral_constant<int, RG>());
            }
            else
            {
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const unsigned char* hd = blkBase + hdrOff + (size_t)(hb + q) * C.hdrStep + lane * 8;
                    const double cv0 = *reinterpret_cast<const double*>(hd + 256);
                    const double cv1 = C.Kg > 1 ? *reinterpret_cast<const double*>(hd + 512) : 0.0;
                    double acc = acc0[q];
#pragma unroll
                    for (int r = 0; r < RG; r++)
                    {
                        const unsigned byte = (unsigned)(meta[q] >> (8 * r)) & 0xffu;
                        const double sh = __shfl_sync(FULL, prev, (int)(byte & 31u));
                        double v = (byte & 0x80u) ? ((byte & 0x20u) ? cv1 : cv0) : sh;
                        if (byte & 0x40u) v = MODE == 2 ? 1.0 : 0.0; // padding term (coefficient 0)
                        acc = sweep_apply<MODE>(acc, cf[q][r], v);
                    }
                    st_relaxed(outPtr, acc);
                    prev = acc;
                    outPtr += outStride;
                }
            }
        }
        const long long w1 = S.stats ? clock64() : 0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&C.empty[slot]); // block done: producers may reuse its hdr, the loader warp refills
        if (++slot == kRB)
        {
            slot = 0;
            par ^= 1u;
        }
        if (S.stats) tTail += clock64() - w1;
    }
    if (S.stats && lane == 0)
    {
        long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        long long* st = S.stats + 8ll * g;
        st[0] = clock64() - t0;
        st[1] = tWait;
        st[2] = g0;
        st[3] = g1;
        st[5] = nT;
        st[6] = tTma;
        st[7] = tTail;
    }
}

// ------------------------------------------------------------------------------------------ producers
template <int MODE, int LG>
__device__ __forceinline__ void split_producer(const PipeDev& S, const SplitCtx& C, const int g, const int h, const int lane, double* out,
                                               int* err)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int LGA = LG > 0 ? LG : 1;
    const int dir = C.dir;
    const int codeOff = LG * 256 + lane * 4, constOff = LG * 384 + lane * 4;
    const int vecIdx = (dir > 0 ? h : kNH - 1 - h) * 32 + lane; // this producer's step inside a raw chunk
    const double neutral = MODE == 2 ? 1.0 : 0.0;
    const int hdrOff = kNH * C.cRec;
    int codes[LGA], codesN[LGA], cc[2], ccN[2];
    double mv[LGA], mvN[LGA], mc[2], mcN[2];
    auto fetch = [&](const unsigned char* rec, int* cd, double* vals, int* kc, double* kv) {
#pragma unroll
        for (int i = 0; i < LG; i++)
        {
            cd[i] = *reinterpret_cast<const int*>(rec + codeOff + i * 128);
            vals[i] = neutral;
            if (cd[i] >= 0) vals[i] = ld_relaxed(out + cd[i]);
        }
#pragma unroll
        for (int k = 0; k < 2; k++)
        {
            kc[k] = (k < C.Kg) ? *reinterpret_cast<const int*>(rec + constOff + k * 128) : -1;
            kv[k] = 0.0;
            if (kc[k] >= 0) kv[k] = ld_relaxed(out + kc[k]);
        }
    };
    int st = 0, slot = 0;
    unsigned parRaw = 0u, parC = 0u;
    mbar_wait(&C.rawBar[0], 0u);
    fetch(C.rawRing + (size_t)h * C.pRec, codes, mv, cc, mc);
    for (int blk = 0; blk < C.nBlocks; blk++)
    {
        const unsigned char* sb = C.rawRing + (size_t)st * C.stageBytes;
        const unsigned char* rec = sb + (size_t)h * C.pRec;
        const double* aP = reinterpret_cast<const double*>(sb + kNH * C.pRec) + vecIdx;
        // ---- prefetch the codes / cross-group values of this producer's step in the next block
        int stN = st + 1;
        unsigned parN = parRaw;
        if (stN == C.NS)
        {
            stN = 0;
            parN ^= 1u;
        }
        if (blk + 1 < C.nBlocks)
        {
            mbar_wait(&C.rawBar[stN], parN);
            fetch(C.rawRing + (size_t)stN * C.stageBytes + (size_t)h * C.pRec, codesN, mvN, ccN, mcN);
        }
        // ---- operands of the current step first, the (possibly late) cross-group values last
        double acc = *aP;
        if (MODE == 0) acc *= aP[kNH * 32];
        double cf[LGA];
#pragma unroll
        for (int i = 0; i < LG; i++) cf[i] = *reinterpret_cast<const double*>(rec + lane * 8 + i * 256);
        bool bad = false;
#pragma unroll
        for (int i = 0; i < LG; i++) bad |= codes[i] >= 0 && is_sentinel(mv[i]);
#pragma unroll
        for (int k = 0; k < 2; k++) bad |= cc[k] >= 0 && is_sentinel(mc[k]);
        if (__any_sync(FULL, bad))
        { // a value had not arrived when it was prefetched: poll for it
            if (S.stats && lane == 0) atomicAdd((unsigned long long*)(S.stats + 8ll * g + 4), 1ull);
#pragma unroll
            for (int i = 0; i < LG; i++)
                if (codes[i] >= 0 && is_sentinel(mv[i])) mv[i] = sweep_spin(out + codes[i], err);
#pragma unroll
            for (int k = 0; k < 2; k++)
                if (cc[k] >= 0 && is_sentinel(mc[k])) mc[k] = sweep_spin(out + cc[k], err);
        }
#pragma unroll
        for (int i = 0; i < LG; i++) acc = sweep_apply<MODE>(acc, cf[i], mv[i]); // padding: coefficient 0, neutral value
        // ---- hand over once the consumer has released the ring block
        if (blk >= kRB) mbar_wait(&C.empty[slot], parC ^ 1u);
        unsigned char* hd = C.cRing + (size_t)slot * C.cBlockBytes + hdrOff + (size_t)h * C.hdrStep + lane * 8;
        *reinterpret_cast<double*>(hd) = acc;
        *reinterpret_cast<double*>(hd + 256) = mc[0];
        if (C.Kg > 1) *reinterpret_cast<double*>(hd + 512) = mc[1];
        __syncwarp();
        if (lane == 0) mbar_arrive(&C.hdrFull[slot * 2 + (h >> 2)]);
        if (++slot == kRB)
        {
            slot = 0;
            parC ^= 1u;
        }
        st = stN;
        parRaw = parN;
#pragma unroll
        for (int i = 0; i < LG; i++)
        {
            codes[i] = codesN[i];
            mv[i] = mvN[i];
        }
#pragma unroll
        for (int k = 0; k < 2; k++)
        {
            cc[k] = ccN[k];
            mc[k] = mcN[k];
        }
    }
}

template <int MODE>
__device__ __forceinline__ void sweep_group_split(const PipeDev& S, const int g, const int warp, const int lane, const double* __restrict__ a,
                                                  const double* __restrict__ b, double* out, int* err, unsigned char* smem)
{
    SplitCtx C;
    C.nT = S.gNT[g];
    C.base = S.gBase[g];
    C.dir = S.dir;
    C.nBlocks = C.nT / kNH;
    C.NS = S.nStages;
    C.stageBytes = S.stageBytes;
    C.Lg = S.gLg[g];
    C.Rg = S.gRg[g];
    C.Kg = S.gKg[g];
    C.pRec = C.Lg * 384 + C.Kg * 128;
    C.cRec = 256 + C.Rg * 256;
    C.hdrStep = 256 * (1 + C.Kg);
    C.cBlockBytes = S.cBlockBytes;
    C.rawChunkBytes = (unsigned)(kNH * C.pRec + kNH * 256 * (MODE == 0 ? 2 : 1));
    C.rawBar = reinterpret_cast<unsigned long long*>(smem);
    C.hdrFull = reinterpret_cast<unsigned long long*>(smem + 128);
    C.empty = C.hdrFull + 2 * kRB;
    C.cTma = C.empty + kRB;
    C.rawRing = smem + 256;
    C.cRing = C.rawRing + (size_t)C.NS * C.stageBytes;
    C.pStream = S.pStream + S.gPOff[g];
    C.cStream = S.cStream + S.gCOff[g];
    if (warp == 0 && lane == 0)
    {
        for (int st = 0; st < C.NS; st++) mbar_init(&C.rawBar[st], 1);
        for (int i = 0; i < 2 * kRB; i++) mbar_init(&C.hdrFull[i], kNH / 2);
        for (int i = 0; i < kRB; i++)
        {
            mbar_init(&C.empty[i], 1);
            mbar_init(&C.cTma[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int c = 0; c < C.NS && c < C.nBlocks; c++) split_issue_raw<MODE>(C, a, b, c, c);
        for (int c = 0; c < kRB && c < C.nBlocks; c++) split_issue_c(C, c, c);
    }
    __syncthreads();
    if (warp == 0)
    {
        switch (C.Rg)
        {
            case 1: split_consumer<MODE, 1>(S, C, g, lane, a, b, out); break;
            case 2: split_consumer<MODE, 2>(S, C, g, lane, a, b, out); break;
            case 3: split_consumer<MODE, 3>(S, C, g, lane, a, b, out); break;
            case 4: split_consumer<MODE, 4>(S, C, g, lane, a, b, out); break;
            case 5: split_consumer<MODE, 5>(S, C, g, lane, a, b, out); break;
            default: split_consumer<MODE, 6>(S, C, g, lane, a, b, out); break;
        }
    }
    else if (warp == 1 + kNH)
    {
        // loader: one thread refills the rings as soon as the consumer has released a block
        if (lane == 0)
        {
            int slot = 0, stRaw = 0;
            unsigned par = 0u;
            for (int blk = 0; blk < C.nBlocks; blk++)
            {
                if (blk + kRB >= C.nBlocks && blk + C.NS >= C.nBlocks) break;
                mbar_wait(&C.empty[slot], par);
                if (blk + kRB < C.nBlocks) split_issue_c(C, blk + kRB, slot);
                if (blk + C.NS < C.nBlocks) split_issue_raw<MODE>(C, a, b, blk + C.NS, stRaw);
                if (++slot == kRB)
                {
                    slot = 0;
                    par ^= 1u;
                }
                if (++stRaw == C.NS) stRaw = 0;
            }
        }
    }
    else
    {
        const int h = warp - 1;
        switch (C.Lg)
        {
            case 0: split_producer<MODE, 0>(S, C, g, h, lane, out, err); break;
            case 1: split_producer<MODE, 1>(S, C, g, h, lane, out, err); break;
            case 2: split_producer<MODE, 2>(S, C, g, h, lane, out, err); break;
            case 3: split_producer<MODE, 3>(S, C, g, h, lane, out, err); break;
            case 4: split_producer<MODE, 4>(S, C, g, h, lane, out, err); break;
            case 5: split_producer<MODE, 5>(S, C, g, h, lane, out, err); break;
            default: split_producer<MODE, 6>(S, C, g, h, lane, out, err); break;
        }
    }
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
    double prev = 0.0;
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
                double v = prev;
                if ((shflMask >> j) & 1u) v = __shfl_sync(FULL, prev, code <= kCodeShfl ? kCodeShfl - code : lane);
                if (code >= 0)
                {
                    v = ld_relaxed(out + code);
                    if (is_sentinel(v)) v = sweep_spin(out + code, err);
                }
                if (code != kCodeNone) acc = sweep_apply<MODE>(acc, cfj, v);
            }
            st_relaxed(outPtr, acc);
            prev = acc;
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
constexpr int kSweepThreads = 32 * (2 + kNH); // consumer + producers + loader
template <int MODE>
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
    if (S.gFast[g])
        sweep_group_split<MODE>(S, g, warp, lane, a, b, out, err, smem);
    else if (warp == 0)
        sweep_group_generic<MODE>(S, g, lane, a, b, out, err, smem);
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
            reinterpret_cast<double*>(cs + (size_t)step * cRec + 256)[rem] = value(face[e], step, rem & 31);
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
    B200_GRID_STRIDE(i, n)
    {
        a[i] = s;
        if (b) b[i] = s;
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
    double d[1] = {0.0};
    B200_GRID_STRIDE(i, n)
    {
        const double ri = r[i];
        p[i] = ri + beta * p[i] - bo * v[i];
        if (fillA) fillA[i] = sen;
        if (fillB) fillB[i] = sen;
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
    B200_GRID_STRIDE(i, n)
    {
        s[i] = r[i] - alpha * v[i];
        if (fillA) fillA[i] = sen;
        if (fillB) fillB[i] = sen;
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
__global__ void __launch_bounds__(256) k_pcg_xr(size_t n, double* __restrict__ x, const double* __restrict__ pA,
                                                 double* __restrict__ rA, const double* __restrict__ wA,
                                                 double* __restrict__ fillA, double* __restrict__ fillB,
                                                 double* partials, int pstride, const DevScalars* sc)
{
    if (sc->done) return;
    const double alpha = sc->alpha;
    const double sen = sentinel();
    double d[1] = {0.0};
    B200_GRID_STRIDE(i, n)
    {
        x[i] += alpha * pA[i];
        const double ri = rA[i] - alpha * wA[i];
        rA[i] = ri;
        d[0] += fabs(ri);
        if (fillA) fillA[i] = sen;
        if (fillB) fillB[i] = sen;
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
