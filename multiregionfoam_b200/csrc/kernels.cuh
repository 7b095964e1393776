// kernels.cuh -- hand-written sm_100a kernels of the coupled LDU Krylov solve.
// All FP64, all HBM-bound; compiled with -fmad=false so that every product/sum rounds exactly
// like the reference build (g++ -O3, x86-64 baseline, no FMA contraction).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

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
        const int row = s * 32 + lane;
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
struct SweepDev
{
    int nWarps;
    int dir;
    const int* nLanes;
    const int* nSteps;
    const int* W;
    const int* laneBase;
    const long long* chainBase;
    const long long* offBase;
    const int* laneStart;
    const int* laneLen;
    const int* chainFace;
    const int* offFace;
    const int* offCol;
    double* chainC; // packed coefficients (mode specific)
    double* offC;
};

// MODE 0: forward  : acc = a[row]*b[row];  acc -= c * w[nbr]      (a = rD, b = rA,  c = rD[row]*lower[f])
// MODE 1: backward : acc = a[row];         acc -= c * w[nbr]      (a = forward result, c = rD[row]*upper[f])
// MODE 2: calcReciprocalD: acc = a[row];   acc -= c / w[nbr]      (a = diag,        c = upper[f]*lower[f])
template <int MODE>
__device__ __forceinline__ double sweep_apply(double acc, double c, double v)
{
    return MODE == 2 ? acc - c / v : acc - c * v;
}

__device__ __noinline__ double sweep_spin(const double* p, int* err)
{
    unsigned ns = 20;
    for (long long tries = 0; tries < (1ll << 22); tries++)
    {
        const double v = ld_volatile(p);
        if (!is_sentinel(v)) return v;
        __nanosleep(ns);
        if (ns < 200) ns += ns;
        if ((tries & 1023) == 1023 && *(volatile int*)err) break;
    }
    atomicExch(err, 1);
    return ld_volatile(p);
}

template <int MODE, int WMAX, int D>
__device__ __forceinline__ void sweep_warp_fast(const SweepDev& S, int w, int lane, const double* __restrict__ a,
                                                const double* __restrict__ b, double* out, int* err)
{
    const int nl = S.nLanes[w], nSteps = S.nSteps[w], W = S.W[w];
    const bool act = lane < nl;
    const int start = act ? S.laneStart[S.laneBase[w] + lane] : 0;
    const int len = act ? S.laneLen[S.laneBase[w] + lane] : 0;
    const long long cb = S.chainBase[w] + lane, ob = S.offBase[w] + lane;
    const int dir = S.dir;
    const double* __restrict__ chainC = S.chainC;
    const double* __restrict__ offC = S.offC;
    const int* __restrict__ offCol = S.offCol;

    double rc[D], ra[D], rb[D], roc[D][WMAX], rv[D][WMAX];
    int rcol[D][WMAX];

#define B200_LOAD_STEP(d, sArg)                                                        \
    {                                                                                  \
        const int s_ = (sArg);                                                         \
        if (s_ < len)                                                                  \
        {                                                                              \
            const int row_ = start + dir * s_;                                         \
            rc[d] = chainC[cb + (long long)s_ * nl];                                   \
            ra[d] = a[row_];                                                           \
            if (MODE == 0) rb[d] = b[row_];                                            \
            _Pragma("unroll") for (int j = 0; j < WMAX; j++)                           \
            {                                                                          \
                rcol[d][j] = -1;                                                       \
                if (j < W)                                                             \
                {                                                                      \
                    const long long idx_ = ob + ((long long)s_ * W + j) * nl;          \
                    rcol[d][j] = offCol[idx_];                                         \
                    roc[d][j] = offC[idx_];                                            \
                }                                                                      \
            }                                                                          \
            _Pragma("unroll") for (int j = 0; j < WMAX; j++)                           \
            {                                                                          \
                if (rcol[d][j] >= 0) rv[d][j] = ld_volatile(out + rcol[d][j]);         \
            }                                                                          \
        }                                                                              \
    }

#pragma unroll
    for (int d = 0; d < D; d++) B200_LOAD_STEP(d, d);

    double prev = 0.0;
    for (int s0 = 0; s0 < nSteps; s0 += D)
    {
#pragma unroll
        for (int d = 0; d < D; d++)
        {
            const int s = s0 + d;
            if (s < len)
            {
                const int row = start + dir * s;
                double acc = MODE == 0 ? ra[d] * rb[d] : ra[d];
#pragma unroll
                for (int j = 0; j < WMAX; j++)
                {
                    if (rcol[d][j] >= 0)
                    {
                        double v = rv[d][j];
                        if (is_sentinel(v)) v = sweep_spin(out + rcol[d][j], err);
                        acc = sweep_apply<MODE>(acc, roc[d][j], v);
                    }
                }
                if (s > 0) acc = sweep_apply<MODE>(acc, rc[d], prev);
                st_volatile(out + row, acc);
                prev = acc;
            }
            B200_LOAD_STEP(d, s + D);
        }
    }
#undef B200_LOAD_STEP
}

template <int MODE>
__device__ __noinline__ void sweep_warp_generic(const SweepDev& S, int w, int lane, const double* __restrict__ a,
                                                const double* __restrict__ b, double* out, int* err)
{
    const int nl = S.nLanes[w], W = S.W[w];
    if (lane >= nl) return;
    const int start = S.laneStart[S.laneBase[w] + lane];
    const int len = S.laneLen[S.laneBase[w] + lane];
    const long long cb = S.chainBase[w] + lane, ob = S.offBase[w] + lane;
    double prev = 0.0;
    for (int s = 0; s < len; s++)
    {
        const int row = start + S.dir * s;
        double acc = MODE == 0 ? a[row] * b[row] : a[row];
        for (int j = 0; j < W; j++)
        {
            const long long idx = ob + ((long long)s * W + j) * nl;
            const int col = S.offCol[idx];
            if (col < 0) break; // entries are packed first
            double v = ld_volatile(out + col);
            if (is_sentinel(v)) v = sweep_spin(out + col, err);
            acc = sweep_apply<MODE>(acc, S.offC[idx], v);
        }
        if (s > 0) acc = sweep_apply<MODE>(acc, S.chainC[cb + (long long)s * nl], prev);
        st_volatile(out + row, acc);
        prev = acc;
    }
}

// One warp per 32 independent chains; CTAs take tickets so that warps start in chain-level order.
template <int MODE>
__global__ void __launch_bounds__(128) k_sweep(SweepDev S, const double* __restrict__ a, const double* __restrict__ b,
                                                double* out, unsigned* ticket, unsigned ticketBase, int* err,
                                                const DevScalars* sc, int force)
{
    __shared__ unsigned sTicket;
    if (threadIdx.x == 0) sTicket = atomicAdd(ticket, 1u) - ticketBase;
    __syncthreads();
    if (sc->done && !force) return;
    const int w = (int)sTicket * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= S.nWarps) return;
    const int lane = threadIdx.x & 31;
    const int W = S.W[w];
    if (W <= 2)
        sweep_warp_fast<MODE, 2, 8>(S, w, lane, a, b, out, err);
    else if (W <= 4)
        sweep_warp_fast<MODE, 4, 4>(S, w, lane, a, b, out, err);
    else
        sweep_warp_generic<MODE>(S, w, lane, a, b, out, err);
}

// Fill the packed sweep coefficients from the face coefficients.
// prodMode 1: c = c1[f]*c2[f]  (calcReciprocalD);  prodMode 0: c = rD[row]*c1[f]
__global__ void __launch_bounds__(256) k_pack_sweep(SweepDev S, const double* __restrict__ c1,
                                                     const double* __restrict__ c2, const double* __restrict__ rD,
                                                     int prodMode)
{
    const int w = blockIdx.x;
    if (w >= S.nWarps) return;
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5, ny = blockDim.x >> 5;
    const int nl = S.nLanes[w], nSteps = S.nSteps[w], W = S.W[w];
    if (lane >= nl) return;
    const int start = S.laneStart[S.laneBase[w] + lane];
    const int len = S.laneLen[S.laneBase[w] + lane];
    const long long cb = S.chainBase[w] + lane, ob = S.offBase[w] + lane;
    for (int s = ty; s < nSteps; s += ny)
    {
        if (s >= len) continue;
        const double scale = prodMode ? 1.0 : rD[start + S.dir * s];
        {
            const long long idx = cb + (long long)s * nl;
            const int f = S.chainFace[idx];
            S.chainC[idx] = f >= 0 ? (prodMode ? c1[f] * c2[f] : scale * c1[f]) : 0.0;
        }
        for (int j = 0; j < W; j++)
        {
            const long long idx = ob + ((long long)s * W + j) * nl;
            const int f = S.offFace[idx];
            S.offC[idx] = f >= 0 ? (prodMode ? c1[f] * c2[f] : scale * c1[f]) : 0.0;
        }
    }
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
__global__ void k_invert(size_t n, const double* __restrict__ in, double* __restrict__ out)
{
    B200_GRID_STRIDE(i, n) out[i] = 1.0 / in[i];
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
