// gs_smoother.cuh -- Gauss-Seidel smoother / smoothSolver for one lduMatrix (include/b200_smooth.h), included by b200_ldu.cu.
//
// Reference: foam-extend 4.1 GaussSeidelSmoother.C (smooth), smoothSolver.C (solve), lduMatrixATmul.C (Amul, residual);
// restated on the CPU by the test oracle (orc_gs_smooth, orc_gs_solve).
//
// Layout: the matrix stays in LDU form (face-ordered upper / lower, cell-ordered diag) with the row pointers ownerStart /
// losortStart / losort.  A sweep is level-scheduled like the block-coupled sweeps (blk_system.cuh): rows in wavefront-level order
// of the lower-neighbour graph, one thread per row, chunks of 256 positions handed out in order by a ticket counter, a row polls
// the new values of its lower neighbours through the output vector (NaN sentinel).  HBM-bound: 8 (diag) + 8 (b) + 8 (old) +
// 8 (new) + per face 2 x (8 coefficient + 4 index) + the polled values: ~100 B per hex cell.
#pragma once

namespace b200
{
constexpr int kGsThreads = 256;
constexpr int kGsRedBlocks = 1184; // 8 x 148: fixed reduction grid

struct GsDev
{
    int n, nf;
    const int *l, *u, *losort, *losortStart, *ownerStart;
    const double *diag, *upper, *lower;
};

// psiNew[c] = ((bPrime[c] - sum_{lower faces, ascending} lower[f]*psiNew[l[f]]) - sum_{owner faces} upper[f]*psiOld[u[f]]) / diag[c]
__global__ void __launch_bounds__(kGsThreads)
    k_gs_sweep(GsDev M, const int* __restrict__ rows, int nPos, const double* __restrict__ bPrime, const double* __restrict__ psiOld, double* psiNew,
               unsigned* ticket, unsigned ticketBase, int* err)
{
    __shared__ unsigned sTicket;
    if (threadIdx.x == 0) sTicket = atomicAdd(ticket, 1u) - ticketBase;
    __syncthreads();
    const long long pos = (long long)sTicket * kGsThreads + threadIdx.x;
    if (pos >= nPos) return;
    const int row = rows[pos];
    if (row < 0) return;
    double acc = bPrime[row];
    const double rd = M.diag[row];
    const int o0 = M.ownerStart[row], o1 = M.ownerStart[row + 1];
    const int k0 = M.losortStart[row], k1 = M.losortStart[row + 1];
    // Everything that does not depend on other rows of this sweep is fetched BEFORE the first poll - the rows of a level wait
    // for the level before, so what a row does after its neighbours' values arrive is the critical path: the lower
    // coefficients and neighbour indices, and the products upper[f]*psiOld[u[f]] (subtracted AFTER the lower terms, in face
    // order, as the reference does; forming the products early does not change them).  Up to kPre faces per side in
    // registers, the rest (polyhedral cells) in the loops below.
    constexpr int kPre = 4;
    double lc[kPre], up[kPre];
    int ln[kPre];
#pragma unroll
    for (int t = 0; t < kPre; t++)
    {
        lc[t] = up[t] = 0.0;
        ln[t] = 0;
        if (k0 + t < k1)
        {
            const int f = M.losort[k0 + t];
            lc[t] = M.lower[f];
            ln[t] = M.l[f];
        }
        if (o0 + t < o1) up[t] = M.upper[o0 + t] * psiOld[M.u[o0 + t]];
    }
    auto wait_for = [&](const double* p) {
        double v = ld_relaxed(p);
        for (long long tries = 0; is_sentinel(v); tries++)
        {
            if (tries >= 32) __nanosleep(tries > 4096 ? 1000 : 100);
            if ((tries & 1023) == 1023 && *(volatile int*)err) break;
            if (tries >= (1ll << 22))
            {
                atomicExch(err, 1);
                break;
            }
            v = ld_relaxed(p);
        }
        return v;
    };
#pragma unroll
    for (int t = 0; t < kPre; t++)
        if (k0 + t < k1) acc -= lc[t] * wait_for(psiNew + ln[t]);
    for (int k = k0 + kPre; k < k1; k++)
    {
        const int f = M.losort[k];
        acc -= M.lower[f] * wait_for(psiNew + M.l[f]);
    }
#pragma unroll
    for (int t = 0; t < kPre; t++)
        if (o0 + t < o1) acc -= up[t];
    for (int f = o0 + kPre; f < o1; f++) acc -= M.upper[f] * psiOld[M.u[f]];
    st_relaxed(psiNew + row, acc / rd);
}

// OP 0: y = A x (lduMatrix::Amul order: diag, lower neighbours ascending, owner faces);  OP 1: y = b - A x in the order of
// lduMatrix::residual (b - diag*x, then the same face order, subtracting)
template <int OP>
__global__ void __launch_bounds__(256) k_gs_amul(GsDev M, const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ y)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= M.n) return;
    double acc = OP ? b[c] - M.diag[c] * x[c] : M.diag[c] * x[c];
    for (int k = M.losortStart[c]; k < M.losortStart[c + 1]; k++)
    {
        const int f = M.losort[k];
        acc = OP ? acc - M.lower[f] * x[M.l[f]] : acc + M.lower[f] * x[M.l[f]];
    }
    for (int f = M.ownerStart[c]; f < M.ownerStart[c + 1]; f++) acc = OP ? acc - M.upper[f] * x[M.u[f]] : acc + M.upper[f] * x[M.u[f]];
    y[c] = acc;
}

// OP 0: sum a; 1: sum |a|; 2: sum |a - c| + |b - c|; 3: sum |b - a|   (fixed grid, fixed tree: run-to-run identical)
template <int OP>
__global__ void __launch_bounds__(256) k_gs_reduce(int n, const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                                                   double* partial)
{
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        acc += OP == 0 ? a[i] : OP == 1 ? fabs(a[i]) : OP == 2 ? fabs(a[i] - c[i]) + fabs(b[i] - c[i]) : fabs(b[i] - a[i]);
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1)
    {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) k_gs_reduce_final(const double* __restrict__ partial, int nBlocks, double* out)
{
    __shared__ double sh[256];
    double acc = 0.0;
    for (int k = threadIdx.x; k < nBlocks; k += 256) acc += partial[k];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1)
    {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}
__global__ void k_gs_fill(long long n, double* p, double v)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
} // namespace b200

struct b200_gs
{
    b200_ctx* ctx = nullptr;
    int n = 0, nf = 0, nPos = 0, nLevels = 0;
    DevBuf<int> l, u, losort, losortStart, ownerStart, rows;
    DevBuf<double> diag, upper, lower, psi, psi2, b, tmp, tmp2, partial, red;
    bool symmetric = true, haveCoeffs = false;
    DevBuf<unsigned> ticket;
    unsigned ticketBase = 0;
    DevBuf<int> devErr;
};

namespace
{
GsDev gs_dev(const b200_gs* s)
{
    return GsDev{s->n, s->nf, s->l.p, s->u.p, s->losort.p, s->losortStart.p, s->ownerStart.p, s->diag.p, s->upper.p, s->symmetric ? s->upper.p : s->lower.p};
}

int gs_check_err(b200_gs* s, const char* what)
{
    int e = 0;
    b200_ctx* ctx = s->ctx;
    CK(ctx, cudaMemcpyAsync(&e, s->devErr.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (e)
    {
        cudaMemsetAsync(s->devErr.p, 0, sizeof(int), ctx->stream);
        return set_err(ctx, B200_EDEVICE, "%s: a Gauss-Seidel sweep timed out waiting for a dependency", what);
    }
    return B200_OK;
}

// one sweep on the device: psi (in) -> other buffer (out); the buffers are swapped
int gs_sweep_dev(b200_gs* s, const double* bPrime)
{
    b200_ctx* ctx = s->ctx;
    if (s->n == 0) return B200_OK;
    ctx->launches += 2;
    k_blk_fill_sentinel<<<blk_grid(s->n, 256), 256, 0, ctx->stream>>>(s->psi2.p, s->n);
    const unsigned ctas = (unsigned)((s->nPos + kGsThreads - 1) / kGsThreads);
    k_gs_sweep<<<ctas, kGsThreads, 0, ctx->stream>>>(gs_dev(s), s->rows.p, s->nPos, bPrime, s->psi.p, s->psi2.p, s->ticket.p, s->ticketBase, s->devErr.p);
    s->ticketBase += ctas;
    CK(ctx, cudaGetLastError());
    std::swap(s->psi.p, s->psi2.p);
    return B200_OK;
}

template <int OP>
int gs_reduce(b200_gs* s, const double* a, const double* b, const double* c, double* out)
{
    b200_ctx* ctx = s->ctx;
    const int blocks = blk_grid(s->n, 256, kGsRedBlocks);
    ctx->launches += 2;
    k_gs_reduce<OP><<<blocks, 256, 0, ctx->stream>>>(s->n, a, b, c, s->partial.p);
    k_gs_reduce_final<<<1, 256, 0, ctx->stream>>>(s->partial.p, blocks, s->red.p);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(out, s->red.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}
} // namespace

extern "C" int b200_gs_create(b200_ctx* ctx, int32_t nCells, int32_t nFaces, const int32_t* lowerAddr, const int32_t* upperAddr, b200_gs** out)
{
    if (!ctx || !out || nCells < 0 || nFaces < 0 || (nFaces && (!lowerAddr || !upperAddr))) return B200_EINVAL;
    for (int f = 0; f < nFaces; f++)
    {
        if (lowerAddr[f] < 0 || upperAddr[f] >= nCells || lowerAddr[f] >= upperAddr[f])
            return set_err(ctx, B200_EINVAL, "b200_gs_create: face %d (%d, %d) is not upper-triangular", f, lowerAddr[f], upperAddr[f]);
        if (f && (lowerAddr[f] < lowerAddr[f - 1] || (lowerAddr[f] == lowerAddr[f - 1] && upperAddr[f] <= upperAddr[f - 1])))
            return set_err(ctx, B200_EINVAL, "b200_gs_create: faces are not in upper-triangular order at face %d", f);
    }
    CK(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<b200_gs> s(new b200_gs);
    s->ctx = ctx;
    s->n = nCells;
    s->nf = nFaces;
    std::vector<int> l(lowerAddr, lowerAddr + nFaces), u(upperAddr, upperAddr + nFaces);
    std::vector<int> losortStart((size_t)nCells + 1, 0), ownerStart((size_t)nCells + 1, 0), losort((size_t)nFaces);
    for (int f = 0; f < nFaces; f++)
    {
        losortStart[(size_t)u[f] + 1]++;
        ownerStart[(size_t)l[f] + 1]++;
    }
    for (int c = 0; c < nCells; c++)
    {
        losortStart[(size_t)c + 1] += losortStart[c];
        ownerStart[(size_t)c + 1] += ownerStart[c];
    }
    {
        std::vector<int> fill(losortStart.begin(), losortStart.end() - 1);
        for (int f = 0; f < nFaces; f++) losort[(size_t)fill[u[f]]++] = f;
    }
    // wavefront levels of the lower-neighbour graph; every level padded to a whole warp, so that no thread polls a value
    // a thread of its own warp produces
    std::vector<int> lev((size_t)nCells, 0);
    int nLev = nCells ? 1 : 0;
    for (int f = 0; f < nFaces; f++)
    {
        lev[u[f]] = std::max(lev[u[f]], lev[l[f]] + 1);
        nLev = std::max(nLev, lev[u[f]] + 1);
    }
    std::vector<long long> start((size_t)nLev + 1, 0);
    for (int c = 0; c < nCells; c++) start[(size_t)lev[c] + 1]++;
    for (int k = 0; k < nLev; k++) start[(size_t)k + 1] = start[k] + (start[(size_t)k + 1] + 31) / 32 * 32;
    std::vector<int> rows((size_t)start[nLev], -1);
    {
        std::vector<long long> fill(start.begin(), start.end() - 1);
        for (int c = 0; c < nCells; c++) rows[(size_t)fill[lev[c]]++] = c;
    }
    s->nPos = (int)rows.size();
    s->nLevels = nLev;
    cudaStream_t st = ctx->stream;
    CK(ctx, s->l.upload(l, st));
    CK(ctx, s->u.upload(u, st));
    CK(ctx, s->losort.upload(losort, st));
    CK(ctx, s->losortStart.upload(losortStart, st));
    CK(ctx, s->ownerStart.upload(ownerStart, st));
    CK(ctx, s->rows.upload(rows, st));
    for (DevBuf<double>* v : {&s->psi, &s->psi2, &s->b, &s->tmp, &s->tmp2, &s->diag}) CK(ctx, v->alloc((size_t)nCells));
    CK(ctx, s->upper.alloc((size_t)nFaces));
    CK(ctx, s->partial.alloc(kGsRedBlocks));
    CK(ctx, s->red.alloc(8));
    CK(ctx, s->ticket.alloc(1));
    CK(ctx, s->devErr.alloc(1));
    CK(ctx, cudaMemsetAsync(s->ticket.p, 0, sizeof(unsigned), st));
    CK(ctx, cudaMemsetAsync(s->devErr.p, 0, sizeof(int), st));
    CK(ctx, cudaStreamSynchronize(st));
    *out = s.release();
    return B200_OK;
}

extern "C" int b200_gs_destroy(b200_gs* s)
{
    if (!s) return B200_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    delete s;
    return B200_OK;
}

extern "C" int b200_gs_set_coeffs(b200_gs* s, const double* diag, const double* upper, const double* lower)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if ((s->n && !diag) || (s->nf && !upper)) return set_err(ctx, B200_EINVAL, "b200_gs_set_coeffs: null array");
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (s->n) CK(ctx, cudaMemcpyAsync(s->diag.p, diag, sizeof(double) * (size_t)s->n, cudaMemcpyHostToDevice, st));
    if (s->nf) CK(ctx, cudaMemcpyAsync(s->upper.p, upper, sizeof(double) * (size_t)s->nf, cudaMemcpyHostToDevice, st));
    s->symmetric = lower == nullptr;
    if (lower)
    {
        CK(ctx, s->lower.alloc((size_t)s->nf));
        if (s->nf) CK(ctx, cudaMemcpyAsync(s->lower.p, lower, sizeof(double) * (size_t)s->nf, cudaMemcpyHostToDevice, st));
    }
    CK(ctx, cudaStreamSynchronize(st));
    s->haveCoeffs = true;
    return B200_OK;
}

static int gs_smooth_impl(b200_gs* s, double* psi, const double* source, int nSweeps, const char* what)
{
    if (!s || nSweeps < 0 || (s->n && (!psi || !source))) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->haveCoeffs) return set_err(ctx, B200_ESTATE, "%s: coefficients not set", what);
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t nb = sizeof(double) * (size_t)s->n;
    if (nb)
    {
        CK(ctx, cudaMemcpyAsync(s->psi.p, psi, nb, cudaMemcpyHostToDevice, ctx->stream));
        CK(ctx, cudaMemcpyAsync(s->b.p, source, nb, cudaMemcpyHostToDevice, ctx->stream));
    }
    for (int k = 0; k < nSweeps; k++)
    {
        int rc = gs_sweep_dev(s, s->b.p);
        if (rc) return rc;
    }
    if (nb) CK(ctx, cudaMemcpyAsync(psi, s->psi.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    return gs_check_err(s, what);
}

extern "C" int b200_gs_sweep(b200_gs* s, double* psi, const double* bPrime) { return gs_smooth_impl(s, psi, bPrime, 1, "b200_gs_sweep"); }
extern "C" int b200_gs_smooth(b200_gs* s, double* psi, const double* source, int nSweeps)
{
    return gs_smooth_impl(s, psi, source, nSweeps, "b200_gs_smooth");
}

extern "C" int b200_gs_solve(b200_gs* s, const b200_solver_opts* o, int nSweeps, double* psi, const double* source, b200_perf* perf,
                             double* history, int cap)
{
    if (!s || !o || !perf || nSweeps <= 0 || (s->n && (!psi || !source))) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->haveCoeffs) return set_err(ctx, B200_ESTATE, "b200_gs_solve: coefficients not set");
    CK(ctx, cudaSetDevice(ctx->device));
    memset(perf, 0, sizeof(*perf));
    const size_t nb = sizeof(double) * (size_t)s->n;
    if (nb)
    {
        CK(ctx, cudaMemcpyAsync(s->psi.p, psi, nb, cudaMemcpyHostToDevice, ctx->stream));
        CK(ctx, cudaMemcpyAsync(s->b.p, source, nb, cudaMemcpyHostToDevice, ctx->stream));
    }
    cudaEvent_t e0, e1;
    CK(ctx, cudaEventCreate(&e0));
    CK(ctx, cudaEventCreate(&e1));
    CK(ctx, cudaEventRecord(e0, ctx->stream));
    const int grid = blk_grid(s->n, 256, 1 << 30);
    const GsDev M = gs_dev(s);
    int rc = B200_OK;
    auto stop = [&]() {
        if (perf->nIterations < o->minIter) return false;
        perf->converged = (perf->finalResidual < o->tolerance || (o->relTol > B200_SMALL_ && perf->finalResidual <= o->relTol * perf->initialResidual)) ? 1 : 0;
        return perf->nIterations >= o->maxIter || perf->converged;
    };
    do
    {
        // normFactor: Ax = A psi, pA = A (gAverage(psi) * 1), sum |Ax - pA| + |b - pA| + SMALL; initial residual sum |b - Ax|
        double sum = 0.0, nt = 0.0, r0 = 0.0;
        if ((rc = gs_reduce<0>(s, s->psi.p, nullptr, nullptr, &sum))) break;
        ctx->launches += 3;
        k_gs_fill<<<blk_grid(s->n, 256), 256, 0, ctx->stream>>>(s->n, s->tmp.p, sum / (double)(s->n ? s->n : 1));
        k_gs_amul<0><<<grid, 256, 0, ctx->stream>>>(M, s->tmp.p, nullptr, s->tmp2.p); // pA
        k_gs_amul<0><<<grid, 256, 0, ctx->stream>>>(M, s->psi.p, nullptr, s->tmp.p);  // Ax
        if ((rc = gs_reduce<2>(s, s->tmp.p, s->b.p, s->tmp2.p, &nt))) break;
        if ((rc = gs_reduce<3>(s, s->tmp.p, s->b.p, nullptr, &r0))) break;
        perf->normFactor = nt + B200_SMALL;
        perf->initialResidual = perf->finalResidual = r0 / perf->normFactor;
        int k = 0;
        if (history && k < cap) history[k] = perf->initialResidual;
        if (stop()) break;
        do
        {
            for (int q = 0; q < nSweeps && !rc; q++) rc = gs_sweep_dev(s, s->b.p);
            if (rc) break;
            ctx->launches++;
            k_gs_amul<1><<<grid, 256, 0, ctx->stream>>>(M, s->psi.p, s->b.p, s->tmp.p); // lduMatrix::residual
            double sr = 0.0;
            if ((rc = gs_reduce<1>(s, s->tmp.p, nullptr, nullptr, &sr))) break;
            perf->finalResidual = sr / perf->normFactor;
            perf->nIterations += nSweeps;
            k++;
            if (history && k < cap) history[k] = perf->finalResidual;
        } while (!stop());
    } while (false);
    if (rc) return rc;
    CK(ctx, cudaEventRecord(e1, ctx->stream));
    if (nb) CK(ctx, cudaMemcpyAsync(psi, s->psi.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    rc = gs_check_err(s, "b200_gs_solve");
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    perf->deviceMs = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}
