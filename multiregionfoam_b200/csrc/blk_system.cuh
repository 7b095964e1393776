// blk_system.cuh -- block-coupled (vector4) LDU systems: kernels, host orchestration and the C ABI of
// include/b200_blk.h.  Included at the end of b200_ldu.cu (one translation unit: shares b200_ctx, CK, DevBuf).
//
// Replaces, for fvBlockMatrix<vector4>::solve (/root/reference/filesToReplace/fvBlockMatrix.C:1360-1388; SURVEY 8
// a18-a19), foam-extend's BlockLduMatrix<vector4>::Amul, BlockCholeskyPrecon / BlockDiagonalPrecon / BlockNoPrecon and
// BlockBiCGStabSolver / BlockCGSolver.  FP64, -fmad=false: Amul, the preconditioner diagonal and precondition are
// bit-identical to oracle/blk_oracle.c; reductions are fixed-shape trees.
//
// Layout: everything stays in the caller's order - fields [N][4], coefficients in FACE order with their active type
// (1 / 4 / 16 doubles per entry).  One QUAD of threads per matrix row, thread i of the quad owns component i: it reads
// row i of every 4x4 coefficient (32 contiguous bytes; a quad reads the whole 128-byte block, consecutive quads
// consecutive blocks), so no repacking pass is needed and every access is a full sector.  A row's terms are applied
// in the reference order: diagonal, lower neighbours by ascending face (losort), upper neighbours by ascending face.
//
// Sweeps (BlockCholesky ILUmultiply, calcPreconDiag): rows are visited in wavefront-LEVEL order (host tables, levels
// padded to whole warps so that a warp never holds two dependent rows); the output vector is its own ready flag
// (sentinel pre-fill, as in the scalar sweeps) and CTAs take tickets, so a row only ever waits for rows of CTAs that
// already run.  Block rows carry 16x the coefficient bytes of a scalar row per dependency, which is why this simple
// scheme is bandwidth- rather than latency-limited on 3-D meshes; the line-pipelined schedule of the scalar sweeps is
// the next step for 2-D meshes.
#pragma once

#include "../../include/b200_blk.h"

namespace b200
{

constexpr int kBlkThreads = 256;          // 64 rows per CTA
constexpr int kBlkRowsPerCta = kBlkThreads / 4;
constexpr int kBlkRedBlocks = 1184;       // 8 x 148: fixed reduction grid (deterministic partial sums)
constexpr int kBlkMaxRed = 5;

__device__ __forceinline__ void ld_cg2(const double* p, double& a, double& b)
{
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}

// Wait for the K values at p (K = 1, 4, 16; written once by another row) to leave the sentinel state.
template <int K>
__device__ __forceinline__ bool blk_poll(const double* p, double* v, int* err)
{
    for (long long tries = 0; tries < (1ll << 22); tries++)
    {
        bool ok = true;
        if (K == 1)
        {
            v[0] = ld_relaxed(p);
            ok = !is_sentinel(v[0]);
        }
        else
        {
#pragma unroll
            for (int j = 0; j < K; j += 2)
            {
                ld_cg2(p + j, v[j], v[j + 1]);
                ok = ok && !is_sentinel(v[j]) && !is_sentinel(v[j + 1]);
            }
        }
        if (ok) return true;
        if (tries >= 32) __nanosleep(tries > 4096 ? 1000 : 100);
        if ((tries & 1023) == 1023 && *(volatile int*)err) break;
    }
    atomicExch(err, 1);
    return false;
}

// x[i] without dynamic register indexing (which would spill the vector to local memory)
__device__ __forceinline__ double blk_sel4(const double* x, int i) { return i == 0 ? x[0] : i == 1 ? x[1] : i == 2 ? x[2] : x[3]; }

// component i of mult(coeff, x) for a coefficient of the given kind (BlockCoeff<Type>::multiply); tr: transposed
// square coefficient (symmetric matrices)
__device__ __forceinline__ double blk_mult_row(int kind, const double* __restrict__ a, bool tr, const double* x, int i)
{
    if (kind == 1) return __ldg(a) * blk_sel4(x, i);
    if (kind == 4) return __ldg(a + i) * blk_sel4(x, i);
    double r0, r1, r2, r3;
    if (tr)
    {
        r0 = __ldg(a + i);
        r1 = __ldg(a + 4 + i);
        r2 = __ldg(a + 8 + i);
        r3 = __ldg(a + 12 + i);
    }
    else
    {
        const double2 p = __ldg(reinterpret_cast<const double2*>(a + 4 * i));
        const double2 q = __ldg(reinterpret_cast<const double2*>(a + 4 * i + 2));
        r0 = p.x;
        r1 = p.y;
        r2 = q.x;
        r3 = q.y;
    }
    double sum = r0 * x[0];
    sum += r1 * x[1];
    sum += r2 * x[2];
    sum += r3 * x[3];
    return sum;
}

// the same with the coefficient row already in registers (kind 16: c = row i; kind 4 / 1: c[0] = the entry)
__device__ __forceinline__ double blk_mult_reg(int kind, const double* c, const double* x, int i)
{
    if (kind != 16) return c[0] * blk_sel4(x, i);
    double sum = c[0] * x[0];
    sum += c[1] * x[1];
    sum += c[2] * x[2];
    sum += c[3] * x[3];
    return sum;
}

__device__ __forceinline__ void blk_load_row(int kind, const double* a, int i, double* c)
{
    if (kind == 16)
    {
        c[0] = a[4 * i];
        c[1] = a[4 * i + 1];
        c[2] = a[4 * i + 2];
        c[3] = a[4 * i + 3];
    }
    else
        c[0] = kind == 4 ? a[i] : a[0];
}

// row i (tr: column i) of a coefficient into registers, for blk_mult_reg
__device__ __forceinline__ void blk_load_row_tr(int kind, const double* __restrict__ a, bool tr, int i, double* c)
{
    if (kind != 16)
    {
        c[0] = kind == 4 ? __ldg(a + i) : __ldg(a);
        c[1] = c[2] = c[3] = 0.0;
    }
    else if (tr)
    {
        c[0] = __ldg(a + i);
        c[1] = __ldg(a + 4 + i);
        c[2] = __ldg(a + 8 + i);
        c[3] = __ldg(a + 12 + i);
    }
    else
    {
        const double2 p = __ldg(reinterpret_cast<const double2*>(a + 4 * i));
        const double2 q = __ldg(reinterpret_cast<const double2*>(a + 4 * i + 2));
        c[0] = p.x;
        c[1] = p.y;
        c[2] = q.x;
        c[3] = q.y;
    }
}

__device__ __forceinline__ void blk_load_x(const double* __restrict__ x, long long c, double* v)
{
    const double2 p = __ldg(reinterpret_cast<const double2*>(x + 4 * c));
    const double2 q = __ldg(reinterpret_cast<const double2*>(x + 4 * c + 2));
    v[0] = p.x;
    v[1] = p.y;
    v[2] = q.x;
    v[3] = q.y;
}

struct BlkDev
{
    int n, nf;
    const int *l, *u, *losort, *losortStart, *ownerStart;
    int dK, uK, lK; // lK == 0: symmetric (lower = transposed upper)
    const double *diag, *upper, *lower;
};

// ---------------------------------------------------------------------------------------------- Amul
// BlockLduMatrix<Type>::AmulCore: y = D x; y[u] += L[f] x[l] (all faces); y[l] += U[f] x[u] (all faces).
__global__ void __launch_bounds__(kBlkThreads) k_blk_amul(BlkDev M, const double* __restrict__ x, double* __restrict__ y)
{
    const long long t = (long long)blockIdx.x * kBlkThreads + threadIdx.x;
    const long long c = t >> 2;
    const int i = (int)(t & 3);
    if (c >= M.n) return;
    double xv[4];
    blk_load_x(x, c, xv);
    double acc = blk_mult_row(M.dK, M.diag + (size_t)c * M.dK, false, xv, i);
    for (int k = M.losortStart[c]; k < M.losortStart[c + 1]; k++)
    {
        const int f = M.losort[k];
        blk_load_x(x, M.l[f], xv);
        acc += M.lK ? blk_mult_row(M.lK, M.lower + (size_t)f * M.lK, false, xv, i) : blk_mult_row(M.uK, M.upper + (size_t)f * M.uK, true, xv, i);
    }
    for (int f = M.ownerStart[c]; f < M.ownerStart[c + 1]; f++)
    {
        blk_load_x(x, M.u[f], xv);
        acc += blk_mult_row(M.uK, M.upper + (size_t)f * M.uK, false, xv, i);
    }
    y[t] = acc;
}

// ---------------------------------------------------------------------------------------------- sweeps
// ILUmultiply of BlockCholeskyPrecon.  BWD = false: out[row] = D b[row] - sum_{lower faces, ascending} D (L x[l]);
// BWD = true: out[row] = a[row] - sum_{owner faces, DESCENDING} D (U x[u]).  rows: level-ordered row list (-1 = pad).
#ifndef B200_BLK_MINCTAS
#define B200_BLK_MINCTAS 2 // resident CTAs per SM the sweep is compiled for (register cap 65536 / (256 * this))
#endif
#ifndef B200_BLK_CHUNK
#define B200_BLK_CHUNK 3 // terms of a row whose coefficients are loaded before the first poll (a hex cell has 3 per sweep)
#endif
#ifndef B200_BLK_SWEEP_THREADS
#define B200_BLK_SWEEP_THREADS 256 // threads of a sweep CTA (4 per row)
#endif
#ifndef B200_BLK_PERSIST
#define B200_BLK_PERSIST 0 // 1: a CTA works through chunk after chunk and fetches the next ticket while the current chunk runs
#endif
constexpr int kBlkSweepThreads = B200_BLK_SWEEP_THREADS;
constexpr int kBlkSweepRows = kBlkSweepThreads / 4;

// the rows of one chunk (ticket) of a sweep: one thread quad per row
// Position-ordered copies for the Cholesky sweeps (built with the preconditioner): pack = the coefficient blocks of the first
// kBlkDirect terms of every sweep position ([pos][kBlkDirect][uK], zero where a row has fewer; for the forward sweep of a
// symmetric matrix already transposed), pDpos = the preconditioner diagonal, nb12 = the neighbour rows of terms 1 and 2 (term
// 0's sits in meta.w).  Their addresses follow from the position alone, so they are requested together with the position
// table, not after it: the chain of dependent loads of a row is  ticket -> {meta, coefficients, diagonal} -> neighbour values.
constexpr int kBlkDirect = 3;
struct BlkPack
{
    const double* pack;  // nullptr: no packed copies (BlockDiagonalPrecon), everything through meta / terms
    const double* pDpos;
    const int2* nb12;
};

template <bool BWD>
__device__ __forceinline__ void blk_sweep_chunk(const BlkDev& M, int pK, int withFaces, const double* __restrict__ pD, const int4* __restrict__ meta,
                                                const int2* __restrict__ terms, const BlkPack& P, int nPos, const double* __restrict__ a,
                                                double* out, int* err, unsigned chunk)
{
    constexpr int kChunk = B200_BLK_CHUNK;
    static_assert(kChunk == kBlkDirect, "the packed first batch holds kBlkDirect terms");
    const long long pos = (long long)chunk * kBlkSweepRows + (threadIdx.x >> 2);
    if (pos >= nPos) return;
    const int i = threadIdx.x & 3;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned qm = 0xFu << (lane & ~3u);
    const int qb = (int)(lane & ~3u);
    double d[4], xv[4], tv[4];
    double cf[kChunk][4];
    int nbv[kChunk];
    int2 nb12v = make_int2(0, 0);
    const bool direct = P.pack != nullptr;
    if (direct)
    { // independent of the position table
#pragma unroll
        for (int t = 0; t < kChunk; t++) blk_load_row_tr(M.uK, P.pack + ((size_t)pos * kBlkDirect + t) * M.uK, false, i, cf[t]);
        blk_load_row(pK, P.pDpos + (size_t)pos * pK, i, d);
        nb12v = __ldg(P.nb12 + pos);
    }
    // {row, number of terms, first term, neighbour row of term 0}
    const int4 mt = __ldg(meta + pos);
    const int row = mt.x;
    if (row < 0) return;
    if (!direct) blk_load_row(pK, pD + (size_t)row * pK, i, d);
    double acc;
    if (!BWD)
    {
        blk_load_x(a, row, xv);
        acc = blk_mult_reg(pK, d, xv, i);
    }
    else
        acc = a[4ll * row + i];
    const int k0 = mt.z;
    const int nTerms = withFaces ? mt.y : 0; // BlockDiagonalPrecon: x = mult(dDiag, b) only
    // The rows of one wavefront level wait for the level before: what a row does AFTER its neighbours' values arrive is
    // the critical path of the whole sweep.  Everything that does not depend on those values - face and neighbour
    // indices, this thread's row of each coefficient block (from DRAM) - is therefore loaded for up to kChunk terms at
    // once BEFORE the first poll; after a poll only multiply, quad shuffles and multiply remain (same order of the
    // terms and of the operations as before).
    const bool tr = !(BWD || M.lK);
    const double* const coefBase = (BWD || !M.lK) ? M.upper : M.lower;
    for (int kk0 = 0; kk0 < nTerms; kk0 += kChunk)
    {
        if (direct && kk0 == 0)
        {
#pragma unroll
            for (int t = 0; t < kChunk; t++) nbv[t] = t == 0 ? mt.w : t == 1 ? nb12v.x : nb12v.y;
        }
        else
        {
#pragma unroll
            for (int t = 0; t < kChunk; t++)
                if (kk0 + t < nTerms)
                {
                    const int2 tm = __ldg(terms + k0 + kk0 + t);
                    const int f = tm.x;
                    nbv[t] = tm.y;
                    blk_load_row_tr(M.uK, coefBase + (size_t)f * M.uK, tr, i, cf[t]);
                }
        }
        // the neighbours of a row mostly sit in the level just before it and arrive together: all of them are polled in
        // one batch of independent loads per round instead of one after the other
        double xn[kChunk][4];
        bool have[kChunk];
#pragma unroll
        for (int t = 0; t < kChunk; t++) have[t] = !(kk0 + t < nTerms);
        for (long long tries = 0;; tries++)
        {
#pragma unroll
            for (int t = 0; t < kChunk; t++)
                if (!have[t])
                {
                    ld_cg2(out + 4ll * nbv[t], xn[t][0], xn[t][1]);
                    ld_cg2(out + 4ll * nbv[t] + 2, xn[t][2], xn[t][3]);
                }
            bool all = true;
#pragma unroll
            for (int t = 0; t < kChunk; t++)
            {
                if (!have[t]) have[t] = !is_sentinel(xn[t][0]) && !is_sentinel(xn[t][1]) && !is_sentinel(xn[t][2]) && !is_sentinel(xn[t][3]);
                all = all && have[t];
            }
            if (all) break;
            if (tries >= 32) __nanosleep(tries > 4096 ? 1000 : 100);
            if ((tries & 1023) == 1023 && *(volatile int*)err) break;
            if (tries >= (1ll << 22))
            {
                atomicExch(err, 1);
                break;
            }
        }
#pragma unroll
        for (int t = 0; t < kChunk; t++)
            if (kk0 + t < nTerms)
            {
                const double ti = blk_mult_reg(M.uK, cf[t], xn[t], i);
#pragma unroll
                for (int j = 0; j < 4; j++) tv[j] = __shfl_sync(qm, ti, qb + j);
                acc -= blk_mult_reg(pK, d, tv, i);
            }
    }
    st_relaxed(out + 4ll * row + i, acc);
}

// ILUmultiply of BlockCholeskyPrecon.  BWD = false: out[row] = D b[row] - sum_{lower faces, ascending} D (L x[l]);
// BWD = true: out[row] = a[row] - sum_{owner faces, DESCENDING} D (U x[u]).  Chunks of kBlkSweepRows sweep positions are
// handed out in order by a ticket counter, so a row only ever waits for rows of CTAs that already run.
template <bool BWD>
__global__ void __launch_bounds__(kBlkSweepThreads, B200_BLK_MINCTAS * 256 / kBlkSweepThreads)
    k_blk_sweep(BlkDev M, int pK, int withFaces, const double* __restrict__ pD, const int4* __restrict__ meta, const int2* __restrict__ terms,
                BlkPack P, int nPos, const double* __restrict__ a, double* out, unsigned* ticket, unsigned ticketBase, int* err)
{
    __shared__ unsigned sTicket[2];
    if (threadIdx.x == 0) sTicket[0] = atomicAdd(ticket, 1u) - ticketBase;
    __syncthreads();
#if B200_BLK_PERSIST
    const unsigned nChunks = (unsigned)((nPos + kBlkSweepRows - 1) / kBlkSweepRows);
    int par = 0;
    for (unsigned chunk = sTicket[0]; chunk < nChunks;)
    {
        unsigned next = 0;
        if (threadIdx.x == 0) next = atomicAdd(ticket, 1u) - ticketBase; // on its way while this chunk runs
        blk_sweep_chunk<BWD>(M, pK, withFaces, pD, meta, terms, P, nPos, a, out, err, chunk);
        if (threadIdx.x == 0) sTicket[par ^ 1] = next;
        __syncthreads();
        par ^= 1;
        chunk = sTicket[par];
    }
#else
    blk_sweep_chunk<BWD>(M, pK, withFaces, pD, meta, terms, P, nPos, a, out, err, sTicket[0]);
#endif
}

// the packed copies of BlkPack
__global__ void k_blk_pack_coef(long long nPos, const int4* __restrict__ meta, const int2* __restrict__ terms, const double* __restrict__ coef, int K,
                                int tr, double* __restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nPos * kBlkDirect * K) return;
    const long long slot = t / K;
    const int j = (int)(t - slot * K);
    const long long pos = slot / kBlkDirect;
    const int k = (int)(slot - pos * kBlkDirect);
    const int4 mt = meta[pos];
    double v = 0.0;
    if (mt.x >= 0 && k < mt.y)
    {
        const int src = (tr && K == 16) ? (j & 3) * 4 + (j >> 2) : j;
        v = coef[(size_t)terms[mt.z + k].x * K + src];
    }
    out[t] = v;
}
__global__ void k_blk_pack_pd(long long nPos, const int4* __restrict__ meta, const double* __restrict__ pD, int pK, double* __restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nPos * pK) return;
    const long long pos = t / pK;
    const int row = meta[pos].x;
    out[t] = row >= 0 ? pD[(size_t)row * pK + (t - pos * pK)] : 0.0;
}

// 4x4 inverse: Gauss-Jordan with partial pivoting, operation by operation as oracle/blk_oracle.c blk_inv4
__device__ __forceinline__ void blk_inv4_dev(double (&m)[4][4], double (&r)[4][4])
{
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) r[i][j] = i == j ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        int p = k;
        double best = fabs(m[k][k]);
#pragma unroll
        for (int q = k + 1; q < 4; q++)
            if (fabs(m[q][k]) > best)
            {
                best = fabs(m[q][k]);
                p = q;
            }
#pragma unroll
        for (int q = k + 1; q < 4; q++)
            if (q == p)
            {
#pragma unroll
                for (int j = 0; j < 4; j++)
                {
                    double tm = m[k][j];
                    m[k][j] = m[q][j];
                    m[q][j] = tm;
                    tm = r[k][j];
                    r[k][j] = r[q][j];
                    r[q][j] = tm;
                }
            }
        const double piv = 1.0 / m[k][k];
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            m[k][j] *= piv;
            r[k][j] *= piv;
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            if (q == k) continue;
            const double f = m[q][k];
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                m[q][j] -= f * m[k][j];
                r[q][j] -= f * r[k][j];
            }
        }
    }
}

// row i of a coefficient of kind k expanded to SQUARE (scalar -> s I, linear -> diagonal); tr: transposed
__device__ __forceinline__ void blk_expand_row(int kind, const double* a, bool tr, int i, double* o)
{
#pragma unroll
    for (int j = 0; j < 4; j++) o[j] = 0.0;
    if (kind == 1)
        o[i] = a[0];
    else if (kind == 4)
        o[i] = a[i];
    else
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = tr ? a[4 * j + i] : a[4 * i + j];
}

// BlockCholeskyPrecon<Type>::calcPreconDiag (chol) / BlockDiagonalPrecon (!chol), square working type:
//   pD[u] = diag[u] - sum_{lower faces, ascending} (L[f] & inv(pD[l])) & U[f];  pD = inv(pD).
// Successors read the INVERTED value: inv() of the same final matrix, the same function the reference applies on the fly.
__global__ void __launch_bounds__(kBlkThreads)
    k_blk_diag16(BlkDev M, int chol, const int* __restrict__ rows, int nPos, double* pD, unsigned* ticket, unsigned ticketBase, int* err)
{
    __shared__ unsigned sTicket;
    if (threadIdx.x == 0) sTicket = atomicAdd(ticket, 1u) - ticketBase;
    __syncthreads();
    const long long pos = (long long)sTicket * kBlkRowsPerCta + (threadIdx.x >> 2);
    if (pos >= nPos) return;
    const int row = rows[pos];
    if (row < 0) return;
    const int i = threadIdx.x & 3;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned qm = 0xFu << (lane & ~3u);
    const int qb = (int)(lane & ~3u);
    double acc[4];
    blk_expand_row(M.dK, M.diag + (size_t)row * M.dK, false, i, acc);
    if (chol)
        for (int k = M.losortStart[row]; k < M.losortStart[row + 1]; k++)
        {
            const int f = M.losort[k];
            double Bi[16], A[4], AB[4], C[4];
            blk_poll<16>(pD + 16ll * M.l[f], Bi, err);
            const double* up = M.upper + (size_t)f * M.uK;
            blk_expand_row(M.uK, M.lK ? M.lower + (size_t)f * M.lK : up, M.lK == 0, i, A);
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                double sum = A[0] * Bi[j];
                sum += A[1] * Bi[4 + j];
                sum += A[2] * Bi[8 + j];
                sum += A[3] * Bi[12 + j];
                AB[j] = sum;
            }
            double T[4];
#pragma unroll
            for (int kq = 0; kq < 4; kq++)
            { // row kq of expand(upper[f]) (every thread needs all of C)
                blk_expand_row(M.uK, up, false, kq, C);
#pragma unroll
                for (int j = 0; j < 4; j++) T[j] = kq == 0 ? AB[0] * C[j] : T[j] + AB[kq] * C[j];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) acc[j] -= T[j];
        }
    double m[4][4], r[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int j = 0; j < 4; j++) m[q][j] = __shfl_sync(qm, acc[j], qb + q);
    blk_inv4_dev(m, r);
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (q == i)
#pragma unroll
            for (int j = 0; j < 4; j++) st_relaxed(pD + 16ll * row + 4 * q + j, r[q][j]);
}

// component-wise working type (pK = 1 or 4): raw[u] = diag[u] - sum (a*c)/raw[l]; pD = 1/raw.  The reference divides
// by the NOT yet inverted value, so successors poll the raw buffer.
__global__ void __launch_bounds__(kBlkThreads) k_blk_diag_cmpt(BlkDev M, int chol, int pK, const int* __restrict__ rows, int nPos, double* raw,
                                                                double* pD, unsigned* ticket, unsigned ticketBase, int* err)
{
    __shared__ unsigned sTicket;
    if (threadIdx.x == 0) sTicket = atomicAdd(ticket, 1u) - ticketBase;
    __syncthreads();
    const long long pos = (long long)sTicket * kBlkRowsPerCta + (threadIdx.x >> 2);
    if (pos >= nPos) return;
    const int row = rows[pos];
    if (row < 0) return;
    const int i = threadIdx.x & 3;
    const int ci = pK == 4 ? i : 0;
    const double* dg = M.diag + (size_t)row * M.dK;
    double acc = M.dK == 1 ? dg[0] : dg[i];
    if (chol)
        for (int k = M.losortStart[row]; k < M.losortStart[row + 1]; k++)
        {
            const int f = M.losort[k];
            double b;
            blk_poll<1>(raw + (size_t)pK * M.l[f] + ci, &b, err);
            const double* up = M.upper + (size_t)f * M.uK;
            const double* lo = M.lK ? M.lower + (size_t)f * M.lK : up;
            const double a = M.uK == 1 ? lo[0] : lo[i], c = M.uK == 1 ? up[0] : up[i];
            acc -= (a * c) / b;
        }
    if (pK == 4 || i == 0)
    {
        st_relaxed(raw + (size_t)pK * row + ci, acc);
        pD[(size_t)pK * row + ci] = 1.0 / acc;
    }
}

// ---------------------------------------------------------------------------------------------- vector kernels
__global__ void k_blk_fill_sentinel(double* p, long long n)
{
    const double s = sentinel();
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) p[t] = s;
}

__global__ void k_blk_fill4(double* p, long long nCells, double v0, double v1, double v2, double v3)
{
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += (long long)gridDim.x * blockDim.x)
    {
        p[4 * c] = v0;
        p[4 * c + 1] = v1;
        p[4 * c + 2] = v2;
        p[4 * c + 3] = v3;
    }
}

// y = a - b (initial residual), p = r + beta p - beta omega v, s = r - alpha v, x += alpha ph + omega sh / r = s - omega t,
// pA = wA + beta pA, x += alpha pA / rA -= alpha wA: expression by expression as the reference's forAll loops
enum BlkVecOp
{
    BV_SUB = 0,
    BV_P,
    BV_S,
    BV_XR,
    BV_CG_P,
    BV_CG_XR,
};

template <int OP>
__global__ void k_blk_vec(long long n4, double* y, double* z, const double* __restrict__ a, const double* __restrict__ b,
                          const double* __restrict__ c, const double* __restrict__ d, double s1, double s2)
{
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x)
    {
        if (OP == BV_SUB) y[t] = a[t] - b[t];
        if (OP == BV_P) y[t] = a[t] + s1 * y[t] - s1 * s2 * b[t];
        if (OP == BV_S) y[t] = a[t] - s1 * b[t];
        if (OP == BV_XR)
        { // x = x + alpha ph + omega sh; r = s - omega t
            y[t] = y[t] + s1 * a[t] + s2 * b[t];
            z[t] = c[t] - s2 * d[t];
        }
        if (OP == BV_CG_P) y[t] = a[t] + s1 * y[t];
        if (OP == BV_CG_XR)
        {
            y[t] += s1 * a[t];
            z[t] -= s1 * b[t];
        }
    }
}

// reductions over cells, NV values per cell, fixed grid and fixed tree: partial[v * gridDim.x + block]
enum BlkRedOp
{
    BR_PROD = 0,  // (a & b)
    BR_PROD2,     // (a & b), (a & a)
    BR_CMPTMAG,   // |a_0| .. |a_3|
    BR_SUM4,      // a_0 .. a_3
    BR_NORMTERMS, // mag(a - b) + mag(c - b)   (a = wA, b = pA, c = source)
};

template <int OP>
__global__ void __launch_bounds__(256) k_blk_reduce(long long nCells, const double* __restrict__ a, const double* __restrict__ b,
                                                    const double* __restrict__ c, double* partial)
{
    constexpr int NV = OP == BR_PROD ? 1 : OP == BR_PROD2 ? 2 : OP == BR_NORMTERMS ? 1 : 4;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) acc[v] = 0.0;
    for (long long cell = (long long)blockIdx.x * 256 + threadIdx.x; cell < nCells; cell += (long long)gridDim.x * 256)
    {
        double av[4], bv[4], cv[4];
        blk_load_x(a, cell, av);
        if (OP == BR_PROD || OP == BR_PROD2 || OP == BR_NORMTERMS) blk_load_x(b, cell, bv);
        if (OP == BR_NORMTERMS) blk_load_x(c, cell, cv);
        if (OP == BR_PROD || OP == BR_PROD2)
        {
            double dsum = av[0] * bv[0];
            dsum += av[1] * bv[1];
            dsum += av[2] * bv[2];
            dsum += av[3] * bv[3];
            acc[0] += dsum;
        }
        if (OP == BR_PROD2)
        {
            double dsum = av[0] * av[0];
            dsum += av[1] * av[1];
            dsum += av[2] * av[2];
            dsum += av[3] * av[3];
            acc[1] += dsum;
        }
        if (OP == BR_CMPTMAG)
#pragma unroll
            for (int v = 0; v < 4; v++) acc[v < NV ? v : 0] += fabs(av[v]);
        if (OP == BR_SUM4)
#pragma unroll
            for (int v = 0; v < 4; v++) acc[v < NV ? v : 0] += av[v];
        if (OP == BR_NORMTERMS)
        {
            double q1 = 0.0, q2 = 0.0;
#pragma unroll
            for (int v = 0; v < 4; v++)
            {
                const double d1 = av[v] - bv[v], d2 = cv[v] - bv[v];
                q1 = v == 0 ? d1 * d1 : q1 + d1 * d1;
                q2 = v == 0 ? d2 * d2 : q2 + d2 * d2;
            }
            acc[0] += sqrt(q1) + sqrt(q2);
        }
    }
    __shared__ double sh[NV][256];
#pragma unroll
    for (int v = 0; v < NV; v++) sh[v][threadIdx.x] = acc[v];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1)
    {
        if ((int)threadIdx.x < w)
#pragma unroll
            for (int v = 0; v < NV; v++) sh[v][threadIdx.x] += sh[v][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0)
#pragma unroll
        for (int v = 0; v < NV; v++) partial[(size_t)v * gridDim.x + blockIdx.x] = sh[v][0];
}

__global__ void __launch_bounds__(256) k_blk_reduce_final(const double* __restrict__ partial, int nBlocks, int nv, double* out, PeerAR ar, int* err)
{
    __shared__ double sh[256];
    for (int v = 0; v < nv; v++)
    {
        double acc = 0.0;
        for (int k = threadIdx.x; k < nBlocks; k += 256) acc += partial[(size_t)v * nBlocks + k];
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1)
        {
            if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[v] = sh[0];
        __syncthreads();
    }
    // gSumProd / gSum: the partial sums of the ranks, added in rank order on every rank (kernels.cuh, peer_allreduce)
    if (ar.nranks > 1) peer_allreduce(ar, nv, out, err);
}

// ---------------------------------------------------------------------------------------------- coupled patches
// BlockLduMatrix<Type>::initInterfaces for a processor patch: the patch-internal field x[faceCells] (4 doubles per face)
// goes straight into the neighbour's receive buffer; the last block publishes the sequence number (cf. k_halo_push).
__global__ void k_blk_halo_push(int nFaces, const int* __restrict__ cells, const double* __restrict__ x, double* __restrict__ remote,
                                unsigned long long* remoteFlag, unsigned long long seq, unsigned* counter)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 4ll * nFaces) remote[t] = x[4ll * cells[t >> 2] + (t & 3)];
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

// BlockLduMatrix<Type>::updateInterfaces(coupleUpper, Ax, x) for one processor patch (processorFvPatchField<Type>::
// updateInterfaceMatrix with switchToLhs = false):  Ax[faceCells[f]] -= coupleUpper[f] * xNbr[f], faces in patch order.
// One thread quad per touched cell (rows = the cells of the patch, each with its faces in patch order).  xNbr comes from
// the receive buffer (nbr, after the neighbour's flag shows this exchange) or, for a pair of patches on the same rank,
// from x[nbrCells[f]].
__global__ void __launch_bounds__(kBlkThreads)
    k_blk_iface(int nRows, const int* __restrict__ rowCell, const int* __restrict__ rowStart, const int* __restrict__ rowFace, int kind,
                const double* __restrict__ coef, const double* nbr, const int* __restrict__ nbrCells, const double* __restrict__ x, double* y,
                const unsigned long long* flag, unsigned long long seq, int* err)
{
    if (flag)
    {
        if (threadIdx.x == 0)
        {
            const volatile unsigned long long* f = flag;
            long long tries = 0;
            while (*f < seq)
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
    }
    const long long t = (long long)blockIdx.x * kBlkThreads + threadIdx.x;
    const long long row = t >> 2;
    const int i = (int)(t & 3);
    if (row >= nRows) return;
    const int c = rowCell[row];
    double acc = y[4ll * c + i];
    for (int k = rowStart[row]; k < rowStart[row + 1]; k++)
    {
        const int f = rowFace[k];
        double xv[4];
        if (nbr)
        {
            ld_cg2(nbr + 4ll * f, xv[0], xv[1]);
            ld_cg2(nbr + 4ll * f + 2, xv[2], xv[3]);
        }
        else
            blk_load_x(x, nbrCells[f], xv);
        acc -= blk_mult_row(kind, coef + (size_t)f * kind, false, xv, i);
    }
    y[4ll * c + i] = acc;
}

} // namespace b200

// ================================================================================================ host side
struct b200_blk
{
    b200_ctx* ctx = nullptr;
    int n = 0, nf = 0;
    DevBuf<int> l, u, losort, losortStart, ownerStart, rowsF, rowsB;
    DevBuf<int4> metaF, metaB;   // per sweep position {row (-1 pad), number of terms, first term, 0}
    DevBuf<int2> termsF, termsB; // {face, neighbour row} in the order the sweep subtracts them
    DevBuf<int2> nb12F, nb12B;   // neighbour rows of terms 1 and 2 of every position (term 0: meta.w)
    DevBuf<double> packF, packB, pDposF, pDposB; // BlkPack copies (Cholesky)
    int nPosF = 0, nPosB = 0, nLevelsF = 0, nLevelsB = 0;
    int dK = 0, uK = 0, lK = 0;
    DevBuf<double> diag, upper, lower;
    int precond = -1, pK = 0;
    DevBuf<double> pD, pRaw;
    // vectors [n][4]
    DevBuf<double> x, b, xSaved, r, p, v, s, t, ph, sh, rw, tmp, tmp2;
    DevBuf<double> partial, red;
    double* hostRed = nullptr; // pinned
    int* hostErr = nullptr;    // pinned copy of devErr, refreshed with every reduction
    DevBuf<unsigned> ticket;
    unsigned ticketBase = 0;
    DevBuf<int> devErr;
    bool haveCoeffs = false, haveVectors = false;
    // coupled patches (processor patches of a decomposed block matrix; a pair on the same rank is served locally)
    struct Iface
    {
        int nFaces = 0, peerRank = 0, peerIface = -1, nRows = 0;
        DevBuf<int> cells, rowCell, rowStart, rowFace;
        int kind = 0;
        DevBuf<double> upper; // coupleUpper
        int64_t recvOff = 0;  // of the patch in the halo buffer, in doubles
    };
    std::deque<Iface> ifaces;
    HaloLink halo;
    bool ifacesFinal = false;
    double nGlobal = 0; // cells of all ranks
    bool profiling = false;
    double clsMs[5] = {0};
    int64_t clsLaunches[5] = {0};
    struct Ev
    {
        int cls;
        cudaEvent_t a, b;
    };
    std::vector<Ev> evRecs;
    std::vector<cudaEvent_t> evPool;
    cudaEvent_t evA = nullptr, evB = nullptr;
};

namespace
{
#define BRC0(call)         \
    do                     \
    {                      \
        int rc0_ = (call); \
        if (rc0_) return rc0_; \
    } while (0)

struct BlkScope
{
    b200_blk* s;
    int cls;
    cudaEvent_t a = nullptr, b = nullptr;
    BlkScope(b200_blk* s_, int cls_) : s(s_), cls(cls_)
    {
        s->ctx->launches++;
        s->clsLaunches[cls]++;
        if (s->profiling)
        {
            auto get = [&]() {
                cudaEvent_t e;
                if (!s->evPool.empty())
                {
                    e = s->evPool.back();
                    s->evPool.pop_back();
                }
                else
                    cudaEventCreate(&e);
                return e;
            };
            a = get();
            b = get();
            cudaEventRecord(a, s->ctx->stream);
        }
    }
    ~BlkScope()
    {
        if (s->profiling)
        {
            cudaEventRecord(b, s->ctx->stream);
            s->evRecs.push_back({cls, a, b});
        }
    }
};

void blk_harvest(b200_blk* s)
{
    for (auto& r : s->evRecs)
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) s->clsMs[r.cls] += ms;
        s->evPool.push_back(r.a);
        s->evPool.push_back(r.b);
    }
    s->evRecs.clear();
}

BlkDev blk_dev(const b200_blk* s)
{
    BlkDev M;
    M.n = s->n;
    M.nf = s->nf;
    M.l = s->l.p;
    M.u = s->u.p;
    M.losort = s->losort.p;
    M.losortStart = s->losortStart.p;
    M.ownerStart = s->ownerStart.p;
    M.dK = s->dK;
    M.uK = s->uK;
    M.lK = s->lK;
    M.diag = s->diag.p;
    M.upper = s->upper.p;
    M.lower = s->lower.p;
    return M;
}

inline int blk_grid(long long n, int threads, int cap = 8 * 148 * 4)
{
    long long g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// level-ordered row list, every level padded to a whole warp (8 rows)
void blk_level_rows(int n, const std::vector<int>& lev, int nLev, std::vector<int>& rows)
{
    std::vector<long long> start((size_t)nLev + 1, 0);
    for (int c = 0; c < n; c++) start[(size_t)lev[c] + 1]++;
    for (int k = 0; k < nLev; k++) start[(size_t)k + 1] = start[k] + (start[(size_t)k + 1] + 7) / 8 * 8;
    rows.assign((size_t)start[nLev], -1);
    std::vector<long long> fill(start.begin(), start.end() - 1);
    for (int c = 0; c < n; c++) rows[(size_t)fill[lev[c]]++] = c;
}

int blk_check_err(b200_blk* s, const char* what)
{
    int e = 0;
    b200_ctx* ctx = s->ctx;
    CK(ctx, cudaMemcpyAsync(&e, s->devErr.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (e)
    {
        cudaMemsetAsync(s->devErr.p, 0, sizeof(int), ctx->stream);
        return set_err(ctx, B200_EDEVICE, e == 2 ? "%s: timed out waiting for a neighbour rank (halo or all-reduce)" : "%s: a block sweep timed out waiting for a dependency", what);
    }
    return B200_OK;
}

// Once per system, at its first operation: pair the coupled patches (same rank) or map the neighbours' receive buffers
// (other ranks; collective), and count the cells of all ranks for BlockIterativeSolver::normFactor's average.
int blk_finalize_ifaces(b200_blk* s)
{
    if (s->ifacesFinal) return B200_OK;
    b200_ctx* ctx = s->ctx;
    const int nI = (int)s->ifaces.size();
    s->nGlobal = (double)s->n;
    for (int i = 0; i < nI; i++)
    {
        const b200_blk::Iface& I = s->ifaces[i];
        if (I.peerRank != ctx->rank) continue;
        if (I.peerIface < 0 || I.peerIface >= nI || I.peerIface == i || s->ifaces[I.peerIface].peerRank != ctx->rank ||
            s->ifaces[I.peerIface].peerIface != i || s->ifaces[I.peerIface].nFaces != I.nFaces)
            return set_err(ctx, B200_EINVAL, "b200_blk: interface %d names interface %d of this rank as its neighbour, which does not match it", i,
                           I.peerIface);
    }
    if (ctx->nranks > 1)
    {
        if (!ctx->peer.enabled)
            return set_err(ctx, B200_EUNSUPPORTED,
                           "b200_blk: block systems on several ranks use the peer-to-peer transport only (B200_TRANSPORT=p2p or auto with peer "
                           "access between the GPUs)");
        PeerLink& L = ctx->peer;
        CK(ctx, cudaMemcpyAsync(s->red.p, &s->nGlobal, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        const PeerAR ar{L.dPeerMbox, L.mbox, ctx->rank, ctx->nranks, ++L.arSeq};
        k_peer_allreduce<<<1, 64, 0, ctx->stream>>>(s->red.p, 1, ar, s->devErr.p);
        CK(ctx, cudaGetLastError());
        CK(ctx, cudaMemcpyAsync(&s->nGlobal, s->red.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        const int k = L.sysCount++;
        bool remote = false;
        std::vector<int> peers(nI), match(nI);
        std::vector<int64_t> sendCount(nI), recvOff((size_t)nI + 1, 0);
        for (int i = 0; i < nI; i++)
        {
            b200_blk::Iface& I = s->ifaces[i];
            const bool rem = I.peerRank != ctx->rank;
            remote = remote || rem;
            peers[i] = I.peerRank;
            match[i] = I.peerIface;
            sendCount[i] = rem ? 4ll * I.nFaces : 0;
            I.recvOff = recvOff[i];
            recvOff[(size_t)i + 1] = recvOff[i] + sendCount[i];
        }
        if (remote) BRC0(setup_halo_link_generic(ctx, k, peers, sendCount, recvOff, match, s->halo));
    }
    else
        for (const auto& I : s->ifaces)
            if (I.peerRank != ctx->rank) return set_err(ctx, B200_EINVAL, "b200_blk: interface to rank %d in a context of one rank", I.peerRank);
    s->ifacesFinal = true;
    return B200_OK;
}

// BlockLduMatrix<Type>::Amul: initInterfaces, AmulCore, updateInterfaces
int blk_amul_dev(b200_blk* s, const double* x, double* y)
{
    b200_ctx* ctx = s->ctx;
    BRC0(blk_finalize_ifaces(s));
    const HaloLink& H = s->halo;
    unsigned long long seq = 0;
    int par = 0;
    if (H.enabled)
    {
        seq = ++s->halo.seq;
        par = (int)(seq & 1ull);
        for (size_t p = 0; p < s->ifaces.size(); p++)
        {
            const b200_blk::Iface& I = s->ifaces[p];
            if (I.peerRank == ctx->rank) continue;
            BlkScope k(s, 0);
            k_blk_halo_push<<<blk_grid(4ll * I.nFaces, 256, 1 << 30), 256, 0, ctx->stream>>>(
                I.nFaces, I.cells.p, x, H.peerData[p] + (size_t)par * H.peerTotal[p], H.peerFlag[p] + (size_t)par * H.peerNPeers[p], seq,
                H.counters + p);
        }
    }
    if (s->n)
    {
        BlkScope k(s, 0);
        k_blk_amul<<<(unsigned)((4ll * s->n + kBlkThreads - 1) / kBlkThreads), kBlkThreads, 0, ctx->stream>>>(blk_dev(s), x, y);
    }
    for (size_t p = 0; p < s->ifaces.size(); p++)
    {
        const b200_blk::Iface& I = s->ifaces[p];
        const bool rem = I.peerRank != ctx->rank;
        if (!I.kind) return set_err(ctx, B200_ESTATE, "b200_blk: coupling coefficients of interface %zu not set", p);
        if (I.nRows == 0 && !rem) continue;
        BlkScope k(s, 0);
        k_blk_iface<<<blk_grid(4ll * I.nRows, kBlkThreads, 1 << 30), kBlkThreads, 0, ctx->stream>>>(
            I.nRows, I.rowCell.p, I.rowStart.p, I.rowFace.p, I.kind, I.upper.p, rem ? H.data(par) + I.recvOff : nullptr,
            rem ? nullptr : s->ifaces[I.peerIface].cells.p, x, y, rem ? H.flags(par, (int)s->ifaces.size()) + p : nullptr, seq, s->devErr.p);
    }
    CK(ctx, cudaGetLastError());
    return B200_OK;
}

int blk_fill_sentinel(b200_blk* s, double* p, long long n)
{
    if (n == 0) return B200_OK;
    BlkScope k(s, 3);
    k_blk_fill_sentinel<<<blk_grid(n, 256), 256, 0, s->ctx->stream>>>(p, n);
    CK(s->ctx, cudaGetLastError());
    return B200_OK;
}

// (re)build the preconditioner diagonal for the current coefficients
int blk_precond_setup(b200_blk* s, int precond)
{
    b200_ctx* ctx = s->ctx;
    if (!s->haveCoeffs) return set_err(ctx, B200_ESTATE, "b200_blk: coefficients not set");
    if (precond != B200_PRECOND_NONE && precond != B200_PRECOND_DIAGONAL && precond != B200_PRECOND_CHOLESKY)
        return set_err(ctx, B200_EINVAL, "b200_blk: unknown block preconditioner %d", precond);
    if (s->precond == precond) return B200_OK;
    s->precond = -1;
    if (precond == B200_PRECOND_NONE)
    {
        s->precond = precond;
        return B200_OK;
    }
    const int chol = precond == B200_PRECOND_CHOLESKY;
    int pK = s->dK;
    if (chol && s->uK > pK) pK = s->uK;
    s->pK = pK;
    CK(ctx, s->pD.alloc((size_t)s->n * pK));
    if (s->n)
    {
        const unsigned ctas = (unsigned)((s->nPosF + kBlkRowsPerCta - 1) / kBlkRowsPerCta);
        if (pK == 16)
        {
            int rc = blk_fill_sentinel(s, s->pD.p, 16ll * s->n);
            if (rc) return rc;
            BlkScope k(s, 4);
            k_blk_diag16<<<ctas, kBlkThreads, 0, ctx->stream>>>(blk_dev(s), chol, s->rowsF.p, s->nPosF, s->pD.p, s->ticket.p, s->ticketBase,
                                                                s->devErr.p);
        }
        else
        {
            CK(ctx, s->pRaw.alloc((size_t)s->n * pK));
            int rc = blk_fill_sentinel(s, s->pRaw.p, (long long)pK * s->n);
            if (rc) return rc;
            BlkScope k(s, 4);
            k_blk_diag_cmpt<<<ctas, kBlkThreads, 0, ctx->stream>>>(blk_dev(s), chol, pK, s->rowsF.p, s->nPosF, s->pRaw.p, s->pD.p, s->ticket.p,
                                                                   s->ticketBase, s->devErr.p);
        }
        s->ticketBase += ctas;
        CK(ctx, cudaGetLastError());
        int rc = blk_check_err(s, "calcPreconDiag");
        if (rc) return rc;
        if (chol)
        { // position-ordered copies for the sweeps (BlkPack)
            const long long nF = (long long)s->nPosF * kBlkDirect * s->uK, nB = (long long)s->nPosB * kBlkDirect * s->uK;
            CK(ctx, s->packF.alloc((size_t)nF));
            CK(ctx, s->packB.alloc((size_t)nB));
            CK(ctx, s->pDposF.alloc((size_t)s->nPosF * pK));
            CK(ctx, s->pDposB.alloc((size_t)s->nPosB * pK));
            BlkScope k(s, 4);
            k_blk_pack_coef<<<blk_grid(nF, 256, 1 << 30), 256, 0, ctx->stream>>>(s->nPosF, s->metaF.p, s->termsF.p, s->lK ? s->lower.p : s->upper.p, s->uK,
                                                                               s->lK ? 0 : 1, s->packF.p);
            k_blk_pack_coef<<<blk_grid(nB, 256, 1 << 30), 256, 0, ctx->stream>>>(s->nPosB, s->metaB.p, s->termsB.p, s->upper.p, s->uK, 0, s->packB.p);
            k_blk_pack_pd<<<blk_grid((long long)s->nPosF * pK, 256, 1 << 30), 256, 0, ctx->stream>>>(s->nPosF, s->metaF.p, s->pD.p, pK, s->pDposF.p);
            k_blk_pack_pd<<<blk_grid((long long)s->nPosB * pK, 256, 1 << 30), 256, 0, ctx->stream>>>(s->nPosB, s->metaB.p, s->pD.p, pK, s->pDposB.p);
            CK(ctx, cudaGetLastError());
        }
    }
    s->precond = precond;
    return B200_OK;
}

// grid of a sweep over `chunks` tickets, and the tickets it takes (persistent CTAs take one more each: the one past the end)
inline unsigned blk_sweep_grid(unsigned chunks)
{
#if B200_BLK_PERSIST
    const unsigned resident = 148u * (unsigned)(B200_BLK_MINCTAS * 256 / kBlkSweepThreads);
    return chunks < resident ? (chunks ? chunks : 1u) : resident;
#else
    return chunks;
#endif
}
inline unsigned blk_sweep_tickets(unsigned chunks, unsigned ctas)
{
#if B200_BLK_PERSIST
    return chunks + ctas;
#else
    (void)chunks;
    return ctas;
#endif
}

// w = M^-1 r (device pointers; r != w)
int blk_precondition_dev(b200_blk* s, const double* r, double* w)
{
    b200_ctx* ctx = s->ctx;
    const long long n4 = 4ll * s->n;
    if (n4 == 0) return B200_OK;
    if (s->precond == B200_PRECOND_NONE)
    {
        CK(ctx, cudaMemcpyAsync(w, r, sizeof(double) * n4, cudaMemcpyDeviceToDevice, ctx->stream));
        return B200_OK;
    }
    const bool chol = s->precond == B200_PRECOND_CHOLESKY;
    static const bool packOff = getenv("B200_BLK_NO_PACK") != nullptr; // developer knob: sweep through the term table only
    const bool usePack = chol && !packOff;
    // forward (Cholesky) or plain diagonal scaling: the same kernel, without faces when !chol
    double* fwdOut = chol ? s->tmp2.p : w;
    int rc = B200_OK;
    if (chol)
    {
        rc = blk_fill_sentinel(s, fwdOut, n4);
        if (rc) return rc;
    }
    BlkDev M = blk_dev(s);
    {
        const unsigned chunks = (unsigned)((s->nPosF + kBlkSweepRows - 1) / kBlkSweepRows);
        const unsigned ctas = blk_sweep_grid(chunks);
        BlkScope k(s, 1);
        k_blk_sweep<false><<<ctas, kBlkSweepThreads, 0, ctx->stream>>>(M, s->pK, chol ? 1 : 0, s->pD.p, s->metaF.p, s->termsF.p,
                                                                   usePack ? BlkPack{s->packF.p, s->pDposF.p, s->nb12F.p} : BlkPack{nullptr, nullptr, nullptr}, s->nPosF, r, fwdOut, s->ticket.p,
                                                                   s->ticketBase, s->devErr.p);
        s->ticketBase += blk_sweep_tickets(chunks, ctas);
        CK(ctx, cudaGetLastError());
    }
    if (chol)
    {
        rc = blk_fill_sentinel(s, w, n4);
        if (rc) return rc;
        const unsigned chunks = (unsigned)((s->nPosB + kBlkSweepRows - 1) / kBlkSweepRows);
        const unsigned ctas = blk_sweep_grid(chunks);
        BlkScope k(s, 2);
        k_blk_sweep<true><<<ctas, kBlkSweepThreads, 0, ctx->stream>>>(M, s->pK, 1, s->pD.p, s->metaB.p, s->termsB.p,
                                                                  usePack ? BlkPack{s->packB.p, s->pDposB.p, s->nb12B.p} : BlkPack{nullptr, nullptr, nullptr}, s->nPosB, fwdOut, w, s->ticket.p, s->ticketBase,
                                                                  s->devErr.p);
        s->ticketBase += blk_sweep_tickets(chunks, ctas);
        CK(ctx, cudaGetLastError());
    }
    return B200_OK;
}

template <int OP>
int blk_reduce_dev(b200_blk* s, const double* a, const double* b, const double* c, int nv, double* hostOut)
{
    b200_ctx* ctx = s->ctx;
    BRC0(blk_finalize_ifaces(s));
    const int blocks = blk_grid(s->n, 256, kBlkRedBlocks);
    {
        BlkScope k(s, 3);
        k_blk_reduce<OP><<<blocks, 256, 0, ctx->stream>>>(s->n, a, b, c, s->partial.p);
    }
    {
        BlkScope k(s, 3);
        PeerAR ar{nullptr, nullptr, ctx->rank, 1, 0};
        if (ctx->nranks > 1) ar = PeerAR{ctx->peer.dPeerMbox, ctx->peer.mbox, ctx->rank, ctx->nranks, ++ctx->peer.arSeq};
        k_blk_reduce_final<<<1, 256, 0, ctx->stream>>>(s->partial.p, blocks, nv, s->red.p, ar, s->devErr.p);
    }
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(s->hostRed, s->red.p, sizeof(double) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaMemcpyAsync(s->hostErr, s->devErr.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int v = 0; v < nv; v++) hostOut[v] = s->hostRed[v];
    if (*s->hostErr) return blk_check_err(s, "block solve"); // a sweep gave up waiting: stop instead of iterating on garbage
    return B200_OK;
}

template <int OP>
int blk_vec(b200_blk* s, double* y, double* z, const double* a, const double* b, const double* c, const double* d, double s1, double s2)
{
    const long long n4 = 4ll * s->n;
    if (n4 == 0) return B200_OK;
    BlkScope k(s, 3);
    k_blk_vec<OP><<<blk_grid(n4, 256), 256, 0, s->ctx->stream>>>(n4, y, z, a, b, c, d, s1, s2);
    CK(s->ctx, cudaGetLastError());
    return B200_OK;
}

int blk_alloc_vectors(b200_blk* s)
{
    if (s->haveVectors) return B200_OK;
    b200_ctx* ctx = s->ctx;
    const size_t n4 = 4 * (size_t)s->n;
    for (DevBuf<double>* v : {&s->x, &s->b, &s->xSaved, &s->r, &s->p, &s->v, &s->s, &s->t, &s->ph, &s->sh, &s->rw, &s->tmp, &s->tmp2})
        CK(ctx, v->alloc(n4));
    s->haveVectors = true;
    return B200_OK;
}

double cmax4(const double* v)
{
    double m = v[0];
    for (int i = 1; i < 4; i++)
        if (v[i] > m) m = v[i];
    return m;
}

// BlockLduSolver::stop + BlockSolverPerformance<Type>::checkConvergence
bool blk_stop(const b200_solver_opts* o, b200_blk_perf* p)
{
    if (p->nIterations < o->minIter) return false;
    const double fin = cmax4(p->finalResidual), ini = cmax4(p->initialResidual);
    p->converged = (fin < o->tolerance || (o->relTol > B200_SMALL_ && fin <= o->relTol * ini)) ? 1 : 0;
    return p->nIterations >= o->maxIter || p->converged;
}

#define BRC(call)              \
    do                         \
    {                          \
        int rc_ = (call);      \
        if (rc_) return rc_;   \
    } while (0)

int blk_residual(b200_blk* s, const double* r, double nf, double* out4)
{
    BRC(blk_reduce_dev<BR_CMPTMAG>(s, r, nullptr, nullptr, 4, out4));
    for (int i = 0; i < 4; i++) out4[i] /= nf;
    return B200_OK;
}

// BlockIterativeSolver<Type>::normFactor
int blk_norm_factor(b200_blk* s, double* nfOut)
{
    double sum4[4];
    BRC(blk_reduce_dev<BR_SUM4>(s, s->x.p, nullptr, nullptr, 4, sum4));
    for (int i = 0; i < 4; i++) sum4[i] /= s->nGlobal > 0 ? s->nGlobal : 1.0; // gAverage
    if (s->n)
    {
        BlkScope k(s, 3);
        k_blk_fill4<<<blk_grid(s->n, 256), 256, 0, s->ctx->stream>>>(s->tmp.p, s->n, sum4[0], sum4[1], sum4[2], sum4[3]);
    }
    BRC(blk_amul_dev(s, s->x.p, s->v.p));   // wA = A x
    BRC(blk_amul_dev(s, s->tmp.p, s->t.p)); // pA = A xRef
    double nt = 0.0;
    BRC(blk_reduce_dev<BR_NORMTERMS>(s, s->v.p, s->t.p, s->b.p, 1, &nt));
    *nfOut = nt + B200_SMALL;
    return B200_OK;
}

int blk_solve_core(b200_blk* s, const b200_solver_opts* o, b200_blk_perf* perf, double* history, int cap)
{
    b200_ctx* ctx = s->ctx;
    const size_t nb = sizeof(double) * 4 * (size_t)s->n;
    auto hist = [&](int it, const double* res) {
        if (history && it < cap)
            for (int i = 0; i < 4; i++) history[4 * it + i] = res[i];
    };
    double nf;
    BRC(blk_norm_factor(s, &nf));
    perf->normFactor = nf;
    BRC(blk_amul_dev(s, s->x.p, s->p.p));
    BRC((blk_vec<BV_SUB>(s, s->r.p, nullptr, s->b.p, s->p.p, nullptr, nullptr, 0, 0)));
    BRC(blk_residual(s, s->r.p, nf, perf->initialResidual));
    memcpy(perf->finalResidual, perf->initialResidual, sizeof(double) * 4);
    hist(0, perf->initialResidual);
    if (blk_stop(o, perf)) return B200_OK;
    BRC(blk_precond_setup(s, o->precond));
    if (o->solver == B200_BLK_SOLVER_BICGSTAB)
    {
        double rho = B200_GREAT, rhoOld = rho, alpha = 0, omega = B200_GREAT, beta;
        for (DevBuf<double>* v : {&s->p, &s->ph, &s->v, &s->s, &s->sh, &s->t}) CK(ctx, cudaMemsetAsync(v->p, 0, nb ? nb : 1, ctx->stream));
        CK(ctx, cudaMemcpyAsync(s->rw.p, s->r.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        do
        {
            rhoOld = rho;
            BRC(blk_reduce_dev<BR_PROD>(s, s->rw.p, s->r.p, nullptr, 1, &rho));
            beta = rho / rhoOld * (alpha / omega);
            if (rho == 0)
            { // restart if breakdown occurs
                CK(ctx, cudaMemcpyAsync(s->rw.p, s->r.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
                BRC(blk_reduce_dev<BR_PROD>(s, s->rw.p, s->r.p, nullptr, 1, &rho));
                alpha = 0;
                omega = 0;
                beta = 0;
            }
            BRC((blk_vec<BV_P>(s, s->p.p, nullptr, s->r.p, s->v.p, nullptr, nullptr, beta, omega)));
            BRC(blk_precondition_dev(s, s->p.p, s->ph.p));
            BRC(blk_amul_dev(s, s->ph.p, s->v.p));
            double rwv;
            BRC(blk_reduce_dev<BR_PROD>(s, s->rw.p, s->v.p, nullptr, 1, &rwv));
            alpha = rho / rwv;
            BRC((blk_vec<BV_S>(s, s->s.p, nullptr, s->r.p, s->v.p, nullptr, nullptr, alpha, 0)));
            BRC(blk_precondition_dev(s, s->s.p, s->sh.p));
            BRC(blk_amul_dev(s, s->sh.p, s->t.p));
            double tstt[2];
            BRC(blk_reduce_dev<BR_PROD2>(s, s->t.p, s->s.p, nullptr, 2, tstt));
            omega = tstt[0] / tstt[1];
            BRC((blk_vec<BV_XR>(s, s->x.p, s->r.p, s->ph.p, s->sh.p, s->s.p, s->t.p, alpha, omega)));
            BRC(blk_residual(s, s->r.p, nf, perf->finalResidual));
            perf->nIterations++;
            hist(perf->nIterations, perf->finalResidual);
        } while (!blk_stop(o, perf));
    }
    else
    { // BlockCGSolver: r = rA, t = wA, p = pA
        double rho = B200_GREAT, rhoOld = rho;
        CK(ctx, cudaMemsetAsync(s->p.p, 0, nb ? nb : 1, ctx->stream));
        do
        {
            rhoOld = rho;
            BRC(blk_precondition_dev(s, s->r.p, s->t.p));
            BRC(blk_reduce_dev<BR_PROD>(s, s->t.p, s->r.p, nullptr, 1, &rho));
            const double beta = rho / rhoOld;
            BRC((blk_vec<BV_CG_P>(s, s->p.p, nullptr, s->t.p, nullptr, nullptr, nullptr, beta, 0)));
            BRC(blk_amul_dev(s, s->p.p, s->t.p));
            double wApA;
            BRC(blk_reduce_dev<BR_PROD>(s, s->t.p, s->p.p, nullptr, 1, &wApA));
            if (!(fabs(wApA) / nf > B200_VSMALL))
            {
                perf->singular = 1;
                break;
            }
            const double alpha = rho / wApA;
            BRC((blk_vec<BV_CG_XR>(s, s->x.p, s->r.p, s->p.p, s->t.p, nullptr, nullptr, alpha, 0)));
            BRC(blk_residual(s, s->r.p, nf, perf->finalResidual));
            perf->nIterations++;
            hist(perf->nIterations, perf->finalResidual);
        } while (!blk_stop(o, perf));
    }
    return blk_check_err(s, "b200_blk_solve");
}
} // namespace

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int b200_blk_create(b200_ctx* ctx, int32_t nCells, int32_t nFaces, const int32_t* lowerAddr, const int32_t* upperAddr, b200_blk** out)
{
    if (!ctx || !out || nCells < 0 || nFaces < 0 || (nFaces > 0 && (!lowerAddr || !upperAddr)))
        return set_err(ctx, B200_EINVAL, "b200_blk_create: bad arguments");
    // lduAddressing: l < u, faces in upper-triangular order
    for (int f = 0; f < nFaces; f++)
    {
        const int l = lowerAddr[f], u = upperAddr[f];
        if (l < 0 || u >= nCells || l >= u) return set_err(ctx, B200_EINVAL, "b200_blk_create: face %d (%d, %d) is not lower < upper < nCells", f, l, u);
        if (f > 0 && (l < lowerAddr[f - 1] || (l == lowerAddr[f - 1] && u < upperAddr[f - 1])))
            return set_err(ctx, B200_EINVAL, "b200_blk_create: faces are not in upper-triangular order at face %d", f);
    }
    CK(ctx, cudaSetDevice(ctx->device));
    std::unique_ptr<b200_blk> s(new b200_blk);
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
    // wavefront levels of the forward (lower neighbours) and backward (upper neighbours) sweeps
    std::vector<int> levF((size_t)nCells, 0), levB((size_t)nCells, 0), rowsF, rowsB;
    int nLevF = nCells ? 1 : 0, nLevB = nCells ? 1 : 0;
    for (int f = 0; f < nFaces; f++) // owner-sorted faces: levF[l] is final when the owner's faces are reached
    {
        levF[u[f]] = std::max(levF[u[f]], levF[l[f]] + 1);
        nLevF = std::max(nLevF, levF[u[f]] + 1);
    }
    for (int f = nFaces - 1; f >= 0; f--)
    {
        levB[l[f]] = std::max(levB[l[f]], levB[u[f]] + 1);
        nLevB = std::max(nLevB, levB[l[f]] + 1);
    }
    blk_level_rows(nCells, levF, nLevF, rowsF);
    blk_level_rows(nCells, levB, nLevB, rowsB);
    s->nPosF = (int)rowsF.size();
    s->nPosB = (int)rowsB.size();
    // sweep-ordered term lists: forward = lower faces of the row ascending (losort order), backward = owner faces DESCENDING
    std::vector<int4> metaF(rowsF.size()), metaB(rowsB.size());
    std::vector<int2> termsF((size_t)nFaces), termsB((size_t)nFaces);
    std::vector<int2> nb12F(rowsF.size(), make_int2(0, 0)), nb12B(rowsB.size(), make_int2(0, 0));
    auto firstThree = [](const std::vector<int2>& terms, int4& mt, int2& nb12) {
        mt.w = mt.y > 0 ? terms[(size_t)mt.z].y : 0;
        nb12 = make_int2(mt.y > 1 ? terms[(size_t)mt.z + 1].y : 0, mt.y > 2 ? terms[(size_t)mt.z + 2].y : 0);
    };
    {
        int kF = 0, kB = 0;
        for (size_t pos = 0; pos < rowsF.size(); pos++)
        {
            const int row = rowsF[pos];
            if (row < 0)
            {
                metaF[pos] = make_int4(-1, 0, 0, 0);
                continue;
            }
            metaF[pos] = make_int4(row, losortStart[(size_t)row + 1] - losortStart[row], kF, 0);
            for (int k = losortStart[row]; k < losortStart[(size_t)row + 1]; k++) termsF[(size_t)kF++] = make_int2(losort[k], l[losort[k]]);
            firstThree(termsF, metaF[pos], nb12F[pos]);
        }
        for (size_t pos = 0; pos < rowsB.size(); pos++)
        {
            const int row = rowsB[pos];
            if (row < 0)
            {
                metaB[pos] = make_int4(-1, 0, 0, 0);
                continue;
            }
            metaB[pos] = make_int4(row, ownerStart[(size_t)row + 1] - ownerStart[row], kB, 0);
            for (int f = ownerStart[(size_t)row + 1] - 1; f >= ownerStart[row]; f--) termsB[(size_t)kB++] = make_int2(f, u[f]);
            firstThree(termsB, metaB[pos], nb12B[pos]);
        }
    }
    s->nLevelsF = nLevF;
    s->nLevelsB = nLevB;
    cudaStream_t st = ctx->stream;
    CK(ctx, s->l.upload(l, st));
    CK(ctx, s->u.upload(u, st));
    CK(ctx, s->losort.upload(losort, st));
    CK(ctx, s->losortStart.upload(losortStart, st));
    CK(ctx, s->ownerStart.upload(ownerStart, st));
    CK(ctx, s->rowsF.upload(rowsF, st));
    CK(ctx, s->rowsB.upload(rowsB, st));
    CK(ctx, s->metaF.upload(metaF, st));
    CK(ctx, s->metaB.upload(metaB, st));
    CK(ctx, s->termsF.upload(termsF, st));
    CK(ctx, s->termsB.upload(termsB, st));
    CK(ctx, s->nb12F.upload(nb12F, st));
    CK(ctx, s->nb12B.upload(nb12B, st));
    CK(ctx, s->partial.alloc((size_t)kBlkMaxRed * kBlkRedBlocks));
    CK(ctx, s->red.alloc(8));
    CK(ctx, s->ticket.alloc(1));
    CK(ctx, s->devErr.alloc(1));
    CK(ctx, cudaMemsetAsync(s->ticket.p, 0, sizeof(unsigned), st));
    CK(ctx, cudaMemsetAsync(s->devErr.p, 0, sizeof(int), st));
    CK(ctx, cudaMallocHost((void**)&s->hostRed, 8 * sizeof(double)));
    CK(ctx, cudaMallocHost((void**)&s->hostErr, sizeof(int)));
    *s->hostErr = 0;
    CK(ctx, cudaEventCreate(&s->evA));
    CK(ctx, cudaEventCreate(&s->evB));
    CK(ctx, cudaStreamSynchronize(st));
    *out = s.release();
    return B200_OK;
}

extern "C" int b200_blk_destroy(b200_blk* s)
{
    if (!s) return B200_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    blk_harvest(s);
    for (auto e : s->evPool) cudaEventDestroy(e);
    if (s->evA) cudaEventDestroy(s->evA);
    if (s->evB) cudaEventDestroy(s->evB);
    if (s->hostRed) cudaFreeHost(s->hostRed);
    if (s->hostErr) cudaFreeHost(s->hostErr);
    s->halo.release();
    delete s;
    return B200_OK;
}

extern "C" int b200_blk_set_coeffs(b200_blk* s, int dK, const double* diag, int uK, const double* upper, int lK, const double* lower)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    auto okKind = [](int k) { return k == 1 || k == 4 || k == 16; };
    if (!okKind(dK) || !okKind(uK) || (s->n && !diag) || (s->nf && !upper)) return set_err(ctx, B200_EINVAL, "b200_blk_set_coeffs: bad kind or null array");
    if (lower && lK != uK)
        return set_err(ctx, B200_EUNSUPPORTED, "b200_blk_set_coeffs: lower and upper must have the same active type (%d vs %d)", lK, uK);
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ctx, s->diag.alloc((size_t)s->n * dK));
    CK(ctx, s->upper.alloc((size_t)s->nf * uK));
    if (s->n) CK(ctx, cudaMemcpyAsync(s->diag.p, diag, sizeof(double) * (size_t)s->n * dK, cudaMemcpyHostToDevice, st));
    if (s->nf) CK(ctx, cudaMemcpyAsync(s->upper.p, upper, sizeof(double) * (size_t)s->nf * uK, cudaMemcpyHostToDevice, st));
    if (lower)
    {
        CK(ctx, s->lower.alloc((size_t)s->nf * lK));
        if (s->nf) CK(ctx, cudaMemcpyAsync(s->lower.p, lower, sizeof(double) * (size_t)s->nf * lK, cudaMemcpyHostToDevice, st));
    }
    else
        s->lower.release();
    CK(ctx, cudaStreamSynchronize(st));
    s->dK = dK;
    s->uK = uK;
    s->lK = lower ? lK : 0;
    s->haveCoeffs = true;
    s->precond = -1;
    return B200_OK;
}

// BlockLduInterfaceField / processorFvPatchField<vector4>: a coupled patch of the block matrix
extern "C" int b200_blk_add_interface(b200_blk* s, int32_t nFaces, const int32_t* faceCells, int32_t peerRank, int32_t peerIface, int32_t* index)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (s->ifacesFinal) return set_err(ctx, B200_ESTATE, "b200_blk_add_interface: the system has been used already");
    if (nFaces < 0 || (nFaces && !faceCells) || peerRank < 0 || peerRank >= ctx->nranks || peerIface < 0)
        return set_err(ctx, B200_EINVAL, "b200_blk_add_interface: bad argument (nFaces %d, peer rank %d, peer interface %d)", nFaces, peerRank, peerIface);
    for (int f = 0; f < nFaces; f++)
        if (faceCells[f] < 0 || faceCells[f] >= s->n) return set_err(ctx, B200_EINVAL, "b200_blk_add_interface: faceCells[%d] = %d out of range", f, faceCells[f]);
    CK(ctx, cudaSetDevice(ctx->device));
    // rows of the patch: its cells, each with its faces in patch order
    std::vector<int> order(nFaces), cells(faceCells, faceCells + nFaces), rowCell, rowStart;
    for (int f = 0; f < nFaces; f++) order[f] = f;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return faceCells[a] < faceCells[b]; });
    for (int k = 0; k < nFaces; k++)
        if (k == 0 || faceCells[order[k]] != faceCells[order[k - 1]])
        {
            rowCell.push_back(faceCells[order[k]]);
            rowStart.push_back(k);
        }
    rowStart.push_back(nFaces);
    s->ifaces.emplace_back();
    b200_blk::Iface& I = s->ifaces.back();
    I.nFaces = nFaces;
    I.peerRank = peerRank;
    I.peerIface = peerIface;
    I.nRows = (int)rowCell.size();
    cudaStream_t st = ctx->stream;
    CK(ctx, I.cells.upload(cells, st));
    CK(ctx, I.rowCell.upload(rowCell, st));
    CK(ctx, I.rowStart.upload(rowStart, st));
    CK(ctx, I.rowFace.upload(order, st));
    CK(ctx, cudaStreamSynchronize(st));
    if (index) *index = (int32_t)s->ifaces.size() - 1;
    return B200_OK;
}

// BlockLduMatrix<vector4>::coupleUpper()[patch] (coupleLower is used by Tmul only, which the solvers of this path never call)
extern "C" int b200_blk_set_interface_coeffs(b200_blk* s, int32_t iface, int kind, const double* coupleUpper)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (iface < 0 || iface >= (int)s->ifaces.size()) return set_err(ctx, B200_EINVAL, "b200_blk_set_interface_coeffs: no interface %d", iface);
    b200_blk::Iface& I = s->ifaces[iface];
    if ((kind != 1 && kind != 4 && kind != 16) || (I.nFaces && !coupleUpper))
        return set_err(ctx, B200_EINVAL, "b200_blk_set_interface_coeffs: bad kind or null array");
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, I.upper.alloc((size_t)I.nFaces * kind));
    if (I.nFaces) CK(ctx, cudaMemcpyAsync(I.upper.p, coupleUpper, sizeof(double) * (size_t)I.nFaces * kind, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    I.kind = kind;
    return B200_OK;
}

extern "C" int b200_blk_amul(b200_blk* s, const double* x, double* y)
{
    if (!s || (s->n && (!x || !y))) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->haveCoeffs) return set_err(ctx, B200_ESTATE, "b200_blk_amul: coefficients not set");
    CK(ctx, cudaSetDevice(ctx->device));
    BRC(blk_alloc_vectors(s));
    const size_t nb = sizeof(double) * 4 * (size_t)s->n;
    if (nb) CK(ctx, cudaMemcpyAsync(s->tmp.p, x, nb, cudaMemcpyHostToDevice, ctx->stream));
    BRC(blk_amul_dev(s, s->tmp.p, s->t.p));
    if (nb) CK(ctx, cudaMemcpyAsync(y, s->t.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_blk_precondition(b200_blk* s, int precond, const double* r, double* w)
{
    if (!s || (s->n && (!r || !w))) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    BRC(blk_alloc_vectors(s));
    BRC(blk_precond_setup(s, precond));
    const size_t nb = sizeof(double) * 4 * (size_t)s->n;
    if (nb) CK(ctx, cudaMemcpyAsync(s->tmp.p, r, nb, cudaMemcpyHostToDevice, ctx->stream));
    BRC(blk_precondition_dev(s, s->tmp.p, s->t.p));
    BRC(blk_check_err(s, "b200_blk_precondition"));
    if (nb) CK(ctx, cudaMemcpyAsync(w, s->t.p, nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_blk_get_precon_diag(b200_blk* s, int precond, double* out, int* kind)
{
    if (!s || !kind || (s->n && !out)) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    if (precond == B200_PRECOND_NONE) return set_err(ctx, B200_EINVAL, "b200_blk_get_precon_diag: no diagonal for BlockNoPrecon");
    BRC(blk_precond_setup(s, precond));
    *kind = s->pK;
    if (s->n) CK(ctx, cudaMemcpyAsync(out, s->pD.p, sizeof(double) * (size_t)s->n * s->pK, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_blk_reduce(b200_blk* s, const double* a, const double* b, double* out5)
{
    if (!s || !out5 || (s->n && (!a || !b))) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    BRC(blk_alloc_vectors(s));
    const size_t nb = sizeof(double) * 4 * (size_t)s->n;
    if (nb) CK(ctx, cudaMemcpyAsync(s->tmp.p, a, nb, cudaMemcpyHostToDevice, ctx->stream));
    if (nb) CK(ctx, cudaMemcpyAsync(s->t.p, b, nb, cudaMemcpyHostToDevice, ctx->stream));
    BRC(blk_reduce_dev<BR_PROD>(s, s->tmp.p, s->t.p, nullptr, 1, out5));
    BRC(blk_reduce_dev<BR_CMPTMAG>(s, s->tmp.p, nullptr, nullptr, 4, out5 + 1));
    return B200_OK;
}

extern "C" int b200_blk_upload(b200_blk* s, const double* x, const double* b)
{
    if (!s || (s->n && (!x || !b))) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    BRC(blk_alloc_vectors(s));
    const size_t nb = sizeof(double) * 4 * (size_t)s->n;
    if (nb) CK(ctx, cudaMemcpyAsync(s->x.p, x, nb, cudaMemcpyHostToDevice, ctx->stream));
    if (nb) CK(ctx, cudaMemcpyAsync(s->b.p, b, nb, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_blk_download(b200_blk* s, double* x)
{
    if (!s || (s->n && !x)) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->haveVectors) return set_err(ctx, B200_ESTATE, "b200_blk_download: nothing uploaded");
    CK(ctx, cudaSetDevice(ctx->device));
    if (s->n) CK(ctx, cudaMemcpyAsync(x, s->x.p, sizeof(double) * 4 * (size_t)s->n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_blk_x_save(b200_blk* s)
{
    if (!s || !s->haveVectors) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (s->n) CK(ctx, cudaMemcpyAsync(s->xSaved.p, s->x.p, sizeof(double) * 4 * (size_t)s->n, cudaMemcpyDeviceToDevice, ctx->stream));
    return B200_OK;
}

extern "C" int b200_blk_x_restore(b200_blk* s)
{
    if (!s || !s->haveVectors) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (s->n) CK(ctx, cudaMemcpyAsync(s->x.p, s->xSaved.p, sizeof(double) * 4 * (size_t)s->n, cudaMemcpyDeviceToDevice, ctx->stream));
    return B200_OK;
}

extern "C" int b200_blk_solve_resident(b200_blk* s, const b200_solver_opts* o, b200_blk_perf* perf, double* history, int cap)
{
    if (!s || !o || !perf) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->haveCoeffs || !s->haveVectors) return set_err(ctx, B200_ESTATE, "b200_blk_solve: coefficients / vectors not set");
    if (o->solver != B200_BLK_SOLVER_CG && o->solver != B200_BLK_SOLVER_BICGSTAB)
        return set_err(ctx, B200_EINVAL, "b200_blk_solve: unknown block solver %d", o->solver);
    if (o->precond != B200_PRECOND_NONE && o->precond != B200_PRECOND_DIAGONAL && o->precond != B200_PRECOND_CHOLESKY)
        return set_err(ctx, B200_EINVAL, "b200_blk_solve: unknown block preconditioner %d", o->precond);
    CK(ctx, cudaSetDevice(ctx->device));
    memset(perf, 0, sizeof(*perf));
    if (history)
        for (int i = 0; i < 4 * cap; i++) history[i] = std::nan("");
    CK(ctx, cudaEventRecord(s->evA, ctx->stream));
    const int rc = blk_solve_core(s, o, perf, history, cap);
    CK(ctx, cudaEventRecord(s->evB, ctx->stream));
    CK(ctx, cudaEventSynchronize(s->evB));
    float ms = 0;
    CK(ctx, cudaEventElapsedTime(&ms, s->evA, s->evB));
    perf->deviceMs = ms;
    return rc;
}

extern "C" int b200_blk_solve(b200_blk* s, const b200_solver_opts* o, double* x, const double* b, b200_blk_perf* perf, double* history, int cap)
{
    BRC(b200_blk_upload(s, x, b));
    BRC(b200_blk_solve_resident(s, o, perf, history, cap));
    return b200_blk_download(s, x);
}

extern "C" int b200_blk_set_profiling(b200_blk* s, int enable)
{
    if (!s) return B200_EINVAL;
    cudaStreamSynchronize(s->ctx->stream);
    blk_harvest(s);
    s->profiling = enable != 0;
    return B200_OK;
}

extern "C" int b200_blk_get_kernel_times(b200_blk* s, double* ms5, int64_t* launches5, int reset)
{
    if (!s) return B200_EINVAL;
    cudaStreamSynchronize(s->ctx->stream);
    blk_harvest(s);
    for (int i = 0; i < 5; i++)
    {
        if (ms5) ms5[i] = s->clsMs[i];
        if (launches5) launches5[i] = s->clsLaunches[i];
        if (reset)
        {
            s->clsMs[i] = 0;
            s->clsLaunches[i] = 0;
        }
    }
    return B200_OK;
}
