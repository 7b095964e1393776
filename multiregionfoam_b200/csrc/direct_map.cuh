// direct_map.cuh -- conformal-interface maps on the device: k_direct_map (first match of an N^2 search, the other zone's
// locations staged through shared memory in ascending tiles so "first" is kept) and the C-ABI entry points
// b200_direct_map_build / b200_direct_map_transfer.  Included by b200_ldu.cu.  Arithmetic: direct_map.hpp.
//
// The reference's search is O(P^2) on one core (C3's interface, P = 146 400: 10^10 distance tests); here every receiving
// location is a thread, a tile of 256 source locations (6 KB) is shared by the block, a block leaves as soon as all of
// its threads have their match.  FP64-compute-bound, not HBM-bound: 2 x 24 B per location are read once per block.
#pragma once

#include "direct_map.hpp"

namespace b200
{

constexpr int kMapTile = 256;

__global__ void __launch_bounds__(kMapTile) k_direct_map(int32_t nTo, const double* __restrict__ to, int32_t nFrom,
                                                         const double* __restrict__ from, double tol, int* __restrict__ map)
{
    __shared__ double sx[kMapTile], sy[kMapTile], sz[kMapTile];
    const double fourTol2 = 4.0 * tol * tol;
    for (int32_t base = blockIdx.x * kMapTile; base < nTo; base += gridDim.x * kMapTile)
    {
        const int32_t i = base + threadIdx.x;
        const bool live = i < nTo;
        double ax = 0, ay = 0, az = 0;
        if (live)
        {
            ax = to[3 * i];
            ay = to[3 * i + 1];
            az = to[3 * i + 2];
        }
        int found = live ? -1 : 0;
        for (int32_t j0 = 0; j0 < nFrom; j0 += kMapTile)
        {
            const int32_t j = j0 + threadIdx.x;
            if (j < nFrom)
            {
                sx[threadIdx.x] = from[3 * j];
                sy[threadIdx.x] = from[3 * j + 1];
                sz[threadIdx.x] = from[3 * j + 2];
            }
            __syncthreads();
            if (found < 0)
            {
                const int n = min(kMapTile, nFrom - j0);
                for (int k = 0; k < n; k++)
                    if (dmap::matches(ax, ay, az, sx[k], sy[k], sz[k], tol, fourTol2))
                    {
                        found = j0 + k; // ascending tiles, ascending k: the first match, as the reference's break
                        break;
                    }
            }
            if (__syncthreads_and(found >= 0)) break;
        }
        if (live) map[i] = found;
        __syncthreads();
    }
}

} // namespace b200

extern "C" int b200_direct_map_build(b200_ctx* ctx, int32_t nTo, const double* toXyz, int32_t nFrom, const double* fromXyz,
                                     double tol, int32_t* map)
{
    if (!ctx || nTo < 0 || nFrom < 0 || (nTo && (!toXyz || !map)) || (nFrom && !fromXyz) || !(tol >= 0))
        return set_err(ctx, B200_EINVAL, "b200_direct_map_build: bad arguments");
    CK(ctx, cudaSetDevice(ctx->device));
    if (nTo == 0) return 0;
    cudaStream_t st = ctx->stream;
    DevBuf<double> dTo, dFrom;
    DevBuf<int> dMap;
    CK(ctx, dTo.alloc(3 * (size_t)nTo));
    CK(ctx, dFrom.alloc(3 * (size_t)nFrom));
    CK(ctx, dMap.alloc(nTo));
    CK(ctx, cudaMemcpyAsync(dTo.p, toXyz, sizeof(double) * 3 * nTo, cudaMemcpyHostToDevice, st));
    if (nFrom) CK(ctx, cudaMemcpyAsync(dFrom.p, fromXyz, sizeof(double) * 3 * nFrom, cudaMemcpyHostToDevice, st));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((nTo + kMapTile - 1) / kMapTile, (int64_t)ctx->smCount * 8));
    ctx->launches++;
    k_direct_map<<<blocks, kMapTile, 0, st>>>(nTo, dTo.p, nFrom, dFrom.p, tol, dMap.p);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(map, dMap.p, sizeof(int) * nTo, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    int unmatched = 0;
    for (int32_t i = 0; i < nTo; i++) unmatched += map[i] < 0;
    return unmatched;
}

extern "C" int b200_direct_map_transfer(b200_ctx* ctx, int32_t nTo, const int32_t* map, int32_t nFrom, const double* from,
                                        int nComp, double* to)
{
    if (!ctx || nTo < 0 || nFrom < 0 || nComp < 1 || (nTo && (!map || !from || !to)))
        return set_err(ctx, B200_EINVAL, "b200_direct_map_transfer: bad arguments");
    // interfaceToInterfaceMapping::checkFieldSizes + a map entry of -1 is the reference's "not conformal" FatalError
    for (int32_t i = 0; i < nTo; i++)
        if (map[i] < 0 || map[i] >= nFrom)
            return set_err(ctx, B200_EINVAL, "b200_direct_map_transfer: map[%d] = %d outside the source zone (%d)", i, map[i], nFrom);
    CK(ctx, cudaSetDevice(ctx->device));
    if (nTo == 0) return B200_OK;
    DevBuf<int> dMap;
    DevBuf<double> dT, dF;
    cudaStream_t st = ctx->stream;
    CK(ctx, dMap.alloc(nTo));
    CK(ctx, dT.alloc((size_t)nTo * nComp));
    CK(ctx, dF.alloc((size_t)nFrom * nComp));
    CK(ctx, cudaMemcpyAsync(dMap.p, map, sizeof(int) * nTo, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(dF.p, from, sizeof(double) * nFrom * nComp, cudaMemcpyHostToDevice, st));
    ctx->launches++;
    k_gather_zone<<<(nTo * nComp + 127) / 128, 128, 0, st>>>(nTo, dMap.p, dF.p, nComp, dT.p);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(to, dT.p, sizeof(double) * nTo * nComp, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return B200_OK;
}
