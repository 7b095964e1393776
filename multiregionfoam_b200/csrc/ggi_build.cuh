// ggi_build.cuh -- GGI weight construction on the device: narrow phase and per-face rescale as
// kernels, C-ABI entry points b200_ggi_build / b200_ggi_fetch.  Included by b200_ldu.cu.  Arithmetic and the host
// broad phase: ggi_build.hpp (shared with the CPU emulator of the tests).
//
// k_ggi_pairs: one thread per candidate pair (a patch of P faces has ~9-25 P candidates): reads the two faces' labels and
// points (<= 8 + 8 points x 24 B, neighbouring threads share them through L1/L2), clips in registers / local memory
// (640 B per thread), writes one area.  Compute-light FP64 (a few hundred flops per pair) on at most a few million pairs:
// launch-latency territory next to the solve; the point of running it here is that the rebuilt tables never leave the
// device's side of the call when the interface moves every interpolatorUpdateFrequency steps.
#pragma once

#include "ggi_build.hpp"

namespace b200
{

__global__ void k_ggi_pairs(int32_t nPairs, const int* __restrict__ pairMaster, const int* __restrict__ cand,
                            const int* __restrict__ mOff, const int* __restrict__ mFp, const double* __restrict__ mPts,
                            const int* __restrict__ sOff, const int* __restrict__ sFp, const double* __restrict__ sPts,
                            double* __restrict__ area, double* __restrict__ mArea)
{
    for (int32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nPairs; k += gridDim.x * blockDim.x)
    {
        const int i = pairMaster[k], j = cand[k];
        double ma;
        area[k] = ggib::pair_area(mPts, mFp + mOff[i], mOff[i + 1] - mOff[i], sPts, sFp + sOff[j], sOff[j + 1] - sOff[j], ma);
        mArea[i] = ma; // every pair of row i writes the same value
    }
}

__global__ void k_ggi_rows(int32_t nM, const int* __restrict__ candOff, const double* __restrict__ area,
                           const double* __restrict__ mArea, double tol, int rescale, double* __restrict__ w)
{
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nM; i += gridDim.x * blockDim.x)
        if (candOff[i + 1] > candOff[i]) ggib::row_weights(area, mArea[i], candOff[i], candOff[i + 1], tol, rescale, w);
}

} // namespace b200

static int ggi_check_patch(b200_ctx* ctx, const char* what, int32_t nF, const int32_t* off, const int32_t* lab, int32_t nP,
                           const double* pts)
{
    if (nF < 0 || nP < 0 || (nF && (!off || !lab || !pts))) return set_err(ctx, B200_EINVAL, "b200_ggi_build: bad %s patch", what);
    for (int32_t f = 0; f < nF; f++)
    {
        const int n = off[f + 1] - off[f];
        if (n < 3 || n > ggib::kMaxV)
            return set_err(ctx, B200_EUNSUPPORTED, "b200_ggi_build: %s face %d has %d points (3..%d supported)", what, f, n, ggib::kMaxV);
        for (int32_t k = off[f]; k < off[f + 1]; k++)
            if (lab[k] < 0 || lab[k] >= nP) return set_err(ctx, B200_EINVAL, "b200_ggi_build: %s face %d: point %d out of range", what, f, lab[k]);
    }
    return B200_OK;
}

extern "C" int b200_ggi_build(b200_ctx* ctx, int32_t nMaster, const int32_t* mFaceOffsets, const int32_t* mFaceLabels,
                              int32_t nMasterPoints, const double* mPoints, int32_t nSlave, const int32_t* sFaceOffsets,
                              const int32_t* sFaceLabels, int32_t nSlavePoints, const double* sPoints, double nonOverlapTol,
                              int rescale)
{
    if (!ctx) return B200_EINVAL;
    int rc;
    if ((rc = ggi_check_patch(ctx, "master", nMaster, mFaceOffsets, mFaceLabels, nMasterPoints, mPoints))) return rc;
    if ((rc = ggi_check_patch(ctx, "slave", nSlave, sFaceOffsets, sFaceLabels, nSlavePoints, sPoints))) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    try
    {
        std::vector<int32_t> candOff, cand;
        ggib::broad_phase(nMaster, mFaceOffsets, mFaceLabels, mPoints, nSlave, sFaceOffsets, sFaceLabels, sPoints, candOff, cand);
        const int32_t nPairs = (int32_t)cand.size();
        std::vector<double> w(nPairs, 0.0);
        if (nPairs)
        {
            std::vector<int32_t> pairMaster(nPairs);
            for (int32_t i = 0; i < nMaster; i++)
                for (int32_t k = candOff[i]; k < candOff[i + 1]; k++) pairMaster[k] = i;
            const std::vector<int32_t> hmOff(mFaceOffsets, mFaceOffsets + nMaster + 1), hmFp(mFaceLabels, mFaceLabels + mFaceOffsets[nMaster]);
            const std::vector<int32_t> hsOff(sFaceOffsets, sFaceOffsets + nSlave + 1), hsFp(sFaceLabels, sFaceLabels + sFaceOffsets[nSlave]);
            const std::vector<double> hmP(mPoints, mPoints + 3 * (size_t)nMasterPoints), hsP(sPoints, sPoints + 3 * (size_t)nSlavePoints);
            DevBuf<int> dPairMaster, dCand, dCandOff, dmOff, dmFp, dsOff, dsFp;
            DevBuf<double> dmP, dsP, dArea, dMArea, dW;
            CK(ctx, dPairMaster.upload(pairMaster, st));
            CK(ctx, dCand.upload(cand, st));
            CK(ctx, dCandOff.upload(candOff, st));
            CK(ctx, dmOff.upload(hmOff, st));
            CK(ctx, dmFp.upload(hmFp, st));
            CK(ctx, dsOff.upload(hsOff, st));
            CK(ctx, dsFp.upload(hsFp, st));
            CK(ctx, dmP.upload(hmP, st));
            CK(ctx, dsP.upload(hsP, st));
            CK(ctx, dArea.alloc(nPairs));
            CK(ctx, dMArea.alloc(nMaster));
            CK(ctx, dW.alloc(nPairs));
            CK(ctx, cudaMemsetAsync(dMArea.p, 0, sizeof(double) * nMaster, st));
            const int blocksP = (int)std::max<int64_t>(1, std::min<int64_t>((nPairs + 127) / 128, (int64_t)ctx->smCount * 16));
            const int blocksR = (int)std::max<int64_t>(1, std::min<int64_t>((nMaster + 127) / 128, (int64_t)ctx->smCount * 16));
            ctx->launches += 2;
            k_ggi_pairs<<<blocksP, 128, 0, st>>>(nPairs, dPairMaster.p, dCand.p, dmOff.p, dmFp.p, dmP.p, dsOff.p, dsFp.p, dsP.p,
                                                 dArea.p, dMArea.p);
            CK(ctx, cudaGetLastError());
            k_ggi_rows<<<blocksR, 128, 0, st>>>(nMaster, dCandOff.p, dArea.p, dMArea.p, nonOverlapTol, rescale, dW.p);
            CK(ctx, cudaGetLastError());
            CK(ctx, cudaMemcpyAsync(w.data(), dW.p, sizeof(double) * nPairs, cudaMemcpyDeviceToHost, st));
            CK(ctx, cudaStreamSynchronize(st));
        }
        ggib::compact(nMaster, candOff, cand, w.data(), ctx->ggiOff, ctx->ggiAddr, ctx->ggiW);
    }
    catch (const std::exception& e)
    {
        return set_err(ctx, B200_ENOMEM, "b200_ggi_build: %s", e.what());
    }
    return (int)ctx->ggiAddr.size();
}

extern "C" int b200_ggi_fetch(b200_ctx* ctx, int32_t nMaster, int32_t* offsets, int32_t* addr, double* weights)
{
    if (!ctx || !offsets) return set_err(ctx, B200_EINVAL, "b200_ggi_fetch: bad arguments");
    if ((int64_t)ctx->ggiOff.size() != (int64_t)nMaster + 1)
        return set_err(ctx, B200_ESTATE, "b200_ggi_fetch: no result of a b200_ggi_build with %d master faces", nMaster);
    std::copy(ctx->ggiOff.begin(), ctx->ggiOff.end(), offsets);
    if (!ctx->ggiAddr.empty())
    {
        if (!addr || !weights) return set_err(ctx, B200_EINVAL, "b200_ggi_fetch: bad arguments");
        std::copy(ctx->ggiAddr.begin(), ctx->ggiAddr.end(), addr);
        std::copy(ctx->ggiW.begin(), ctx->ggiW.end(), weights);
    }
    return B200_OK;
}
