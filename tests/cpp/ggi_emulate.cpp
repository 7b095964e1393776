// CPU emulation of the device-side GGI weight construction (TEST INFRASTRUCTURE): the same ggi_build.hpp the kernels
// k_ggi_pairs / k_ggi_rows of multiregionfoam_b200/csrc/ggi_build.cuh are built from, thread loops replaced by plain
// loops.  Build: g++ -O2 -ffp-contract=off (multiregionfoam_b200/build.py).
#include "../../multiregionfoam_b200/csrc/ggi_build.hpp"

// returns nnz (>= 0); when nnz > cap only offsets are written
extern "C" int emu_ggi_build(int32_t nM, const int32_t* mOff, const int32_t* mFp, const double* mPts, int32_t nS,
                             const int32_t* sOff, const int32_t* sFp, const double* sPts, double tol, int rescale,
                             int32_t* offsets, int32_t cap, int32_t* addr, double* weights)
{
    for (int32_t i = 0; i < nM; i++)
        if (mOff[i + 1] - mOff[i] < 3 || mOff[i + 1] - mOff[i] > ggib::kMaxV) return -1;
    for (int32_t j = 0; j < nS; j++)
        if (sOff[j + 1] - sOff[j] < 3 || sOff[j + 1] - sOff[j] > ggib::kMaxV) return -1;
    std::vector<int32_t> candOff, cand;
    ggib::broad_phase(nM, mOff, mFp, mPts, nS, sOff, sFp, sPts, candOff, cand);
    std::vector<double> area(cand.size()), mArea(nM, 0.0), w(cand.size());
    for (int32_t i = 0; i < nM; i++) // k_ggi_pairs: one thread per candidate pair
        for (int32_t k = candOff[i]; k < candOff[i + 1]; k++)
        {
            const int32_t j = cand[k];
            double ma;
            area[k] = ggib::pair_area(mPts, mFp + mOff[i], mOff[i + 1] - mOff[i], sPts, sFp + sOff[j], sOff[j + 1] - sOff[j], ma);
            mArea[i] = ma;
        }
    for (int32_t i = 0; i < nM; i++) // k_ggi_rows: one thread per master face
        ggib::row_weights(area.data(), mArea[i], candOff[i], candOff[i + 1], tol, rescale, w.data());
    std::vector<int32_t> off, a;
    std::vector<double> ww;
    ggib::compact(nM, candOff, cand, w.data(), off, a, ww);
    for (int32_t i = 0; i <= nM; i++) offsets[i] = off[i];
    if ((int32_t)a.size() <= cap)
        for (size_t k = 0; k < a.size(); k++)
        {
            addr[k] = a[k];
            weights[k] = ww[k];
        }
    return (int)a.size();
}
