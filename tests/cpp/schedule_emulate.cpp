// CPU emulation of the device tables built by csrc/schedule.hpp (TEST INFRASTRUCTURE).
// Walks the SELL layout and the sweep schedules exactly as the CUDA kernels do (warps in ticket
// order, lanes, steps) so that the table construction can be checked against the oracle on a
// machine without a GPU.  Not linked into libb200ldu.so.
#include "../../multiregionfoam_b200/csrc/schedule.hpp"

#include <cmath>
#include <cstdio>

using namespace b200;

extern "C" {

// single-region system; precondMode: 0 = DIC (upper both ways), 1 = DILU
// out: y = A x (SELL walk), w = M^-1 r (chain sweeps), rDout
int emu_run(int nCells, int nFaces, const int* l, const int* u, const double* diag, const double* upper,
            const double* lower, int dilu, const double* x, const double* r, double* y, double* w,
            double* rDout, int* stats)
{
    try
    {
        std::vector<RegionHost> regs(1);
        regs[0].nCells = nCells;
        regs[0].nFaces = nFaces;
        regs[0].l.assign(l, l + nFaces);
        regs[0].u.assign(u, u + nFaces);
        GlobalLdu g;
        g.build(regs);
        SellLayout sell;
        sell.build(g);
        std::vector<double> coef(2 * (size_t)nFaces);
        for (int f = 0; f < nFaces; f++)
        {
            coef[f] = upper[f];
            coef[nFaces + f] = lower ? lower[f] : upper[f];
        }
        // ---- Amul through SELL
        for (int c = 0; c < nCells; c++)
        {
            int64_t base = int64_t(sell.sliceOff[c >> 5]) * 32 + (c & 31);
            int width = sell.sliceOff[(c >> 5) + 1] - sell.sliceOff[c >> 5];
            double acc = diag[c] * x[c];
            for (int j = 0; j < width; j++)
            {
                int col = sell.col[base + int64_t(j) * 32];
                if (col >= 0) acc += coef[sell.src[base + int64_t(j) * 32]] * x[col];
            }
            y[c] = acc;
        }
        // ---- sweeps
        SweepSchedule fwd, bwd;
        fwd.build(g, +1);
        bwd.build(g, -1);
        const double* coefL = dilu ? coef.data() + nFaces : coef.data();
        const double* coefU = coef.data();
        const double NOTSET = -1.2345e300;
        // rD: forward schedule, division mode
        std::vector<double> rD(nCells, NOTSET);
        auto walk = [&](const SweepSchedule& S, int mode, const double* a, const double* b, const double* cA,
                        const double* cB, std::vector<double>& out) -> int {
            // mode 0: acc = a*b, acc -= (rD*cA[f])*out[col]; mode 1: acc = a, same; mode 2: acc = a, acc -= (cA*cB)/out
            for (int64_t wq = 0; wq < S.nWarps; wq++)
            {
                int nl = S.warpNLanes[wq], W = S.warpW[wq];
                for (int lane = 0; lane < nl; lane++)
                {
                    int start = S.laneStart[S.warpLaneBase[wq] + lane], len = S.laneLen[S.warpLaneBase[wq] + lane];
                    double prev = 0;
                    for (int s = 0; s < len; s++)
                    {
                        int row = start + S.dir * s;
                        double acc = mode == 0 ? a[row] * b[row] : a[row];
                        for (int j = 0; j < W; j++)
                        {
                            int64_t idx = S.warpOffBase[wq] + (int64_t(s) * W + j) * nl + lane;
                            int col = S.offCol[idx];
                            if (col < 0) continue;
                            int f = S.offFace[idx];
                            double v = out[col];
                            if (v == NOTSET) return -10; // dependency not yet produced: schedule order broken
                            if (mode == 2)
                                acc -= cA[f] * cB[f] / v;
                            else
                                acc -= rD[row] * cA[f] * v;
                        }
                        int cf = S.chainFace[S.warpChainBase[wq] + int64_t(s) * nl + lane];
                        if (cf >= 0)
                        {
                            if (mode == 2)
                                acc -= cA[cf] * cB[cf] / prev;
                            else
                                acc -= rD[row] * cA[cf] * prev;
                        }
                        else if (s > 0)
                            return -11;
                        out[row] = acc;
                        prev = acc;
                    }
                }
            }
            return 0;
        };
        int rc = walk(fwd, 2, diag, nullptr, coefU, coefL, rD);
        if (rc) return rc;
        for (int c = 0; c < nCells; c++) rD[c] = 1.0 / rD[c];
        for (int c = 0; c < nCells; c++) rDout[c] = rD[c];
        std::vector<double> tmp(nCells, NOTSET), wv(nCells, NOTSET);
        rc = walk(fwd, 0, rD.data(), r, coefL, nullptr, tmp);
        if (rc) return rc;
        rc = walk(bwd, 1, tmp.data(), nullptr, coefU, nullptr, wv);
        if (rc) return rc;
        for (int c = 0; c < nCells; c++) w[c] = wv[c];
        stats[0] = int(fwd.nWarps);
        stats[1] = fwd.nLevels;
        stats[2] = fwd.maxW;
        stats[3] = int(fwd.nChains);
        stats[4] = int(bwd.nWarps);
        stats[5] = bwd.nLevels;
        stats[6] = bwd.maxW;
        stats[7] = int(sell.nSlots);
        return 0;
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "emu_run: %s\n", e.what());
        return -1;
    }
}
}
