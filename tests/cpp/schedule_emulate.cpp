// CPU emulation of the device tables built by csrc/schedule.hpp (TEST INFRASTRUCTURE).
// Walks the slot-ordered SELL layout and the sweep streams exactly as the CUDA kernels do (warps in
// ticket order, time steps, lanes, shuffle / own / memory term codes) so that the table
// construction can be checked against the oracle on a machine without a GPU.
// Not linked into libb200ldu.so.
#include "../../multiregionfoam_b200/csrc/schedule.hpp"

#include <cmath>
#include <cstdio>

using namespace b200;

namespace
{
const double NOTSET = -1.2345e300;

// mode 0: acc = a*b, acc -= c*v ; mode 1: acc = a, acc -= c*v ; mode 2: acc = a, acc -= c/v
// coefficient c of a term: mode 0/1: rD[slot]*cA[face] ; mode 2: cA[face]*cB[face]
int walk(const PipeSchedule& S, const PipeSchedule::Dir& D, int dir, int mode, const std::vector<double>& a,
         const std::vector<double>& b, const double* cA, const double* cB, const std::vector<double>& rD,
         std::vector<double>& out)
{
    for (int ticket = 0; ticket < S.nGroups; ticket++)
    {
        const int g = dir > 0 ? ticket : S.orderB[ticket];
        const int W = D.gW[g], nT = S.gNT[g];
        double prev[32], cur[32];
        for (int l = 0; l < 32; l++) prev[l] = 0.0;
        for (int step = 0; step < nT; step++)
        {
            const int t = dir > 0 ? step : nT - 1 - step;
            for (int lane = 0; lane < 32; lane++)
            {
                const int64_t slot = (int64_t(S.gBase[g]) + t) * 32 + lane;
                double acc = mode == 0 ? a[slot] * b[slot] : a[slot];
                for (int j = 0; j < W; j++)
                {
                    const int64_t idx = D.gTermOff[g] + (int64_t(step) * W + j) * 32 + lane;
                    const int32_t code = D.code[idx];
                    if (code == kCodeNone) continue;
                    const int32_t f = D.face[idx];
                    double v;
                    if (code >= 0)
                    {
                        v = out[code];
                        if (v == NOTSET) return -10; // dependency not produced yet: ticket order broken
                    }
                    else if (code == kCodeOwn)
                        v = prev[lane];
                    else
                        v = prev[kCodeShfl - code];
                    if (mode == 2)
                        acc -= cA[f] * cB[f] / v;
                    else
                        acc -= rD[slot] * cA[f] * v;
                }
                if (S.cellOfSlot[slot] < 0 && acc != 0.0) return -12; // padding slots must stay zero
                cur[lane] = acc;
            }
            for (int lane = 0; lane < 32; lane++)
            {
                out[(int64_t(S.gBase[g]) + t) * 32 + lane] = cur[lane];
                prev[lane] = cur[lane];
            }
        }
    }
    return 0;
}
} // namespace

extern "C" {

// single-region system; dilu: 0 = DIC (upper both ways), 1 = DILU
// out: y = A x (SELL walk), w = M^-1 r (sweeps), rDout ; all in CELL order
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
        regs[0].set = true;
        GlobalLdu g;
        g.build(regs);
        PipeSchedule S;
        S.build(g, regs);
        SellLayout sell;
        sell.build(g, S);
        std::vector<double> coef(2 * (size_t)nFaces + 1);
        for (int f = 0; f < nFaces; f++)
        {
            coef[f] = upper[f];
            coef[nFaces + f] = lower ? lower[f] : upper[f];
        }
        const int64_t nS = S.nSlots;
        auto toSlots = [&](const double* v) {
            std::vector<double> s(nS, 0.0);
            for (int c = 0; c < nCells; c++) s[S.slotOfCell[c]] = v[c];
            return s;
        };
        // ---- Amul through SELL over slots
        {
            std::vector<double> xs = toSlots(x), ds = toSlots(diag), ys(nS, 0.0);
            for (int64_t sl = 0; sl < nS; sl++)
            {
                const int64_t base = int64_t(sell.sliceOff[sl >> 5]) * 32 + (sl & 31);
                const int width = sell.sliceOff[(sl >> 5) + 1] - sell.sliceOff[sl >> 5];
                double acc = ds[sl] * xs[sl];
                for (int j = 0; j < width; j++)
                {
                    const int col = sell.col[base + int64_t(j) * 32];
                    if (col >= 0) acc += coef[sell.src[base + int64_t(j) * 32]] * xs[col];
                }
                ys[sl] = acc;
            }
            for (int c = 0; c < nCells; c++) y[c] = ys[S.slotOfCell[c]];
            for (int64_t sl = 0; sl < nS; sl++)
                if (S.cellOfSlot[sl] < 0 && ys[sl] != 0.0) return -13;
        }
        // ---- sweeps
        const double* coefL = dilu ? coef.data() + nFaces : coef.data();
        const double* coefU = coef.data();
        std::vector<double> ds = toSlots(diag), rs = toSlots(r), none;
        std::vector<double> rDraw(nS, NOTSET), rD(nS, 0.0);
        int rc = walk(S, S.fwd, +1, 2, ds, none, coefU, coefL, none, rDraw);
        if (rc) return rc;
        for (int64_t sl = 0; sl < nS; sl++) rD[sl] = S.cellOfSlot[sl] >= 0 ? 1.0 / rDraw[sl] : 0.0;
        for (int c = 0; c < nCells; c++) rDout[c] = rD[S.slotOfCell[c]];
        std::vector<double> tmp(nS, NOTSET), wv(nS, NOTSET);
        rc = walk(S, S.fwd, +1, 0, rD, rs, coefL, nullptr, rD, tmp);
        if (rc) return rc;
        rc = walk(S, S.bwd, -1, 1, tmp, none, coefU, nullptr, rD, wv);
        if (rc) return rc;
        for (int c = 0; c < nCells; c++) w[c] = wv[S.slotOfCell[c]];
        stats[0] = S.nGroups;
        stats[1] = S.fwd.nLevels;
        stats[2] = S.fwd.maxW;
        stats[3] = int(S.nPaths);
        stats[4] = int(S.nLinkedGroups);
        stats[5] = S.bwd.nLevels;
        stats[6] = S.bwd.maxW;
        stats[7] = int(S.nSlots);
        stats[8] = S.nLineRegions;
        stats[9] = int(S.nMemTermsF);
        stats[10] = int(S.nShflTermsF);
        stats[11] = int(S.nOwnTermsF);
        return 0;
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "emu_run: %s\n", e.what());
        return -1;
    }
}
}
