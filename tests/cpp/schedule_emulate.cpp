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
        // hist[k][lane]: value lane produced k+1 time steps ago (own-lane terms use k = 0, shuffled terms k = kSkew-1)
        double hist[kSkew][32], cur[32];
        for (int k = 0; k < kSkew; k++)
            for (int l = 0; l < 32; l++) hist[k][l] = 0.0;
        const double(&prev)[32] = hist[0];
        const double(&prevS)[32] = hist[kSkew - 1];
        const bool fast = D.gFast[g] != 0;
        const int Lg = D.gLg[g], Rg = D.gRg[g], Kg = D.gKg[g];
        for (int step = 0; step < nT; step++)
        {
            const int t = dir > 0 ? step : nT - 1 - step;
            auto coefOf = [&](int32_t f, int64_t slot) { return mode == 2 ? cA[f] * cB[f] : rD[slot] * cA[f]; };
            for (int lane = 0; lane < 32; lane++)
            {
                const int64_t slot = (int64_t(S.gBase[g]) + t) * 32 + lane;
                double acc = mode == 0 ? a[slot] * b[slot] : a[slot];
                if (fast)
                {
                    // producer part: leading cross-group terms from the P-stream
                    const unsigned char* pr = D.pStream.data() + D.gPOff[g] + int64_t(step) * PipeSchedule::p_rec_bytes(Lg, Kg);
                    const int32_t* pCode = reinterpret_cast<const int32_t*>(pr + Lg * 256);
                    const int32_t* pConstArr = reinterpret_cast<const int32_t*>(pr + Lg * 384);
                    for (int i = 0; i < Lg; i++)
                    {
                        const int32_t code = pCode[i * 32 + lane];
                        const int32_t f = D.pFace[D.gPFaceOff[g] + (int64_t(step) * Lg + i) * 32 + lane];
                        if ((code >= 0) != (f >= 0)) return -20;
                        if (code < 0) continue;
                        if (out[code] == NOTSET) return -10;
                        acc = mode == 2 ? acc - coefOf(f, slot) / out[code] : acc - coefOf(f, slot) * out[code];
                    }
                    // consumer part: remaining terms from the C-stream
                    const unsigned char* cr = D.cStream.data() + D.gCOff[g] + PipeSchedule::c_meta_off(Rg, step);
                    const uint64_t meta = reinterpret_cast<const uint64_t*>(cr)[lane];
                    const int Rt = int((meta >> 48) & 0xff);
                    const bool general = (meta >> 56) & 1;
                    if (Rt > Rg) return -21;
                    const double acc0 = acc;
                    const bool dualStep = (meta >> 61) & 1, seamLane = (meta >> 58) & 1;
                    if (seamLane && !dualStep) return -27;
                    if (dualStep && general) return -28;
                    if (seamLane)
                    {
                        // seam form of a dual step: reference order own | A | B lives in planes 1 | 2 | 0
                        const unsigned selB = unsigned(meta >> 59) & 3u;
                        const int32_t fB = D.cFace[D.gCFaceOff[g] + (int64_t(step) * Rg + 0) * 32 + lane];
                        const int32_t fO = D.cFace[D.gCFaceOff[g] + (int64_t(step) * Rg + 1) * 32 + lane];
                        const int32_t fA = Rg > 2 ? D.cFace[D.gCFaceOff[g] + (int64_t(step) * Rg + 2) * 32 + lane] : -1;
                        if (fO < 0 || acc != (mode == 0 ? a[slot] * b[slot] : a[slot])) return -29; // no leading terms before the own-lane term
                        auto sub = [&](int32_t f, double v) { acc = mode == 2 ? acc - coefOf(f, slot) / v : acc - coefOf(f, slot) * v; };
                        sub(fO, prev[lane]);
                        if (fA >= 0)
                        {
                            const int32_t pc = pConstArr[lane];
                            if (pc < 0 || out[pc] == NOTSET) return -30;
                            sub(fA, out[pc]);
                        }
                        if (fB >= 0)
                        {
                            double v;
                            if (selB == 0)
                                v = prevS[(lane - dir) & 31];
                            else
                            {
                                const int32_t pc = pConstArr[(selB - 1) * 32 + lane];
                                if (pc < 0 || out[pc] == NOTSET) return -31;
                                v = out[pc];
                            }
                            sub(fB, v);
                        }
                        else if (selB != 0)
                            return -32;
                        if (mode != 2)
                        {
                            // the consumer's select-free evaluation of the seam form must give the same bits
                            const double cO = coefOf(fO, slot), cA = fA >= 0 ? coefOf(fA, slot) : 0.0, cB = fB >= 0 ? coefOf(fB, slot) : 0.0;
                            const double cv0 = Kg > 0 && pConstArr[lane] >= 0 ? out[pConstArr[lane]] : 0.0;
                            const double cv1 = Kg > 1 && pConstArr[32 + lane] >= 0 ? out[pConstArr[32 + lane]] : 0.0;
                            const double vB = selB == 2 ? cv1 : (selB == 1 ? cv0 : prevS[(lane - dir) & 31]);
                            const double fastAcc = ((acc0 - cO * prev[lane]) - cA * cv0) - cB * vB;
                            if (std::memcmp(&fastAcc, &acc, 8) != 0 && !(fastAcc == 0.0 && acc == 0.0)) return -33;
                        }
                    }
                    else
                    for (int r = 0; r < Rg; r++)
                    {
                        const unsigned byte = unsigned((meta >> (8 * r)) & 0xff);
                        const int32_t f = D.cFace[D.gCFaceOff[g] + (int64_t(step) * Rg + r) * 32 + lane];
                        if (byte & kMetaPad)
                        {
                            if (f >= 0) return -22;
                            continue; // padding
                        }
                        if (r >= Rt || f < 0) return -23;
                        double v;
                        if (byte & kMetaConst)
                        {
                            const int32_t pConst = pConstArr[((byte & 0x20u) ? 1 : 0) * 32 + lane];
                            if (!general || pConst < 0 || out[pConst] == NOTSET) return -24;
                            v = out[pConst];
                        }
                        else if (byte & kMetaOwn)
                            v = prev[lane];
                        else
                            v = prevS[byte & kMetaLane];
                        acc = mode == 2 ? acc - coefOf(f, slot) / v : acc - coefOf(f, slot) * v;
                    }
                    if (!general && !seamLane && mode != 2)
                    {
                        // the consumer's descriptor-free evaluation of a canonical step must give the same bits
                        const int32_t f0 = D.cFace[D.gCFaceOff[g] + (int64_t(step) * Rg + 0) * 32 + lane];
                        const int32_t f1 = D.cFace[D.gCFaceOff[g] + (int64_t(step) * Rg + 1) * 32 + lane];
                        const double c0 = f0 >= 0 ? coefOf(f0, slot) : 0.0, c1 = f1 >= 0 ? coefOf(f1, slot) : 0.0;
                        const double fastAcc = (acc0 - c0 * prevS[(lane - dir) & 31]) - c1 * prev[lane];
                        if (std::memcmp(&fastAcc, &acc, 8) != 0 && !(fastAcc == 0.0 && acc == 0.0)) return -25;
                        for (int r = 2; r < Rg; r++)
                            if (D.cFace[D.gCFaceOff[g] + (int64_t(step) * Rg + r) * 32 + lane] >= 0) return -26;
                    }
                }
                else
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
                            v = prevS[kCodeShfl - code];
                        acc = mode == 2 ? acc - coefOf(f, slot) / v : acc - coefOf(f, slot) * v;
                    }
                if (S.cellOfSlot[slot] < 0 && acc != 0.0) return -12; // padding slots must stay zero
                cur[lane] = acc;
            }
            for (int lane = 0; lane < 32; lane++)
            {
                out[(int64_t(S.gBase[g]) + t) * 32 + lane] = cur[lane];
                for (int k = kSkew - 1; k > 0; k--) hist[k][lane] = hist[k - 1][lane];
                hist[0][lane] = cur[lane];
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
        stats[12] = S.fwd.nFastGroups;
        stats[13] = S.bwd.nFastGroups;
        stats[14] = int(S.fwd.nGeneralSteps);
        stats[15] = int(S.bwd.nGeneralSteps);
        return 0;
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "emu_run: %s\n", e.what());
        return -1;
    }
}
}

// ------------------------------------------------------------------------------------------------
// One rank of a decomposed multi-region system on the CPU, driven by the SAME tables the device uses
// (slot permutation, SELL layout, interface plan, halo send/receive layout).  Used by the
// world_size-2 gloo test: pack -> exchange over torch.distributed -> Amul + interface update.
namespace
{
struct EmuSys
{
    int rank = 0;
    std::vector<RegionHost> regs;
    std::vector<std::vector<double>> diag, upper, lower;
    std::vector<std::vector<std::vector<double>>> bou;
    GlobalLdu g;
    PipeSchedule S;
    SellLayout sell;
    IfacePlan P;
    std::vector<double> coef, ifCoef, diagSlots;
};
} // namespace

extern "C" {

void* emu_sys_create(int nRegions, int rank)
{
    EmuSys* s = new EmuSys;
    s->rank = rank;
    s->regs.resize(nRegions);
    s->diag.resize(nRegions);
    s->upper.resize(nRegions);
    s->lower.resize(nRegions);
    s->bou.resize(nRegions);
    return s;
}

void emu_sys_destroy(void* h) { delete static_cast<EmuSys*>(h); }

int emu_sys_set_region(void* h, int r, int nCells, int nFaces, const int* l, const int* u, const double* diag,
                       const double* upper, const double* lower)
{
    EmuSys* s = static_cast<EmuSys*>(h);
    RegionHost& R = s->regs[r];
    R.nCells = nCells;
    R.nFaces = nFaces;
    R.l.assign(l, l + nFaces);
    R.u.assign(u, u + nFaces);
    R.set = true;
    s->diag[r].assign(diag, diag + nCells);
    s->upper[r].assign(upper, upper + nFaces);
    s->lower[r].assign(lower ? lower : upper, (lower ? lower : upper) + nFaces);
    return 0;
}

int emu_sys_add_iface(void* h, int r, int kind, int nFaces, const int* fc, int peerRank, int peerRegion, int peerIface,
                      int nPeerFaces, const int* go, const int* ga, const double* gw, const double* bou)
{
    EmuSys* s = static_cast<EmuSys*>(h);
    IfaceHost I;
    I.kind = kind;
    I.nFaces = nFaces;
    I.faceCells.assign(fc, fc + nFaces);
    I.peerRank = peerRank;
    I.peerRegion = peerRegion;
    I.peerIface = peerIface;
    I.nPeerFaces = nPeerFaces;
    I.identity = (go == nullptr);
    if (go)
    {
        I.ggiOffsets.assign(go, go + nFaces + 1);
        I.ggiAddr.assign(ga, ga + go[nFaces]);
        I.ggiWeights.assign(gw, gw + go[nFaces]);
    }
    s->regs[r].ifaces.push_back(I);
    s->bou[r].emplace_back(bou, bou + nFaces);
    return int(s->regs[r].ifaces.size()) - 1;
}

int emu_sys_finalize(void* h)
{
    EmuSys* s = static_cast<EmuSys*>(h);
    try
    {
        int64_t co = 0, fo = 0;
        for (auto& R : s->regs)
        {
            R.cellOffset = co;
            R.faceOffset = fo;
            co += R.nCells;
            fo += R.nFaces;
        }
        s->g.build(s->regs);
        s->S.build(s->g, s->regs);
        s->sell.build(s->g, s->S);
        s->P.build(s->regs, s->rank, s->S);
        s->coef.assign(2 * size_t(s->g.F) + 1, 0.0);
        s->diagSlots.assign(s->S.nSlots, 0.0);
        s->ifCoef.assign(size_t(s->P.nCoefs) + 1, 0.0);
        for (size_t r = 0; r < s->regs.size(); r++)
        {
            const RegionHost& R = s->regs[r];
            for (int f = 0; f < R.nFaces; f++)
            {
                s->coef[R.faceOffset + f] = s->upper[r][f];
                s->coef[s->g.F + R.faceOffset + f] = s->lower[r][f];
            }
            for (int c = 0; c < R.nCells; c++) s->diagSlots[s->S.slotOfCell[R.cellOffset + c]] = s->diag[r][c];
            for (size_t i = 0; i < R.ifaces.size(); i++)
                for (int f = 0; f < R.ifaces[i].nFaces; f++) s->ifCoef[R.ifaces[i].coefOffset + f] = s->bou[r][i][f];
        }
        return 0;
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "emu_sys_finalize: %s\n", e.what());
        return -1;
    }
}

int emu_sys_npeers(void* h) { return int(static_cast<EmuSys*>(h)->P.peers.size()); }

int emu_sys_peer(void* h, int p, int* rank, int* sendOff, int* nSend, int* recvOff, int* nRecv)
{
    EmuSys* s = static_cast<EmuSys*>(h);
    *rank = s->P.peers[p];
    *sendOff = s->P.sendOff[p];
    *nSend = s->P.sendOff[p + 1] - s->P.sendOff[p];
    *recvOff = s->P.recvOff[p];
    *nRecv = s->P.recvOff[p + 1] - s->P.recvOff[p];
    return 0;
}

// k_halo_pack: sendbuf[i] = x[sendCells[i]]   (x given in cell order)
int emu_sys_pack(void* h, const double* xCells, double* sendbuf)
{
    EmuSys* s = static_cast<EmuSys*>(h);
    for (size_t i = 0; i < s->P.sendCells.size(); i++) sendbuf[i] = xCells[s->S.cellOfSlot[s->P.sendCells[i]]];
    return 0;
}

// k_amul + k_iface over slots, result back in cell order
int emu_sys_amul(void* h, const double* xCells, const double* recv, double* yCells)
{
    EmuSys* s = static_cast<EmuSys*>(h);
    const int64_t nS = s->S.nSlots;
    std::vector<double> x(nS, 0.0), y(nS, 0.0);
    for (int64_t c = 0; c < s->g.N; c++) x[s->S.slotOfCell[c]] = xCells[c];
    for (int64_t sl = 0; sl < nS; sl++)
    {
        const int64_t base = int64_t(s->sell.sliceOff[sl >> 5]) * 32 + (sl & 31);
        const int width = s->sell.sliceOff[(sl >> 5) + 1] - s->sell.sliceOff[sl >> 5];
        double acc = s->diagSlots[sl] * x[sl];
        for (int j = 0; j < width; j++)
        {
            const int col = s->sell.col[base + int64_t(j) * 32];
            if (col >= 0) acc += s->coef[s->sell.src[base + int64_t(j) * 32]] * x[col];
        }
        y[sl] = acc;
    }
    const IfacePlan& P = s->P;
    for (size_t t = 0; t < P.rows.size(); t++)
    {
        double acc = y[P.rows[t]];
        for (int e = P.rowStart[t]; e < P.rowStart[t + 1]; e++)
        {
            double pnf;
            auto val = [&](int code) { return code >= 0 ? x[code] : recv[-1 - code]; };
            if (P.entCnt[e] == 0)
                pnf = val(P.entSrc[e]);
            else
            {
                pnf = 0.0;
                for (int k = 0; k < P.entCnt[e]; k++) pnf += val(P.gSrc[P.entSrc[e] + k]) * P.gW[P.entSrc[e] + k];
            }
            acc -= s->ifCoef[P.entCoef[e]] * pnf;
        }
        y[P.rows[t]] = acc;
    }
    for (int64_t c = 0; c < s->g.N; c++) yCells[c] = y[s->S.slotOfCell[c]];
    return 0;
}
}

// placement and per-group classification of a single-region system (debug / tests)
extern "C" int emu_placement(int nCells, int nFaces, const int* l, const int* u, int* grp, int* tim, int* lan, int* gInfo /* 8 per group */,
                             int capGroups)
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
        for (int c = 0; c < nCells; c++)
        {
            grp[c] = S.place_.grp[c];
            tim[c] = S.place_.tim[c];
            lan[c] = S.place_.lan[c];
        }
        for (int i = 0; i < S.nGroups && i < capGroups; i++)
        {
            gInfo[8 * i + 0] = S.fwd.gW[i];
            gInfo[8 * i + 1] = S.fwd.gFast[i];
            gInfo[8 * i + 2] = S.fwd.gLg[i];
            gInfo[8 * i + 3] = S.fwd.gRg[i];
            gInfo[8 * i + 4] = S.fwd.gKg[i];
            gInfo[8 * i + 5] = S.gNT[i];
            gInfo[8 * i + 6] = S.bwd.gFast[i];
            gInfo[8 * i + 7] = S.fwd.gShflMask[i] < 0;
        }
        return S.nGroups;
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "emu_placement: %s\n", e.what());
        return -1;
    }
}
