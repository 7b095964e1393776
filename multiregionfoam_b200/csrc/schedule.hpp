// schedule.hpp -- host-side (pure C++) construction of the device tables of one rank's coupled
// LDU system: the row-packed (sliced-ELL) Amul layout, the interface / halo plan and the
// chain-pipelined DIC/DILU sweep schedules.  No CUDA in this file: it is also compiled into the
// CPU schedule-emulation test (tests/cpp/schedule_emulate.cpp).
//
// Reference semantics restated (file list: DESIGN.md section 1):
//  * lduMatrix::Amul face loop: row c accumulates diag, then lower neighbours by ascending
//    column, then upper neighbours by ascending column (faces are in upper-triangular order).
//  * DIC/DILU precondition: forward sweep row c consumes its lower neighbours in ascending
//    column order (losort order), backward sweep row c consumes its upper neighbours in
//    DESCENDING column order; product order (rD*coef)*w.
//  * coupledLduMatrix::updateMatrixInterfaces: per row, non-processor interfaces in patch-list
//    order, then processor interfaces.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200
{

struct IfaceHost
{
    int kind = 0;
    int nFaces = 0;
    std::vector<int32_t> faceCells;
    int peerRank = 0, peerRegion = 0, peerIface = 0;
    int nPeerFaces = 0;
    bool identity = true;
    std::vector<int32_t> ggiOffsets, ggiAddr;
    std::vector<double> ggiWeights;
    int64_t coefOffset = 0; // into the concatenated interface-coefficient arrays
};

struct RegionHost
{
    int32_t nCells = 0, nFaces = 0;
    std::vector<int32_t> l, u;
    int64_t cellOffset = 0, faceOffset = 0;
    std::vector<IfaceHost> ifaces;
    bool set = false;
};

// ------------------------------------------------------------------------------------------
// Global (concatenated) LDU addressing with CSR views of the lower / upper entries of each row.
struct GlobalLdu
{
    int64_t N = 0, F = 0;
    std::vector<int32_t> L, U;       // [F] global lower / upper cell of each face
    std::vector<int32_t> ownerStart; // [N+1] faces of row c as lower cell: [ownerStart[c], ownerStart[c+1])
    std::vector<int32_t> losort;     // [F] faces sorted by upper cell (stable)
    std::vector<int32_t> losortStart; // [N+1]

    void build(const std::vector<RegionHost>& regs)
    {
        N = 0;
        F = 0;
        for (auto& r : regs)
        {
            N += r.nCells;
            F += r.nFaces;
        }
        if (N >= (int64_t(1) << 31) || F >= (int64_t(1) << 30))
            throw std::runtime_error("system too large for int32 labels on one rank");
        L.resize(F);
        U.resize(F);
        for (auto& r : regs)
            for (int64_t f = 0; f < r.nFaces; f++)
            {
                L[r.faceOffset + f] = int32_t(r.cellOffset + r.l[f]);
                U[r.faceOffset + f] = int32_t(r.cellOffset + r.u[f]);
            }
        ownerStart.assign(N + 1, 0);
        losortStart.assign(N + 1, 0);
        for (int64_t f = 0; f < F; f++)
        {
            ownerStart[L[f] + 1]++;
            losortStart[U[f] + 1]++;
        }
        for (int64_t c = 0; c < N; c++)
        {
            ownerStart[c + 1] += ownerStart[c];
            losortStart[c + 1] += losortStart[c];
        }
        for (int64_t f = 1; f < F; f++)
            if (L[f] < L[f - 1] || (L[f] == L[f - 1] && U[f] <= U[f - 1]))
                throw std::runtime_error("faces are not in upper-triangular order (sorted by owner, neighbour)");
        losort.resize(F);
        std::vector<int32_t> pos(losortStart.begin(), losortStart.end() - 1);
        for (int64_t f = 0; f < F; f++) losort[pos[U[f]]++] = int32_t(f);
    }
};

// ------------------------------------------------------------------------------------------
// Row-packed Amul layout: slices of 32 rows, column-major inside a slice (slot = base + j*32 + lane).
struct SellLayout
{
    int64_t nSlices = 0, nSlots = 0;
    std::vector<int32_t> sliceOff; // [nSlices+1] in units of 32 slots
    std::vector<int32_t> col;      // [nSlots] column or -1 (padding)
    std::vector<int32_t> src;      // [nSlots] index into coef = [upper(F) | lower(F)], -1 for padding

    void build(const GlobalLdu& g)
    {
        const int64_t N = g.N, F = g.F;
        nSlices = (N + 31) / 32;
        sliceOff.assign(nSlices + 1, 0);
        for (int64_t s = 0; s < nSlices; s++)
        {
            int w = 0;
            for (int64_t c = s * 32; c < std::min<int64_t>(N, s * 32 + 32); c++)
            {
                int cnt = (g.losortStart[c + 1] - g.losortStart[c]) + (g.ownerStart[c + 1] - g.ownerStart[c]);
                w = std::max(w, cnt);
            }
            int64_t next = int64_t(sliceOff[s]) + w;
            if (next >= (int64_t(1) << 31) / 32) throw std::runtime_error("SELL layout exceeds int32 slots");
            sliceOff[s + 1] = int32_t(next);
        }
        nSlots = int64_t(sliceOff[nSlices]) * 32;
        col.assign(nSlots, -1);
        src.assign(nSlots, -1);
        for (int64_t c = 0; c < N; c++)
        {
            int64_t base = int64_t(sliceOff[c >> 5]) * 32 + (c & 31);
            int j = 0;
            for (int32_t k = g.losortStart[c]; k < g.losortStart[c + 1]; k++, j++)
            {
                int32_t f = g.losort[k]; // row c = U[f], column L[f], coefficient lower[f]
                col[base + int64_t(j) * 32] = g.L[f];
                src[base + int64_t(j) * 32] = int32_t(F + f);
            }
            for (int32_t f = g.ownerStart[c]; f < g.ownerStart[c + 1]; f++, j++)
            {
                col[base + int64_t(j) * 32] = g.U[f]; // row c = L[f], column U[f], coefficient upper[f]
                src[base + int64_t(j) * 32] = f;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// Interface plan: rows touched by coupled patches, their ordered entries, and the halo layout.
struct IfacePlan
{
    std::vector<int32_t> rows;     // unique touched rows, ascending
    std::vector<int32_t> rowStart; // [rows+1]
    std::vector<int32_t> entCoef;  // index into the concatenated interface coefficient array
    std::vector<int32_t> entSrc;   // cnt==0: source code (>=0: x index, <0: recv index -1-k); else start into g*
    std::vector<int32_t> entCnt;   // 0 = identity
    std::vector<int32_t> gSrc;     // source codes of GGI terms
    std::vector<double> gW;
    std::vector<uint32_t> sliceMask; // per SELL slice: lanes whose row is touched
    // halo
    std::vector<int> peers;
    std::vector<int32_t> sendOff, recvOff; // [peers+1]
    std::vector<int32_t> sendCells;        // x indices packed into the send buffer
    int64_t nCoefs = 0;

    void build(std::vector<RegionHost>& regs, int myRank, int64_t N)
    {
        // coefficient offsets
        nCoefs = 0;
        for (auto& r : regs)
            for (auto& I : r.ifaces)
            {
                I.coefOffset = nCoefs;
                nCoefs += I.nFaces;
            }
        // peers
        for (auto& r : regs)
            for (auto& I : r.ifaces)
                if (I.peerRank != myRank && std::find(peers.begin(), peers.end(), I.peerRank) == peers.end())
                    peers.push_back(I.peerRank);
        std::sort(peers.begin(), peers.end());
        sendOff.assign(peers.size() + 1, 0);
        recvOff.assign(peers.size() + 1, 0);
        // per remote interface: offset of what the peer sends for it inside the peer's segment
        struct Rem
        {
            int region, iface;
        };
        std::vector<std::vector<int32_t>> recvSegOff(regs.size());
        for (size_t r = 0; r < regs.size(); r++) recvSegOff[r].assign(regs[r].ifaces.size(), -1);
        for (size_t p = 0; p < peers.size(); p++)
        {
            // send layout: my interfaces to this peer in (region, iface) order
            int32_t so = sendOff[p];
            for (size_t r = 0; r < regs.size(); r++)
                for (auto& I : regs[r].ifaces)
                    if (I.peerRank == peers[p])
                    {
                        for (int i = 0; i < I.nFaces; i++) sendCells.push_back(int32_t(regs[r].cellOffset + I.faceCells[i]));
                        so += I.nFaces;
                    }
            sendOff[p + 1] = so;
            // receive layout: the peer's interfaces to me in ITS (region, iface) order = my
            // interfaces to it sorted by (peerRegion, peerIface)
            std::vector<Rem> mine;
            for (size_t r = 0; r < regs.size(); r++)
                for (size_t i = 0; i < regs[r].ifaces.size(); i++)
                    if (regs[r].ifaces[i].peerRank == peers[p]) mine.push_back({int(r), int(i)});
            std::sort(mine.begin(), mine.end(), [&](const Rem& a, const Rem& b) {
                const IfaceHost& A = regs[a.region].ifaces[a.iface];
                const IfaceHost& B = regs[b.region].ifaces[b.iface];
                return A.peerRegion != B.peerRegion ? A.peerRegion < B.peerRegion : A.peerIface < B.peerIface;
            });
            int32_t ro = recvOff[p];
            for (auto& m : mine)
            {
                recvSegOff[m.region][m.iface] = ro;
                ro += regs[m.region].ifaces[m.iface].nPeerFaces;
            }
            recvOff[p + 1] = ro;
        }
        // entries, keyed (row, phase, iface, face)
        struct Ent
        {
            int32_t row;
            int32_t key;
            int32_t region, iface, face;
        };
        std::vector<Ent> ents;
        for (size_t r = 0; r < regs.size(); r++)
            for (size_t i = 0; i < regs[r].ifaces.size(); i++)
            {
                const IfaceHost& I = regs[r].ifaces[i];
                int32_t key = (I.kind == 1 ? (1 << 20) : 0) + int32_t(i);
                for (int f = 0; f < I.nFaces; f++)
                    ents.push_back({int32_t(regs[r].cellOffset + I.faceCells[f]), key, int32_t(r), int32_t(i), f});
            }
        std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) {
            return a.row != b.row ? a.row < b.row : a.key < b.key;
        });
        sliceMask.assign((N + 31) / 32, 0u);
        rowStart.push_back(0);
        for (size_t e = 0; e < ents.size(); e++)
        {
            const Ent& E = ents[e];
            if (rows.empty() || rows.back() != E.row)
            {
                if (!rows.empty()) rowStart.push_back(int32_t(entCoef.size()));
                rows.push_back(E.row);
                sliceMask[E.row >> 5] |= (1u << (E.row & 31));
            }
            const IfaceHost& I = regs[E.region].ifaces[E.iface];
            entCoef.push_back(int32_t(I.coefOffset + E.face));
            auto srcCode = [&](int peerFace) -> int32_t {
                if (I.peerRank == myRank)
                {
                    if (I.peerRegion < 0 || I.peerRegion >= int(regs.size()) || I.peerIface < 0 ||
                        I.peerIface >= int(regs[I.peerRegion].ifaces.size()))
                        throw std::runtime_error("interface peer (region, iface) does not exist on this rank");
                    const IfaceHost& Q = regs[I.peerRegion].ifaces[I.peerIface];
                    if (peerFace < 0 || peerFace >= Q.nFaces) throw std::runtime_error("GGI address out of range of the shadow patch");
                    return int32_t(regs[I.peerRegion].cellOffset + Q.faceCells[peerFace]);
                }
                if (peerFace < 0 || peerFace >= I.nPeerFaces) throw std::runtime_error("GGI address out of range of the shadow patch");
                return -1 - (recvSegOff[E.region][E.iface] + peerFace);
            };
            if (I.identity)
            {
                entSrc.push_back(srcCode(E.face));
                entCnt.push_back(0);
            }
            else
            {
                entSrc.push_back(int32_t(gSrc.size()));
                int cnt = I.ggiOffsets[E.face + 1] - I.ggiOffsets[E.face];
                entCnt.push_back(cnt);
                for (int k = I.ggiOffsets[E.face]; k < I.ggiOffsets[E.face + 1]; k++)
                {
                    gSrc.push_back(srcCode(I.ggiAddr[k]));
                    gW.push_back(I.ggiWeights[k]);
                }
                if (cnt == 0)
                {
                    // uncovered face: contributes coeff*0 (bridgeOverlap is off in every shipped case)
                    entCnt.back() = -1;
                }
            }
        }
        if (!rows.empty()) rowStart.push_back(int32_t(entCoef.size()));
    }
};

// ------------------------------------------------------------------------------------------
// Chain-pipelined sweep schedule (one per direction).
//
// A *chain* is a maximal run of consecutive rows c-1 -> c joined by a face; one thread walks a
// chain keeping the previous row's value in a register (it is always the LAST neighbour in the
// reference's order: the largest lower / smallest upper index).  All other neighbours are read
// from global memory, where every row value doubles as its own ready flag (sentinel until
// written).  Chains of equal *chain level* are independent and are packed 32 to a warp; warps are
// issued in level order through a ticket counter, so a warp only ever waits on warps that are
// already running: deadlock-free without co-residency requirements.
struct SweepSchedule
{
    int dir = +1; // +1 forward (lower neighbours), -1 backward (upper neighbours)
    int64_t nWarps = 0;
    int nLevels = 0;
    int maxW = 0;
    int64_t nChains = 0;
    std::vector<int32_t> warpNLanes, warpNSteps, warpW; // [nWarps]
    std::vector<int32_t> warpLaneBase;                  // [nWarps] into laneStart/laneLen
    std::vector<int64_t> warpChainBase;                 // [nWarps] into chainFace   (index = base + s*nl + lane)
    std::vector<int64_t> warpOffBase;                   // [nWarps] into offFace/Col (index = base + (s*W+j)*nl + lane)
    std::vector<int32_t> laneStart, laneLen;            // per lane: first row processed, chain length
    std::vector<int32_t> chainFace;                     // face of the in-register neighbour, -1 if none
    std::vector<int32_t> offFace, offCol;               // other neighbours in reference order, -1 padding
    int64_t nChainSlots = 0, nOffSlots = 0;

    void build(const GlobalLdu& g, int direction)
    {
        dir = direction;
        const int64_t N = g.N;
        // continuation flags: row c continues the chain of c-1 iff face (c-1, c) exists
        std::vector<uint8_t> cont(N, 0);
        for (int64_t c = 1; c < N; c++)
        {
            int32_t e = g.losortStart[c + 1];
            if (e > g.losortStart[c] && g.L[g.losort[e - 1]] == c - 1) cont[c] = 1;
        }
        // chains as [first,last] in ascending row order
        std::vector<int32_t> cFirst, cLast;
        std::vector<int32_t> chainOf(N);
        for (int64_t c = 0; c < N; c++)
        {
            if (!cont[c])
            {
                cFirst.push_back(int32_t(c));
                cLast.push_back(int32_t(c));
            }
            else
                cLast.back() = int32_t(c);
            chainOf[c] = int32_t(cFirst.size() - 1);
        }
        nChains = int64_t(cFirst.size());
        // chain levels
        std::vector<int32_t> lvl(nChains, 0);
        if (dir > 0)
        {
            for (int64_t ch = 0; ch < nChains; ch++)
            {
                int32_t lv = 0;
                for (int32_t c = cFirst[ch]; c <= cLast[ch]; c++)
                    for (int32_t k = g.losortStart[c]; k < g.losortStart[c + 1]; k++)
                    {
                        int32_t nb = g.L[g.losort[k]];
                        if (chainOf[nb] != ch) lv = std::max(lv, lvl[chainOf[nb]] + 1);
                    }
                lvl[ch] = lv;
            }
        }
        else
        {
            for (int64_t ch = nChains - 1; ch >= 0; ch--)
            {
                int32_t lv = 0;
                for (int32_t c = cFirst[ch]; c <= cLast[ch]; c++)
                    for (int32_t f = g.ownerStart[c]; f < g.ownerStart[c + 1]; f++)
                    {
                        int32_t nb = g.U[f];
                        if (chainOf[nb] != ch) lv = std::max(lv, lvl[chainOf[nb]] + 1);
                    }
                lvl[ch] = lv;
            }
        }
        nLevels = nChains ? (*std::max_element(lvl.begin(), lvl.end()) + 1) : 0;
        // order chains by (level, length descending, first row)
        std::vector<int32_t> order(nChains);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            if (lvl[a] != lvl[b]) return lvl[a] < lvl[b];
            int32_t la = cLast[a] - cFirst[a], lb = cLast[b] - cFirst[b];
            if (la != lb) return la > lb;
            return dir > 0 ? a < b : a > b;
        });
        // pack warps
        auto nOff = [&](int32_t c, bool isChainHead) -> int {
            // number of neighbours read from memory for row c
            if (dir > 0)
            {
                int n = g.losortStart[c + 1] - g.losortStart[c];
                return isChainHead ? n : n - 1;
            }
            int n = g.ownerStart[c + 1] - g.ownerStart[c];
            return isChainHead ? n : n - 1;
        };
        nChainSlots = 0;
        nOffSlots = 0;
        int64_t i = 0;
        while (i < nChains)
        {
            int64_t j = i;
            while (j < nChains && j - i < 32 && lvl[order[j]] == lvl[order[i]]) j++;
            int nl = int(j - i);
            int steps = 0, W = 0;
            for (int64_t k = i; k < j; k++)
            {
                int32_t ch = order[k];
                int len = cLast[ch] - cFirst[ch] + 1;
                steps = std::max(steps, len);
                for (int s = 0; s < len; s++)
                {
                    int32_t c = dir > 0 ? cFirst[ch] + s : cLast[ch] - s;
                    W = std::max(W, nOff(c, s == 0));
                }
            }
            warpNLanes.push_back(nl);
            warpNSteps.push_back(steps);
            warpW.push_back(W);
            warpLaneBase.push_back(int32_t(laneStart.size()));
            warpChainBase.push_back(nChainSlots);
            warpOffBase.push_back(nOffSlots);
            for (int64_t k = i; k < j; k++)
            {
                int32_t ch = order[k];
                laneStart.push_back(dir > 0 ? cFirst[ch] : cLast[ch]);
                laneLen.push_back(cLast[ch] - cFirst[ch] + 1);
            }
            nChainSlots += int64_t(steps) * nl;
            nOffSlots += int64_t(steps) * W * nl;
            maxW = std::max(maxW, W);
            i = j;
        }
        nWarps = int64_t(warpNLanes.size());
        chainFace.assign(nChainSlots, -1);
        offFace.assign(nOffSlots, -1);
        offCol.assign(nOffSlots, -1);
        for (int64_t w = 0; w < nWarps; w++)
        {
            const int nl = warpNLanes[w], W = warpW[w];
            for (int lane = 0; lane < nl; lane++)
            {
                const int32_t start = laneStart[warpLaneBase[w] + lane];
                const int32_t len = laneLen[warpLaneBase[w] + lane];
                for (int s = 0; s < len; s++)
                {
                    const int32_t c = start + dir * s;
                    int64_t cb = warpChainBase[w] + int64_t(s) * nl + lane;
                    int64_t ob = warpOffBase[w] + int64_t(s) * W * nl + lane;
                    int jj = 0;
                    if (dir > 0)
                    {
                        int32_t b = g.losortStart[c], e = g.losortStart[c + 1];
                        if (s > 0)
                        {
                            chainFace[cb] = g.losort[e - 1];
                            e--;
                        }
                        for (int32_t k = b; k < e; k++, jj++)
                        {
                            offFace[ob + int64_t(jj) * nl] = g.losort[k];
                            offCol[ob + int64_t(jj) * nl] = g.L[g.losort[k]];
                        }
                    }
                    else
                    {
                        int32_t b = g.ownerStart[c], e = g.ownerStart[c + 1];
                        if (s > 0)
                        {
                            chainFace[cb] = b; // smallest upper neighbour = c+1, consumed last
                            b++;
                        }
                        for (int32_t f = e - 1; f >= b; f--, jj++) // descending upper index
                        {
                            offFace[ob + int64_t(jj) * nl] = f;
                            offCol[ob + int64_t(jj) * nl] = g.U[f];
                        }
                    }
                }
            }
        }
    }
};

} // namespace b200
