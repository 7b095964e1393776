// schedule.hpp -- host-side (pure C++) construction of the device tables of one rank's coupled
// LDU system.  No CUDA in this file: it is also compiled into the CPU schedule-emulation test
// (tests/cpp/schedule_emulate.cpp), which walks the tables exactly as the kernels do.
//
// Central idea: every vector of the solver lives in SLOT ORDER, the order in which the DIC/DILU
// sweeps visit the rows.  Rows are packed into *groups* (one warp each) of 32 lanes x nT time
// steps; slot = (groupBase + t) * 32 + lane.  At time step t a warp finalises the (up to) 32 rows
// of that step; the forward sweep walks t upwards, the backward sweep walks the same groups with
// time reversed.  Two ways of forming groups, chosen per region:
//   LINE  mode (structured / blockMesh-like numbering): a lane walks a *path* of rows in which each
//         row has the previous one as a neighbour (x-lines, continued across block seams); lanes
//         of a warp are paths linked row-by-row (line j depends on line j-1) and are skewed by kSkew
//         time steps per lane: the own-lane dependency is one step old and stays in a register, the
//         neighbour-lane dependency is kSkew steps old and travels through a warp shuffle that is
//         issued ahead of its use.  Only dependencies on other warps go through memory.
//   BLOCK mode (unstructured numbering, no long lines): rows sorted by wavefront level are cut
//         into blocks of 32 independent rows; a group is 16 consecutive blocks.
// Cross-warp dependencies are resolved through the output vector itself: it is pre-filled with a
// sentinel, and a consumer re-reads until the value is there.  Warps start in a topological order
// of the group graph (ticket counter), so a warp only ever waits on warps that already run.
//
// Reference semantics preserved (file list: DESIGN.md section 1):
//  * lduMatrix::Amul face loop: row c accumulates diag, then lower neighbours by ascending
//    column, then upper neighbours by ascending column (faces are in upper-triangular order).
//  * DIC/DILU precondition: forward sweep row c consumes its lower neighbours in ascending
//    column order (losort order), backward sweep row c consumes its upper neighbours in
//    DESCENDING column order; product order (rD*coef)*w.
//  * coupledLduMatrix::updateMatrixInterfaces: per row, non-processor interfaces in patch-list
//    order, then processor interfaces.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <utility>
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
    bool attached = true;   // regionInterfaceType::attach()/detach(): a detached regionCouple patch takes no part in a solve
    // A shadow patch that decomposePar spread over several ranks (regionCouple / ggi with a global zone): the shadow ZONE has
    // nPeerFaces faces, the GGI addresses are zone face labels, and every piece says which zone faces it holds
    // (zoneAddressing of that rank's patch).  Empty: one piece {peerRank, peerRegion, peerIface} holding faces 0 .. nPeerFaces-1.
    struct Piece
    {
        int rank = 0, region = 0, iface = 0;
        std::vector<int32_t> zoneAddr;
    };
    std::vector<Piece> pieces;
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
    std::vector<int32_t> L, U;        // [F] global lower / upper cell of each face
    std::vector<int32_t> ownerStart;  // [N+1] faces of row c as lower cell: [ownerStart[c], ownerStart[c+1])
    std::vector<int32_t> losort;      // [F] faces sorted by upper cell (stable)
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
        if (N >= (int64_t(1) << 30) || F >= (int64_t(1) << 30))
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
    // lower neighbours of row c in ascending column order: faces losort[losortStart[c] .. losortStart[c+1])
    // upper neighbours of row c in ascending column order: faces [ownerStart[c] .. ownerStart[c+1])
};

// term codes of the sweep streams
constexpr int kSweepBlock = 8;     // steps per producer/consumer block of the sweep kernel (= producer warps)
#ifndef B200_SKEW
#define B200_SKEW 1
#endif
constexpr int kSkew = B200_SKEW;           // time steps between linked lanes of a warp: a value shuffled from another lane is kSkew
                                   // steps old.  2 takes the shuffle off the dependent chain of a time step (only mul+sub of
                                   // the own-lane term remain); 1 keeps it on the chain but halves the lag of a hop in j and
                                   // the spread of a block seam - measured faster on C2 (164.7 / 157.7 us per sweep against
                                   // 170 / 162.5) once the producers, not the consumer, set the pace
constexpr int32_t kCodeNone = -1; // padding term
constexpr int32_t kCodeOwn = -2;  // value this lane produced in the previous time step (register)
constexpr int32_t kCodeShfl = -3; // -3 - m: value lane m produced kSkew time steps ago (shuffle)
                                  // >= 0 : slot of a value produced elsewhere (memory, sentinel-guarded)
// term descriptor bytes of the consumer records (C-stream meta)
constexpr unsigned kMetaLane = 0x1fu;  // source lane of a shuffled term
constexpr unsigned kMetaOwn = 0x20u;   // (without kMetaConst) own-lane value of the previous step
constexpr unsigned kMetaPad = 0x40u;   // padding term (coefficient 0)
constexpr unsigned kMetaConst = 0x80u; // cross-group value handed over by the producers: constCode[bit 5]

// ------------------------------------------------------------------------------------------
struct PipeSchedule
{
    static constexpr int CH = 16;              // time steps per group are padded to a multiple of CH
    static constexpr int kStageBudget = 24576; // shared-memory bytes of one pipeline stage of the sweep kernel
    static constexpr int kLineModeMinAvgChain = 8;
    bool forceGeneric = false; // debug: every group takes the single-warp generic path
    bool mergeSeams = true;    // continue a lane's path across block seams (head of a chain has the tail of another as neighbour)

    int64_t N = 0, nSlots = 0;
    int nGroups = 0;
    std::vector<int32_t> slotOfCell, cellOfSlot;
    std::vector<int32_t> gBase, gNT; // per group, groups numbered in forward ticket order
    std::vector<int32_t> orderB;     // backward ticket order (group ids)

    struct Dir
    {
        std::vector<int32_t> gW, gCH, gShflMask;
        std::vector<int64_t> gTermOff;   // [nGroups+1] in terms; stream byte offset = 12 * gTermOff
        std::vector<int32_t> code, face; // [nTerms] index = gTermOff[g] + (step*W + j)*32 + lane, step in PROCESSING order
        int64_t nTerms = 0;
        int maxW = 0, maxStageBytes = 0, nLevels = 0;
        // ---- split streams of the fast path (groups with gFast != 0).  Everything about a row's terms that
        //      does not depend on the coefficient VALUES is decided here, once, on the host:
        //  P-stream (producer warps), per step:  coef[Lg][32] f64 | code[Lg][32] i32 | constCode[Kg][32] i32
        //      the LEADING terms of each row: its first run of cross-group (memory) terms, in reference order;
        //      constCode[k] = slot of the k-th cross-group value the row's remaining terms need (-1: none)
        //  C-stream (consumer warp), per step:   meta[32] u64 | coef[Rg][32] f64  (plane-major per block, c_meta_off)
        //      the REMAINING terms.  meta byte r < 6 describes the term of plane r: bits 0-4 source lane of a value
        //      shuffled from kSkew steps ago, kMetaOwn = own-lane value of the previous step, kMetaPad = padding
        //      (coefficient 0), kMetaConst = the value of constCode[bit 5]; byte 6 = number of planes in use;
        //      byte 7 bit 0 = the step needs the descriptor-driven path (uniform over the lanes, see build_split);
        //      bit 1 = every shuffled term of the step reads the linked lane; bit 5 = DUAL step (uniform): every lane is
        //      either in canonical form or in SEAM form - own-lane term FIRST, then up to two more terms - flagged per
        //      lane by bit 2, its planes then hold: 1 = own, 2 = first of two trailing terms (always cval 0), 0 = last
        //      trailing term, whose value bits 3-4 select (0 shuffled from the linked lane, 1 cval 0, 2 cval 1).
        std::vector<uint8_t> gFast;
        std::vector<int32_t> gLg, gRg, gKg;
        std::vector<int64_t> gPOff, gCOff;           // [nGroups+1] byte offsets into pStream / cStream
        std::vector<int64_t> gPFaceOff, gCFaceOff;   // [nGroups+1] offsets into pFace / cFace
        std::vector<unsigned char> pStream, cStream; // host images: codes / meta filled in, coefficients zero
        std::vector<int32_t> pFace, cFace;           // face of every coefficient slot (-1: none)
        std::vector<int64_t> gGenOff;                // [nGroups+1] term offsets of the unified stream: generic groups only
        int64_t nGenTerms = 0;
        int maxPStage = 0, maxCStep = 0, maxGenStage = 0, maxFastStage = 0, nFastGroups = 0;
        int64_t nGeneralSteps = 0; // steps of fast groups that are not canonical (see build_split)
        int64_t nDualSteps = 0;          // ... of which dual (canonical / seam form per lane, meta bit 61)
        int64_t nLinkedGeneralSteps = 0; // ... of which every shuffled term reads the linked lane (meta bit 57)
    } fwd, bwd;

    // statistics
    int nLineRegions = 0, nBlockRegions = 0;
    int64_t nPaths = 0, nLinkedGroups = 0, nMemTermsF = 0, nShflTermsF = 0, nOwnTermsF = 0;

    struct Placement
    {
        std::vector<int32_t> grp, tim, lan; // per cell
    };
    Placement place_;

    static int stage_steps(int W, int nVec)
    {
        int ch = CH;
        while (ch > 1 && ch * (W * 384 + nVec * 256) > kStageBudget) ch >>= 1;
        if (W <= 6 && ch < kSweepBlock) ch = kSweepBlock; // the fast path works on blocks of kSweepBlock steps
        return ch;
    }

    // Warps of linked x-lines are filled across the end of a sequence of linked lines (the j-lines of one k-plane: 164 lines
    // = 5 full warps + one of 4 lanes) with the head of a later sequence (a plane further on), where the dependencies allow
    // it: padding lanes cost slots, and slots cost bandwidth in every kernel of the solve.  If the merged warps make the group
    // graph cyclic (unstructured corner cases) the schedule is rebuilt without merging.
    bool slotOrderChain = false; // slot layout of the groups: ticket order, or chain order (see order_and_number; B200_SLOT_ORDER=chain)
    bool mergeSequences = true;
    double mergeGap = 1.5;      // measured on C3 (64 M cells): 0.5 .. 8 scanned, DESIGN.md section 4
    int mergeMinGroups = 444;   // 1.5 x the 296 resident sweep CTAs: below that the sweep is latency-bound and merging only adds waits
    void build(const GlobalLdu& g, const std::vector<RegionHost>& regs)
    {
        if (const char* e = getenv("B200_SLOT_ORDER")) slotOrderChain = std::string(e) == "chain"; // developer knobs
        if (const char* e = getenv("B200_MERGE_SEQ")) mergeSequences = atoi(e) != 0;
        if (const char* e = getenv("B200_MERGE_GAP")) mergeGap = atof(e);
        if (const char* e = getenv("B200_MERGE_MIN_GROUPS")) mergeMinGroups = atoi(e);
        if (mergeSequences)
        {
            try
            {
                build_impl(g, regs);
                return;
            }
            catch (const std::runtime_error&)
            {
                mergeSequences = false;
            }
        }
        build_impl(g, regs);
    }
    void build_impl(const GlobalLdu& g, const std::vector<RegionHost>& regs)
    {
        N = g.N;
        nLineRegions = nBlockRegions = 0;
        nPaths = nLinkedGroups = nMemTermsF = nShflTermsF = nOwnTermsF = 0;
        Placement P;
        P.grp.assign(N, -1);
        P.tim.assign(N, -1);
        P.lan.assign(N, -1);
        std::vector<int32_t> grpNT; // provisional groups (unordered)
        for (auto& R : regs)
        {
            if (R.nCells == 0) continue;
            const int64_t c0 = R.cellOffset, c1 = R.cellOffset + R.nCells;
            int64_t nChains = 0;
            for (int64_t c = c0; c < c1; c++)
                if (!continues_chain(g, c, c0)) nChains++;
            bool line = (R.nCells / std::max<int64_t>(1, nChains)) >= kLineModeMinAvgChain;
            bool ok = false;
            if (line)
            {
                const size_t mark = grpNT.size();
                ok = place_lines(g, c0, c1, P, grpNT);
                if (!ok)
                { // undo and fall back
                    grpNT.resize(mark);
                    for (int64_t c = c0; c < c1; c++) P.grp[c] = P.tim[c] = P.lan[c] = -1;
                }
                else
                    nLineRegions++;
            }
            if (!ok)
            {
                place_blocks(g, c0, c1, P, grpNT);
                nBlockRegions++;
            }
        }
        order_and_number(g, P, grpNT);
        build_terms(g, +1, fwd);
        build_terms(g, -1, bwd);
        build_split(+1, fwd);
        build_split(-1, bwd);
    }

  private:
    static bool continues_chain(const GlobalLdu& g, int64_t c, int64_t c0)
    {
        if (c == c0) return false;
        const int32_t e = g.losortStart[c + 1];
        return e > g.losortStart[c] && g.L[g.losort[e - 1]] == c - 1;
    }

    // ------------------------------------------------------------------ LINE mode
    bool place_lines(const GlobalLdu& g, int64_t c0, int64_t c1, Placement& P, std::vector<int32_t>& grpNT)
    {
        const int64_t n = c1 - c0;
        // chains: maximal runs of consecutive rows c-1 -> c joined by a face
        std::vector<int32_t> cFirst, cLast, chainOf(n);
        for (int64_t c = c0; c < c1; c++)
        {
            if (!continues_chain(g, c, c0))
            {
                cFirst.push_back(int32_t(c));
                cLast.push_back(int32_t(c));
            }
            else
                cLast.back() = int32_t(c);
            chainOf[c - c0] = int32_t(cFirst.size() - 1);
        }
        const int nCh = int(cFirst.size());
        // merge chains into paths across seams: head(B) has tail(A) as a lower neighbour
        std::vector<int32_t> nextCh(nCh, -1), prevCh(nCh, -1);
        for (int b = 0; b < nCh && mergeSeams; b++)
        {
            const int32_t h = cFirst[b];
            for (int32_t k = g.losortStart[h + 1] - 1; k >= g.losortStart[h]; k--)
            {
                const int32_t nb = g.L[g.losort[k]];
                const int a = chainOf[nb - c0];
                if (cLast[a] == nb && nextCh[a] < 0 && a != b)
                {
                    nextCh[a] = b;
                    prevCh[b] = a;
                    break;
                }
            }
        }
        for (int attempt = 0; attempt < 2; attempt++)
        {
            std::vector<int32_t> pathOf(n, -1), posOf(n, -1);
            std::vector<std::vector<int32_t>> pathChains;
            for (int a = 0; a < nCh; a++)
                if (prevCh[a] < 0)
                {
                    pathChains.emplace_back();
                    int32_t pos = 0;
                    for (int b = a; b >= 0; b = nextCh[b])
                    {
                        pathChains.back().push_back(b);
                        for (int32_t c = cFirst[b]; c <= cLast[b]; c++)
                        {
                            pathOf[c - c0] = int32_t(pathChains.size() - 1);
                            posOf[c - c0] = pos++;
                        }
                    }
                }
            const int nP = int(pathChains.size());
            std::vector<int32_t> pLen(nP, 0), pFirst(nP);
            for (int p = 0; p < nP; p++)
            {
                pFirst[p] = cFirst[pathChains[p][0]];
                for (int b : pathChains[p]) pLen[p] += cLast[b] - cFirst[b] + 1;
            }
            // path dependencies: (pred path, all dependencies aligned = same position in both paths)
            std::vector<std::vector<std::pair<int32_t, bool>>> preds(nP);
            {
                std::vector<int32_t> seen(nP, -1), idx(nP, -1);
                for (int p = 0; p < nP; p++)
                    for (int b : pathChains[p])
                        for (int32_t c = cFirst[b]; c <= cLast[b]; c++)
                            for (int32_t k = g.losortStart[c]; k < g.losortStart[c + 1]; k++)
                            {
                                const int32_t nb = g.L[g.losort[k]];
                                const int q = pathOf[nb - c0];
                                if (q == p) continue;
                                const bool aligned = posOf[nb - c0] == posOf[c - c0];
                                if (seen[q] != p)
                                {
                                    seen[q] = p;
                                    idx[q] = int32_t(preds[p].size());
                                    preds[p].push_back({q, aligned});
                                }
                                else if (!aligned)
                                    preds[p][idx[q]].second = false;
                            }
            }
            // path levels, and acyclicity of the path graph
            std::vector<int32_t> pLevel(nP, 0), indeg(nP, 0);
            std::vector<std::vector<int32_t>> succs(nP);
            for (int p = 0; p < nP; p++)
                for (auto& e : preds[p])
                {
                    succs[e.first].push_back(p);
                    indeg[p]++;
                }
            std::vector<int32_t> queue;
            for (int p = 0; p < nP; p++)
                if (!indeg[p]) queue.push_back(p);
            for (size_t qi = 0; qi < queue.size(); qi++)
            {
                const int p = queue[qi];
                for (int s : succs[p])
                {
                    pLevel[s] = std::max(pLevel[s], pLevel[p] + 1);
                    if (--indeg[s] == 0) queue.push_back(s);
                }
            }
            if (int(queue.size()) != nP)
            {
                if (attempt == 0)
                { // seam merging made the path graph cyclic: use plain chains
                    std::fill(nextCh.begin(), nextCh.end(), -1);
                    std::fill(prevCh.begin(), prevCh.end(), -1);
                    continue;
                }
                return false;
            }
            // link predecessor: the nearest aligned predecessor path; each path links at most one successor
            std::vector<int32_t> linkPred(nP, -1), linkSucc(nP, -1);
            for (int p = 0; p < nP; p++)
            {
                int best = -1;
                for (auto& e : preds[p])
                    if (e.second && linkSucc[e.first] < 0 && pFirst[e.first] < pFirst[p] && (best < 0 || pFirst[e.first] > pFirst[best]))
                        best = e.first;
                if (best >= 0)
                {
                    linkPred[p] = best;
                    linkSucc[best] = p;
                }
            }
            // sequences of linked paths, cut into warps of <= 32 lanes (skew = lane)
            std::vector<std::vector<int32_t>> groups; // lanes -> path
            std::vector<int32_t> singles;
            std::vector<std::vector<int32_t>> seqs;
            for (int p = 0; p < nP; p++)
            {
                if (linkPred[p] >= 0) continue;
                std::vector<int32_t> seq;
                for (int q = p; q >= 0; q = linkSucc[q]) seq.push_back(q);
                seqs.push_back(std::move(seq));
            }
            // admit q as the next lane only if all its dependencies on lanes already in this warp are exactly kSkew
            // time steps old - the aligned link to the previous lane - and no lane of the warp depends on it
            auto admit = [&](std::vector<int32_t>& lanes, int q) {
                for (auto& e : preds[q])
                {
                    auto it = std::find(lanes.begin(), lanes.end(), e.first);
                    if (it == lanes.end()) continue;
                    if (!(e.second && size_t(it - lanes.begin()) == lanes.size() - 1)) return false;
                }
                for (int sq : succs[q])
                    if (std::find(lanes.begin(), lanes.end(), sq) != lanes.end()) return false;
                lanes.push_back(q);
                return true;
            };
            // The last warp of a sequence (the j-lines of one k-plane: 164 lines = 5 warps + 4 lanes) is filled with the
            // first lines of a LATER sequence.  The warp starts when the sequence's chain of warps has reached its end, so
            // the partner must be a sequence that would not have started earlier anyway: its head lies mergeGap levels per
            // warp of this sequence further down the path graph (the next plane would wait for this whole plane and
            // serialise the sweep; measured: 4.6 x slower).
            const int S = int(seqs.size());
            std::vector<size_t> headOff(S, 0);
            // only where the sweep is bound by throughput: many more warps than the device keeps resident (the region's
            // count scaled to the whole system)
            int64_t estWarps = 0;
            for (const auto& seq : seqs) estWarps += int64_t(seq.size() + 31) / 32;
            const bool merge = mergeSequences && double(estWarps) * double(N) / double(n) >= double(mergeMinGroups);
            for (int si = 0; si < S; si++)
            {
                const auto& seq = seqs[si];
                size_t i = headOff[si];
                int nWarps = 0;
                while (i < seq.size())
                {
                    std::vector<int32_t> lanes{seq[i]};
                    size_t j = i + 1;
                    for (; j < seq.size() && lanes.size() < 32; j++)
                        if (!admit(lanes, seq[j])) break;
                    i = j;
                    nWarps++;
                    if (merge && i >= seq.size() && lanes.size() < 32)
                    {
                        const int need = pLevel[seq[0]] + int(std::ceil(mergeGap * nWarps));
                        int tried = 0;
                        for (int sj = si + 1; sj < S && tried < 4; sj++)
                        {
                            if (headOff[sj] != 0 || pLevel[seqs[sj][0]] < need) continue;
                            tried++;
                            size_t t = 0;
                            while (t < seqs[sj].size() && lanes.size() < 32 && admit(lanes, seqs[sj][t])) t++;
                            if (t > 0)
                            {
                                headOff[sj] = t;
                                break;
                            }
                        }
                    }
                    if (lanes.size() == 1)
                        singles.push_back(lanes[0]);
                    else
                        groups.push_back(lanes);
                }
            }
            nLinkedGroups += int64_t(groups.size());
            nPaths += nP;
            // singles: pack by (level, length descending); no skew
            std::stable_sort(singles.begin(), singles.end(), [&](int a, int b) {
                if (pLevel[a] != pLevel[b]) return pLevel[a] < pLevel[b];
                return pLen[a] > pLen[b];
            });
            std::vector<uint8_t> skewed(groups.size(), 1);
            for (size_t i = 0; i < singles.size();)
            {
                size_t j = i;
                std::vector<int32_t> lanes;
                while (j < singles.size() && lanes.size() < 32 && pLevel[singles[j]] == pLevel[singles[i]]) lanes.push_back(singles[j++]);
                groups.push_back(lanes);
                skewed.push_back(0);
                i = j;
            }
            for (size_t gi = 0; gi < groups.size(); gi++)
            {
                const int gid = int(grpNT.size());
                int nT = 0;
                for (size_t ln = 0; ln < groups[gi].size(); ln++)
                {
                    const int p = groups[gi][ln];
                    const int skew = skewed[gi] ? kSkew * int(ln) : 0;
                    nT = std::max(nT, skew + pLen[p]);
                    int32_t pos = 0;
                    for (int b : pathChains[p])
                        for (int32_t c = cFirst[b]; c <= cLast[b]; c++, pos++)
                        {
                            P.grp[c] = gid;
                            P.tim[c] = skew + pos;
                            P.lan[c] = int32_t(ln);
                        }
                }
                grpNT.push_back(nT);
            }
            return true;
        }
        return false;
    }

    // ------------------------------------------------------------------ BLOCK mode
    void place_blocks(const GlobalLdu& g, int64_t c0, int64_t c1, Placement& P, std::vector<int32_t>& grpNT)
    {
        const int64_t n = c1 - c0;
        std::vector<int32_t> lev(n, 0);
        int32_t nLev = 0;
        for (int64_t c = c0; c < c1; c++)
        {
            int32_t lv = 0;
            for (int32_t k = g.losortStart[c]; k < g.losortStart[c + 1]; k++) lv = std::max(lv, lev[g.L[g.losort[k]] - c0] + 1);
            lev[c - c0] = lv;
            nLev = std::max(nLev, lv + 1);
        }
        std::vector<int32_t> cnt(nLev + 1, 0);
        for (int64_t i = 0; i < n; i++) cnt[lev[i] + 1]++;
        for (int32_t l = 0; l < nLev; l++) cnt[l + 1] += cnt[l];
        std::vector<int32_t> byLevel(n);
        {
            std::vector<int32_t> pos(cnt.begin(), cnt.end() - 1);
            for (int64_t i = 0; i < n; i++) byLevel[pos[lev[i]]++] = int32_t(c0 + i);
        }
        int gid = -1, t = CH;
        for (int32_t l = 0; l < nLev; l++)
            for (int32_t i = cnt[l]; i < cnt[l + 1]; i += 32)
            {
                if (t == CH)
                {
                    gid = int(grpNT.size());
                    grpNT.push_back(0);
                    t = 0;
                }
                const int32_t e = std::min(i + 32, cnt[l + 1]);
                for (int32_t k = i; k < e; k++)
                {
                    const int32_t c = byLevel[k];
                    P.grp[c] = gid;
                    P.tim[c] = t;
                    P.lan[c] = k - i;
                }
                t++;
                grpNT[gid] = t;
            }
    }

    // ------------------------------------------------------------------ ordering, slots
    void order_and_number(const GlobalLdu& g, Placement& P, std::vector<int32_t>& grpNT)
    {
        const int nG = int(grpNT.size());
        // group graph: edge H -> G if a row of G has a lower neighbour in H
        std::vector<std::vector<int32_t>> succ(nG);
        std::vector<int32_t> indeg(nG, 0);
        std::unordered_map<uint64_t, int64_t> pairFaces; // faces between two groups (unordered pair)
        {
            uint64_t lastKey = ~uint64_t(0);
            int64_t lastCount = 0;
            std::vector<int64_t> edges;
            for (int64_t f = 0; f < g.F; f++)
            {
                const int32_t a = P.grp[g.L[f]], b = P.grp[g.U[f]];
                if (a != b)
                {
                    edges.push_back((int64_t(a) << 32) | uint32_t(b));
                    const uint64_t key = (uint64_t(uint32_t(std::min(a, b))) << 32) | uint32_t(std::max(a, b));
                    if (key == lastKey)
                        lastCount++;
                    else
                    {
                        if (lastCount) pairFaces[lastKey] += lastCount;
                        lastKey = key;
                        lastCount = 1;
                    }
                }
                else if (P.tim[g.L[f]] >= P.tim[g.U[f]])
                    throw std::runtime_error("internal: in-warp dependency does not point back in time");
                if (edges.size() > (size_t(1) << 22))
                {
                    std::sort(edges.begin(), edges.end());
                    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
                }
            }
            if (lastCount) pairFaces[lastKey] += lastCount;
            std::sort(edges.begin(), edges.end());
            edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
            for (int64_t e : edges)
            {
                const int32_t a = int32_t(e >> 32), b = int32_t(e & 0xffffffff);
                succ[a].push_back(b);
                indeg[b]++;
            }
        }
        std::vector<int32_t> levF(nG, 0), levB(nG, 0), topo;
        {
            std::vector<int32_t> deg(indeg);
            for (int a = 0; a < nG; a++)
                if (!deg[a]) topo.push_back(a);
            for (size_t qi = 0; qi < topo.size(); qi++)
            {
                const int a = topo[qi];
                for (int b : succ[a])
                {
                    levF[b] = std::max(levF[b], levF[a] + 1);
                    if (--deg[b] == 0) topo.push_back(b);
                }
            }
            if (int(topo.size()) != nG) throw std::runtime_error("internal: group graph is cyclic");
            for (int i = nG - 1; i >= 0; i--)
            {
                const int a = topo[i];
                for (int b : succ[a]) levB[a] = std::max(levB[a], levB[b] + 1);
            }
        }
        std::vector<int32_t> ordF(nG), newId(nG);
        std::iota(ordF.begin(), ordF.end(), 0);
        std::stable_sort(ordF.begin(), ordF.end(), [&](int a, int b) { return levF[a] < levF[b]; });
        for (int i = 0; i < nG; i++) newId[ordF[i]] = i;
        nGroups = nG;
        gNT.resize(nG);
        gBase.resize(nG);
        std::vector<int32_t> lb(nG);
        for (int i = 0; i < nG; i++)
        {
            const int old = ordF[i];
            gNT[i] = (grpNT[old] + CH - 1) / CH * CH;
            lb[i] = levB[old];
        }
        // Where the groups lie in slot space does not have to follow the ticket order.  Chain order (optional): the next
        // group in memory is the unplaced group that shares the most faces with the one just placed (the k-column of a
        // j-block), starting a new chain at the lowest ticket left - the groups Amul gathers x from then lie next to each
        // other.  Measured on the B200: no difference (C3 Amul 1 254 -> 1 247 us; Amul already moves its bytes at 0.94 of the
        // copy bandwidth, its traffic is the 104 B per slot it needs plus ~10 %), so ticket order stays the default.
        std::vector<int32_t> layout;
        layout.reserve(nG);
        if (slotOrderChain)
        {
            std::vector<std::vector<std::pair<int32_t, int64_t>>> nbrs(nG); // in ticket ids
            for (auto& kv : pairFaces)
            {
                const int a = newId[int32_t(kv.first >> 32)], b = newId[int32_t(kv.first & 0xffffffffu)];
                nbrs[a].push_back({b, kv.second});
                nbrs[b].push_back({a, kv.second});
            }
            std::vector<uint8_t> placed(nG, 0);
            for (int start = 0; start < nG; start++)
            {
                for (int cur = placed[start] ? -1 : start; cur >= 0;)
                {
                    placed[cur] = 1;
                    layout.push_back(cur);
                    int best = -1;
                    int64_t bestW = 0;
                    for (auto& e : nbrs[cur])
                        if (!placed[e.first] && (e.second > bestW || (e.second == bestW && best >= 0 && e.first < best)))
                        {
                            best = e.first;
                            bestW = e.second;
                        }
                    cur = best;
                }
            }
        }
        else
            for (int i = 0; i < nG; i++) layout.push_back(i);
        int64_t base = 0;
        for (int i : layout)
        {
            gBase[i] = int32_t(base);
            base += gNT[i];
        }
        if (base * 32 >= (int64_t(1) << 31)) throw std::runtime_error("slot space exceeds int32");
        nSlots = base * 32;
        fwd.nLevels = nG ? *std::max_element(levF.begin(), levF.end()) + 1 : 0;
        bwd.nLevels = nG ? *std::max_element(levB.begin(), levB.end()) + 1 : 0;
        orderB.resize(nG);
        std::iota(orderB.begin(), orderB.end(), 0);
        std::stable_sort(orderB.begin(), orderB.end(), [&](int a, int b) { return lb[a] != lb[b] ? lb[a] < lb[b] : a > b; });
        slotOfCell.resize(N);
        cellOfSlot.assign(nSlots, -1);
        for (int64_t c = 0; c < N; c++)
        {
            P.grp[c] = newId[P.grp[c]];
            const int64_t s = (int64_t(gBase[P.grp[c]]) + P.tim[c]) * 32 + P.lan[c];
            slotOfCell[c] = int32_t(s);
            cellOfSlot[s] = int32_t(c);
        }
        place_ = std::move(P);
    }

    // ------------------------------------------------------------------ sweep terms
    void build_terms(const GlobalLdu& g, int dir, Dir& D)
    {
        const int nG = nGroups;
        D.gW.assign(nG, 0);
        for (int64_t c = 0; c < N; c++)
        {
            const int cnt = dir > 0 ? g.losortStart[c + 1] - g.losortStart[c] : g.ownerStart[c + 1] - g.ownerStart[c];
            int32_t& w = D.gW[place_.grp[c]];
            w = std::max(w, cnt);
        }
        D.gTermOff.assign(nG + 1, 0);
        D.gCH.assign(nG, CH);
        D.gShflMask.assign(nG, 0);
        D.maxW = 0;
        D.maxStageBytes = 0;
        const int nVec = dir > 0 ? 2 : 1;
        for (int i = 0; i < nG; i++)
        {
            D.gTermOff[i + 1] = D.gTermOff[i] + int64_t(gNT[i]) * D.gW[i] * 32;
            D.gCH[i] = stage_steps(D.gW[i], nVec);
            D.maxW = std::max(D.maxW, D.gW[i]);
            D.maxStageBytes = std::max(D.maxStageBytes, D.gCH[i] * (D.gW[i] * 384 + nVec * 256));
        }
        D.nTerms = D.gTermOff[nG];
        D.code.assign(D.nTerms, kCodeNone);
        D.face.assign(D.nTerms, -1);
        for (int64_t c = 0; c < N; c++)
        {
            const int gI = place_.grp[c], t = place_.tim[c], ln = place_.lan[c];
            const int W = D.gW[gI];
            const int step = dir > 0 ? t : gNT[gI] - 1 - t;
            const int64_t base = D.gTermOff[gI] + int64_t(step) * W * 32 + ln;
            int j = 0;
            auto put = [&](int32_t f, int32_t nb) {
                int32_t code;
                if (place_.grp[nb] == gI && place_.lan[nb] == ln && place_.tim[nb] == t - dir)
                {
                    code = kCodeOwn;
                    if (dir > 0) nOwnTermsF++;
                }
                else if (place_.grp[nb] == gI && place_.tim[nb] == t - kSkew * dir && (kSkew > 1 || place_.lan[nb] != ln))
                {
                    code = kCodeShfl - place_.lan[nb];
                    if (j < 31) D.gShflMask[gI] |= (1 << j);
                    if (dir > 0) nShflTermsF++;
                }
                else
                {
                    code = slotOfCell[nb];
                    if (dir > 0) nMemTermsF++;
                    // a memory dependency on a step of the same group inside the same block of kSweepBlock steps would
                    // dead-lock the producer/consumer CTA (the consumer needs the whole block): such groups
                    // take the single-warp generic path (bit 31 of gShflMask)
                    if (place_.grp[nb] == gI)
                    {
                        const int stepNb = dir > 0 ? place_.tim[nb] : gNT[gI] - 1 - place_.tim[nb];
                        if (stepNb / kSweepBlock == step / kSweepBlock) D.gShflMask[gI] |= int32_t(0x80000000u);
                    }
                }
                D.code[base + int64_t(j) * 32] = code;
                D.face[base + int64_t(j) * 32] = f;
                j++;
            };
            if (dir > 0)
                for (int32_t k = g.losortStart[c]; k < g.losortStart[c + 1]; k++) put(g.losort[k], g.L[g.losort[k]]); // ascending lower column
            else
                for (int32_t f = g.ownerStart[c + 1] - 1; f >= g.ownerStart[c]; f--) put(f, g.U[f]); // DESCENDING upper column
        }
    }
    // ------------------------------------------------------------------ split streams (fast path)
  public:
    static constexpr int kMaxConst = 2; // cross-group values a row's remaining terms may need on the fast path
    static int p_rec_bytes(int Lg, int Kg) { return Lg * 384 + Kg * 128; }
    static int c_rec_bytes(int Rg) { return 256 + Rg * 256; }
    // The C-stream is PLANE-MAJOR inside a block of kSweepBlock steps: meta[steps][32] u64 | coef plane 0 [steps][32] |
    // coef plane 1 [steps][32] | ... - the consumer warp then reaches the operands of all the steps of a block
    // (plane 0 / plane 1 of a canonical step) at compile-time offsets from one base address, whatever Rg is.
    static int64_t c_meta_off(int Rg, int step)
    {
        return int64_t(step / kSweepBlock) * kSweepBlock * c_rec_bytes(Rg) + int64_t(step % kSweepBlock) * 256;
    }
    static int64_t c_coef_off(int Rg, int step, int r)
    {
        return int64_t(step / kSweepBlock) * kSweepBlock * c_rec_bytes(Rg) + int64_t(kSweepBlock) * 256 * (1 + r) +
               int64_t(step % kSweepBlock) * 256;
    }
    static int c_ring_step_bytes(int Rg, int Kg) { return 256 * (1 + Kg) + c_rec_bytes(Rg); } // hdr acc0[32], cval[Kg][32] + record
    static int hdr_step_bytes(int Kg) { return 256 * (1 + Kg); }                              // acc0[32] | cval[Kg][32]
    // one shared-memory stage of a fast group = one block of kSweepBlock steps:
    //   P-records | C-records | a[steps][32] | b[steps][32] (forward only) | hdr[steps] (written by the producer warps)
    static int fast_stage_bytes(int Lg, int Rg, int Kg, int nVec)
    {
        return kSweepBlock * (p_rec_bytes(Lg, Kg) + c_rec_bytes(Rg) + nVec * 256 + hdr_step_bytes(Kg));
    }

  private:
    void build_split(int dir, Dir& D)
    {
        const int nG = nGroups;
        const int nVec = dir > 0 ? 2 : 1;
        D.gFast.assign(nG, 0);
        D.gLg.assign(nG, 0);
        D.gRg.assign(nG, 0);
        D.gKg.assign(nG, 0);
        D.gPOff.assign(nG + 1, 0);
        D.gCOff.assign(nG + 1, 0);
        D.gPFaceOff.assign(nG + 1, 0);
        D.gCFaceOff.assign(nG + 1, 0);
        D.gGenOff.assign(nG + 1, 0);
        D.maxPStage = D.maxCStep = D.maxGenStage = D.maxFastStage = D.nFastGroups = 0;
        D.nGeneralSteps = 0;
        D.nLinkedGeneralSteps = 0;
        D.nDualSteps = 0;
        // pass 1: eligibility, Lg, Rg
        for (int gI = 0; gI < nG; gI++)
        {
            const int W = D.gW[gI], nT = gNT[gI];
            bool fast = W >= 1 && W <= 6 && D.gShflMask[gI] >= 0 && !forceGeneric;
            int Lg = 0, Rg = 0, Kg = 0;
            for (int step = 0; step < nT && fast; step++)
                for (int lane = 0; lane < 32; lane++)
                {
                    const int64_t b0 = D.gTermOff[gI] + int64_t(step) * W * 32 + lane;
                    int ld = 0, nTerms = 0, nConst = 0;
                    for (int j = 0; j < W; j++)
                        if (D.code[b0 + int64_t(j) * 32] != kCodeNone) nTerms = j + 1;
                    while (ld < nTerms && D.code[b0 + int64_t(ld) * 32] >= 0) ld++;
                    for (int j = ld; j < nTerms; j++)
                        if (D.code[b0 + int64_t(j) * 32] >= 0) nConst++;
                    if (nConst > kMaxConst) fast = false;
                    Kg = std::max(Kg, nConst);
                    Lg = std::max(Lg, ld);
                    Rg = std::max(Rg, nTerms - ld);
                }
            D.gFast[gI] = fast ? 1 : 0;
            D.gLg[gI] = fast ? Lg : 0;
            Rg = std::max(Rg, 2); // planes 0 / 1 hold the shuffled / own-lane coefficient of a canonical step
            D.gRg[gI] = fast ? Rg : 0;
            D.gKg[gI] = fast ? Kg : 0;
            if (fast)
            {
                D.nFastGroups++;
                D.gCH[gI] = kSweepBlock;
                D.gPOff[gI + 1] = D.gPOff[gI] + int64_t(nT) * p_rec_bytes(Lg, Kg);
                D.gCOff[gI + 1] = D.gCOff[gI] + int64_t(nT) * c_rec_bytes(Rg);
                D.gPFaceOff[gI + 1] = D.gPFaceOff[gI] + int64_t(nT) * Lg * 32;
                D.gCFaceOff[gI + 1] = D.gCFaceOff[gI] + int64_t(nT) * Rg * 32;
                D.gGenOff[gI + 1] = D.gGenOff[gI];
                D.maxPStage = std::max(D.maxPStage, kSweepBlock * (p_rec_bytes(Lg, Kg) + nVec * 256));
                D.maxCStep = std::max(D.maxCStep, c_ring_step_bytes(Rg, Kg));
                D.maxFastStage = std::max(D.maxFastStage, fast_stage_bytes(Lg, Rg, Kg, nVec));
            }
            else
            {
                D.gPOff[gI + 1] = D.gPOff[gI];
                D.gCOff[gI + 1] = D.gCOff[gI];
                D.gPFaceOff[gI + 1] = D.gPFaceOff[gI];
                D.gCFaceOff[gI + 1] = D.gCFaceOff[gI];
                D.gGenOff[gI + 1] = D.gGenOff[gI] + int64_t(nT) * W * 32;
                D.maxGenStage = std::max(D.maxGenStage, D.gCH[gI] * (W * 384 + nVec * 256));
            }
        }
        D.nGenTerms = D.gGenOff[nG];
        D.pStream.assign(size_t(D.gPOff[nG]), 0);
        D.cStream.assign(size_t(D.gCOff[nG]), 0);
        D.pFace.assign(size_t(D.gPFaceOff[nG]), -1);
        D.cFace.assign(size_t(D.gCFaceOff[nG]), -1);
        // pass 2: fill
        for (int gI = 0; gI < nG; gI++)
        {
            if (!D.gFast[gI]) continue;
            const int W = D.gW[gI], nT = gNT[gI], Lg = D.gLg[gI], Rg = D.gRg[gI], Kg = D.gKg[gI];
            const int pRec = p_rec_bytes(Lg, Kg);
            for (int step = 0; step < nT; step++)
            {
                unsigned char* pr = D.pStream.data() + D.gPOff[gI] + int64_t(step) * pRec;
                unsigned char* cr = D.cStream.data() + D.gCOff[gI] + c_meta_off(Rg, step);
                int32_t* pCode = reinterpret_cast<int32_t*>(pr + Lg * 256);
                int32_t* pConst = reinterpret_cast<int32_t*>(pr + Lg * 384);
                uint64_t* meta = reinterpret_cast<uint64_t*>(cr);
                // A step is CANONICAL if the remaining terms of every lane are, in reference order, an optional
                // term shuffled from the linked neighbour lane (lane - dir) followed by an optional own-lane term.
                // Its record then holds the shuffle coefficient in plane 0 and the own-lane coefficient in plane 1
                // (0 where the term is absent) and the consumer runs  acc = (acc0 - c0*shfl) - c1*own  without
                // looking at the descriptors.  Any other step keeps its terms in reference order in planes 0.. and is
                // flagged (meta byte 7 bit 0); the consumer then interprets the descriptor bytes.
                int Rt = 0;
                bool canonical = true;
                int ldOf[32], nOf[32];
                for (int lane = 0; lane < 32; lane++)
                {
                    const int64_t b0 = D.gTermOff[gI] + int64_t(step) * W * 32 + lane;
                    int ld = 0, nTerms = 0;
                    for (int j = 0; j < W; j++)
                        if (D.code[b0 + int64_t(j) * 32] != kCodeNone) nTerms = j + 1;
                    while (ld < nTerms && D.code[b0 + int64_t(ld) * 32] >= 0) ld++;
                    ldOf[lane] = ld;
                    nOf[lane] = nTerms;
                    Rt = std::max(Rt, nTerms - ld);
                    int j = ld;
                    if (j < nTerms && D.code[b0 + int64_t(j) * 32] == kCodeShfl - ((lane - dir) & 31) && lane - dir >= 0 && lane - dir < 32) j++;
                    if (j < nTerms && D.code[b0 + int64_t(j) * 32] == kCodeOwn) j++;
                    if (j != nTerms) canonical = false;
                }
                if (canonical) Rt = 2;
                // DUAL step: not canonical, but every lane is in canonical form or in seam form (block seams of a
                // multi-block mesh: the i-1 neighbour lies in the previous mesh block and has the lowest column, so
                // the own-lane term comes first: own, k-1, j-1).  The skew of the lanes spreads one seam over 62
                // steps, so these steps deserve a descriptor-free evaluation of their own.
                bool dual = false;
                int formOf[32], idxA[32], idxB[32];
                if (!canonical)
                {
                    dual = true;
                    for (int lane = 0; lane < 32 && dual; lane++)
                    {
                        const int64_t b0 = D.gTermOff[gI] + int64_t(step) * W * 32 + lane;
                        const int ld = ldOf[lane], nTerms = nOf[lane];
                        auto codeAt = [&](int j) { return D.code[b0 + int64_t(j) * 32]; };
                        auto linkedShfl = [&](int32_t c) { return lane - dir >= 0 && lane - dir < 32 && c == kCodeShfl - (lane - dir); };
                        formOf[lane] = 0;
                        idxA[lane] = idxB[lane] = -1;
                        int j = ld;
                        if (j < nTerms && linkedShfl(codeAt(j))) j++;
                        if (j < nTerms && codeAt(j) == kCodeOwn) j++;
                        if (j == nTerms) continue; // canonical form
                        j = ld;
                        if (!(j < nTerms && codeAt(j) == kCodeOwn))
                        {
                            dual = false;
                            break;
                        }
                        j++;
                        int t[2], nt = 0;
                        while (j < nTerms && nt < 2 && (codeAt(j) >= 0 || linkedShfl(codeAt(j)))) t[nt++] = j++;
                        if (j != nTerms || (nt == 2 && (codeAt(t[0]) < 0 || Rg < 3))) dual = false;
                        formOf[lane] = 1;
                        if (nt == 2)
                        {
                            idxA[lane] = t[0];
                            idxB[lane] = t[1];
                        }
                        else if (nt == 1)
                            idxB[lane] = t[0];
                    }
                }
                for (int lane = 0; lane < 32; lane++)
                {
                    const int64_t b0 = D.gTermOff[gI] + int64_t(step) * W * 32 + lane;
                    const int ld = ldOf[lane], nTerms = nOf[lane];
                    for (int i = 0; i < Lg; i++)
                    {
                        pCode[i * 32 + lane] = i < ld ? D.code[b0 + int64_t(i) * 32] : kCodeNone;
                        D.pFace[D.gPFaceOff[gI] + (int64_t(step) * Lg + i) * 32 + lane] = i < ld ? D.face[b0 + int64_t(i) * 32] : -1;
                    }
                    for (int k = 0; k < Kg; k++) pConst[k * 32 + lane] = -1;
                    int nc = 0;
                    uint64_t m = 0;
                    int32_t* cFaceRow = &D.cFace[D.gCFaceOff[gI] + int64_t(step) * Rg * 32 + lane];
                    auto describe = [&](int j) -> unsigned {
                        const int32_t code = D.code[b0 + int64_t(j) * 32];
                        if (code >= 0)
                        {
                            const unsigned byte = unsigned(lane) | kMetaConst | (nc ? 0x20u : 0u);
                            pConst[nc * 32 + lane] = code;
                            nc++;
                            return byte;
                        }
                        if (code == kCodeOwn) return unsigned(lane) | kMetaOwn;
                        return unsigned(kCodeShfl - code);
                    };
                    unsigned bytes[6];
                    for (int r = 0; r < 6; r++) bytes[r] = unsigned(lane) | kMetaPad; // padding: coefficient 0
                    uint64_t laneFlags = 0;
                    if (dual && formOf[lane] == 1)
                    { // seam form: own | A | B in reference order -> planes 1 | 2 | 0
                        bytes[1] = describe(ld);
                        cFaceRow[32] = D.face[b0 + int64_t(ld) * 32];
                        if (idxA[lane] >= 0)
                        {
                            bytes[2] = describe(idxA[lane]); // a cross-group value: becomes cval 0
                            cFaceRow[64] = D.face[b0 + int64_t(idxA[lane]) * 32];
                        }
                        uint64_t selB = 0;
                        if (idxB[lane] >= 0)
                        {
                            if (D.code[b0 + int64_t(idxB[lane]) * 32] >= 0) selB = nc ? 2 : 1;
                            bytes[0] = describe(idxB[lane]);
                            cFaceRow[0] = D.face[b0 + int64_t(idxB[lane]) * 32];
                        }
                        laneFlags = (uint64_t(1) << 58) | (selB << 59);
                    }
                    else if (canonical || dual)
                    {
                        int j = ld;
                        if (j < nTerms && D.code[b0 + int64_t(j) * 32] != kCodeOwn)
                        {
                            bytes[0] = describe(j);
                            cFaceRow[0] = D.face[b0 + int64_t(j) * 32];
                            j++;
                        }
                        if (j < nTerms)
                        {
                            bytes[1] = describe(j);
                            cFaceRow[32] = D.face[b0 + int64_t(j) * 32];
                        }
                    }
                    else
                        for (int r = 0; r < Rg && ld + r < nTerms; r++)
                        {
                            bytes[r] = describe(ld + r);
                            cFaceRow[int64_t(r) * 32] = D.face[b0 + int64_t(ld + r) * 32];
                        }
                    for (int r = 0; r < 6; r++) m |= uint64_t(bytes[r]) << (8 * r);
                    meta[lane] = m | laneFlags;
                }
                // bit 57: every shuffled term of this step reads the linked neighbour lane (lane - dir): the consumer
                // then issues one shuffle per step instead of one per term
                bool linked = true;
                for (int lane = 0; lane < 32; lane++)
                    for (int r = 0; r < 6; r++)
                    {
                        const unsigned byte = unsigned(meta[lane] >> (8 * r)) & 0xffu;
                        if (!(byte & (kMetaConst | kMetaOwn | kMetaPad)) && (lane - dir < 0 || lane - dir > 31 || int(byte & kMetaLane) != lane - dir))
                            linked = false;
                    }
                for (int lane = 0; lane < 32; lane++)
                    meta[lane] |= (uint64_t(Rt) << 48) | (uint64_t(canonical || dual ? 0 : 1) << 56) | (uint64_t(linked ? 1 : 0) << 57) |
                                  (uint64_t(dual ? 1 : 0) << 61);
                if (!canonical) D.nGeneralSteps++;
                if (!canonical && linked) D.nLinkedGeneralSteps++;
                if (dual) D.nDualSteps++;
            }
        }
    }

  public:
};


// ------------------------------------------------------------------------------------------
// Row-packed Amul layout over slots: slice s = the 32 slots of one (group, time step);
// column-major inside a slice (entry = sliceOff[s]*32 + j*32 + lane); columns are slots.
struct SellLayout
{
    int64_t nSlices = 0, nEntries = 0;
    std::vector<int32_t> sliceOff; // [nSlices+1] in units of 32 entries
    std::vector<int32_t> col;      // [nEntries] column slot or -1 (padding)
    std::vector<int32_t> src;      // [nEntries] index into coef = [upper(F) | lower(F)], -1 for padding

    void build(const GlobalLdu& g, const PipeSchedule& S)
    {
        const int64_t F = g.F;
        nSlices = S.nSlots / 32;
        sliceOff.assign(nSlices + 1, 0);
        std::vector<int32_t> width(nSlices, 0);
        for (int64_t c = 0; c < g.N; c++)
        {
            const int cnt = (g.losortStart[c + 1] - g.losortStart[c]) + (g.ownerStart[c + 1] - g.ownerStart[c]);
            int32_t& w = width[S.slotOfCell[c] >> 5];
            w = std::max(w, cnt);
        }
        for (int64_t s = 0; s < nSlices; s++)
        {
            const int64_t next = int64_t(sliceOff[s]) + width[s];
            if (next >= (int64_t(1) << 31) / 32) throw std::runtime_error("SELL layout exceeds int32 entries");
            sliceOff[s + 1] = int32_t(next);
        }
        nEntries = int64_t(sliceOff[nSlices]) * 32;
        col.assign(nEntries, -1);
        src.assign(nEntries, -1);
        for (int64_t c = 0; c < g.N; c++)
        {
            const int32_t sl = S.slotOfCell[c];
            const int64_t base = int64_t(sliceOff[sl >> 5]) * 32 + (sl & 31);
            int j = 0;
            for (int32_t k = g.losortStart[c]; k < g.losortStart[c + 1]; k++, j++)
            {
                const int32_t f = g.losort[k]; // row c = U[f], column L[f], coefficient lower[f]
                col[base + int64_t(j) * 32] = S.slotOfCell[g.L[f]];
                src[base + int64_t(j) * 32] = int32_t(F + f);
            }
            for (int32_t f = g.ownerStart[c]; f < g.ownerStart[c + 1]; f++, j++)
            {
                col[base + int64_t(j) * 32] = S.slotOfCell[g.U[f]]; // row c = L[f], column U[f], coefficient upper[f]
                src[base + int64_t(j) * 32] = f;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// Interface plan: rows touched by coupled patches, their ordered entries, and the halo layout.
// All row / source indices are SLOTS.
struct IfacePlan
{
    std::vector<int32_t> rows;     // slots of the touched rows (ascending cell order)
    std::vector<int32_t> rowStart; // [rows+1]
    std::vector<int32_t> entCoef;  // index into the concatenated interface coefficient array
    std::vector<int32_t> entSrc;   // cnt==0: source code (>=0: x slot, <0: recv index -1-k); else start into g*
    std::vector<int32_t> entCnt;   // 0 = identity
    std::vector<int32_t> gSrc;     // source codes of GGI terms
    std::vector<double> gW;
    std::vector<uint32_t> sliceMask; // per SELL slice: lanes whose row is touched
    // halo
    std::vector<int> peers;
    std::vector<int32_t> sendOff, recvOff; // [peers+1]
    std::vector<int32_t> sendCells;        // x slots packed into the send buffer
    int64_t nCoefs = 0;

    void build(std::vector<RegionHost>& regs, int myRank, const PipeSchedule& S) { build(regs, myRank, S.slotOfCell, S.nSlots); }
    // (the plan needs the slot of every cell only: it can be rebuilt after finalize, when the GGI interpolation of an
    // interface is replaced, without the sweep schedule)
    struct SlotView
    {
        const std::vector<int32_t>& slotOfCell;
        int64_t nSlots;
    };
    void build(std::vector<RegionHost>& regs, int myRank, const std::vector<int32_t>& slotOfCell_, int64_t nSlots_)
    {
        const SlotView S{slotOfCell_, nSlots_};
        nCoefs = 0;
        for (auto& r : regs)
            for (auto& I : r.ifaces)
            {
                I.coefOffset = nCoefs;
                nCoefs += I.nFaces;
            }
        // sources of an interface: (rank, region, iface) patches whose patch-internal values it reads, n values each
        struct Src
        {
            int rank, region, iface;
            int32_t n;
        };
        auto sourcesOf = [&](const IfaceHost& I) {
            std::vector<Src> out;
            if (I.pieces.empty())
                out.push_back({I.peerRank, I.peerRegion, I.peerIface, int32_t(I.nPeerFaces)});
            else
                for (const auto& P : I.pieces) out.push_back({P.rank, P.region, P.iface, int32_t(P.zoneAddr.size())});
            return out;
        };
        for (auto& r : regs)
            for (auto& I : r.ifaces)
                for (const Src& q : sourcesOf(I))
                    if (q.rank != myRank && std::find(peers.begin(), peers.end(), q.rank) == peers.end()) peers.push_back(q.rank);
        std::sort(peers.begin(), peers.end());
        sendOff.assign(peers.size() + 1, 0);
        recvOff.assign(peers.size() + 1, 0);
        // offset of what a remote patch (rank, region, iface) sends, inside the receive buffer
        std::map<std::tuple<int, int, int>, int32_t> recvSegOff;
        for (size_t p = 0; p < peers.size(); p++)
        {
            // send layout: my interfaces that rank peers[p] reads, in (region, iface) order, the whole patch each.  A patch is
            // read by the ranks it reads from (processor patch: the neighbour; zone piece: the ranks holding pieces of the
            // shadow zone, which list this piece among THEIR pieces)
            int32_t so = sendOff[p];
            for (size_t r = 0; r < regs.size(); r++)
                for (auto& I : regs[r].ifaces)
                {
                    bool reads = false;
                    for (const Src& q : sourcesOf(I)) reads = reads || q.rank == peers[p];
                    if (!reads) continue;
                    for (int i = 0; i < I.nFaces; i++) sendCells.push_back(S.slotOfCell[regs[r].cellOffset + I.faceCells[i]]);
                    so += I.nFaces;
                }
            sendOff[p + 1] = so;
            // receive layout: the peer's patches I read, in ITS (region, iface) order
            std::map<std::pair<int, int>, int32_t> theirs;
            for (auto& r : regs)
                for (auto& I : r.ifaces)
                    for (const Src& q : sourcesOf(I))
                        if (q.rank == peers[p])
                        {
                            auto it = theirs.find({q.region, q.iface});
                            if (it != theirs.end() && it->second != q.n)
                                throw std::runtime_error("two interfaces disagree about the size of a remote patch");
                            theirs[{q.region, q.iface}] = q.n;
                        }
            int32_t ro = recvOff[p];
            for (auto& kv : theirs)
            {
                recvSegOff[std::make_tuple(peers[p], kv.first.first, kv.first.second)] = ro;
                ro += kv.second;
            }
            recvOff[p + 1] = ro;
        }
        // zone face -> (piece, position) of the interfaces whose shadow is spread over pieces
        std::vector<std::vector<std::vector<int32_t>>> zonePiece(regs.size()), zonePos(regs.size());
        for (size_t r = 0; r < regs.size(); r++)
        {
            zonePiece[r].resize(regs[r].ifaces.size());
            zonePos[r].resize(regs[r].ifaces.size());
            for (size_t i = 0; i < regs[r].ifaces.size(); i++)
            {
                const IfaceHost& I = regs[r].ifaces[i];
                if (I.pieces.empty()) continue;
                zonePiece[r][i].assign((size_t)I.nPeerFaces, -1);
                zonePos[r][i].assign((size_t)I.nPeerFaces, -1);
                for (size_t k = 0; k < I.pieces.size(); k++)
                    for (size_t q = 0; q < I.pieces[k].zoneAddr.size(); q++)
                    {
                        const int32_t z = I.pieces[k].zoneAddr[q];
                        if (z < 0 || z >= I.nPeerFaces) throw std::runtime_error("zone address of a shadow piece out of range");
                        if (zonePiece[r][i][z] >= 0) throw std::runtime_error("two shadow pieces hold the same zone face");
                        zonePiece[r][i][z] = int32_t(k);
                        zonePos[r][i][z] = int32_t(q);
                    }
            }
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
        sliceMask.assign(S.nSlots / 32, 0u);
        rowStart.push_back(0);
        int32_t lastRow = -1;
        for (size_t e = 0; e < ents.size(); e++)
        {
            const Ent& E = ents[e];
            if (rows.empty() || lastRow != E.row)
            {
                if (!rows.empty()) rowStart.push_back(int32_t(entCoef.size()));
                const int32_t sl = S.slotOfCell[E.row];
                rows.push_back(sl);
                lastRow = E.row;
                sliceMask[sl >> 5] |= (1u << (sl & 31));
            }
            const IfaceHost& I = regs[E.region].ifaces[E.iface];
            entCoef.push_back(int32_t(I.coefOffset + E.face));
            auto srcCode = [&](int peerFace) -> int32_t {
                int rank = I.peerRank, region = I.peerRegion, iface = I.peerIface, pos = peerFace;
                if (peerFace < 0 || peerFace >= I.nPeerFaces) throw std::runtime_error("GGI address out of range of the shadow patch");
                if (!I.pieces.empty())
                {
                    const int32_t k = zonePiece[E.region][E.iface][peerFace];
                    if (k < 0) throw std::runtime_error("GGI address names a zone face that no shadow piece holds");
                    rank = I.pieces[k].rank;
                    region = I.pieces[k].region;
                    iface = I.pieces[k].iface;
                    pos = zonePos[E.region][E.iface][peerFace];
                }
                if (rank == myRank)
                {
                    if (region < 0 || region >= int(regs.size()) || iface < 0 || iface >= int(regs[region].ifaces.size()))
                        throw std::runtime_error("interface peer (region, iface) does not exist on this rank");
                    const IfaceHost& Q = regs[region].ifaces[iface];
                    if (pos < 0 || pos >= Q.nFaces) throw std::runtime_error("GGI address out of range of the shadow patch");
                    return S.slotOfCell[regs[region].cellOffset + Q.faceCells[pos]];
                }
                return -1 - (recvSegOff.at(std::make_tuple(rank, region, iface)) + pos);
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
                if (cnt == 0) entCnt.back() = -1; // uncovered face: contributes coeff*0 (bridgeOverlap off in every shipped case)
            }
        }
        if (!rows.empty()) rowStart.push_back(int32_t(entCoef.size()));
    }
};

} // namespace b200
