"""Processor decomposition of a serial coupled LDU system (decomposePar semantics).

Every region is decomposed separately into the same number of sub-domains; rank g owns
sub-domain g of every region (/root/reference/tutorials/conjugateHeatTransfer/
flowOverHeatedPlate/Allrun:97-100, system/*/decomposeParDict).  Inside a sub-domain cells and
internal faces keep their relative order (so the local LDU addressing is still
upper-triangular and the local DIC/DILU is foam-extend's block-Jacobi one); faces cut by
the decomposition become processor patches, appended after the physical (regionCouple)
patches, ordered by neighbour rank, faces in global face order on both sides.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .case import Case, Interface, PROCESSOR, RankSystem, REGION_COUPLE, Region


def simple_cell_to_rank(coords: Sequence[np.ndarray], n: Sequence[int]) -> np.ndarray:
    """``method simple; n (nx ny nz)``: split along each direction into equal-count slabs of the
    cells sorted by that coordinate (simpleGeomDecomp)."""
    N = coords[0].size
    rank = np.zeros(N, dtype=np.int32)
    mult = 1
    for xyz, parts in zip(coords, n):
        if parts > 1:
            order = np.argsort(xyz, kind="stable")
            slab = np.empty(N, dtype=np.int32)
            slab[order] = (np.arange(N, dtype=np.int64) * parts // N).astype(np.int32)
            rank += slab * mult
        mult *= parts
    return rank


def decompose(case: Case, cellToRank: List[np.ndarray], nRanks: int) -> Case:
    if case.nRanks != 1:
        raise ValueError("decompose expects a serial case")
    serial = case.ranks[0].regions
    nReg = len(serial)
    # local numbering
    localIdx = []
    cellsOf: List[List[np.ndarray]] = []
    for ri, reg in enumerate(serial):
        c2r = cellToRank[ri]
        li = np.empty(reg.nCells, dtype=np.int32)
        per = []
        for g in range(nRanks):
            cells = np.nonzero(c2r == g)[0].astype(np.int32)
            li[cells] = np.arange(cells.size, dtype=np.int32)
            per.append(cells)
        localIdx.append(li)
        cellsOf.append(per)

    ranks = [RankSystem(g, nRanks, []) for g in range(nRanks)]
    # per (region) : dict[(a,b)] -> face indices cut between rank a (owner side l) and b
    procFaces: List[Dict[Tuple[int, int], np.ndarray]] = []
    for ri, reg in enumerate(serial):
        c2r = cellToRank[ri]
        rl, ru = c2r[reg.lowerAddr], c2r[reg.upperAddr]
        lo = reg.upper if reg.lower is None else reg.lower
        cut = np.nonzero(rl != ru)[0]
        pf: Dict[Tuple[int, int], np.ndarray] = {}
        if cut.size:
            key = rl[cut].astype(np.int64) * nRanks + ru[cut]
            for kv in np.unique(key):
                pf[(int(kv // nRanks), int(kv % nRanks))] = cut[key == kv]
        procFaces.append(pf)
        for g in range(nRanks):
            cells = cellsOf[ri][g]
            inner = np.nonzero((rl == g) & (ru == g))[0]
            sub = Region(
                reg.name, int(cells.size),
                localIdx[ri][reg.lowerAddr[inner]], localIdx[ri][reg.upperAddr[inner]],
                reg.diag[cells].copy(), reg.upper[inner].copy(),
                None if reg.lower is None else reg.lower[inner].copy(),
                reg.source[cells].copy(), reg.psi[cells].copy(), [], cells)
            ranks[g].regions.append(sub)
        del lo

    # regionCouple pieces
    for ri, reg in enumerate(serial):
        for ii, itf in enumerate(reg.interfaces):
            if itf.kind != REGION_COUPLE:
                raise ValueError("serial case may only hold regionCouple interfaces")
            peer = serial[itf.peerRegion].interfaces[itf.peerIface]
            myRank = cellToRank[ri][itf.faceCells]
            peerRank = cellToRank[itf.peerRegion][peer.faceCells]
            if itf.ggiOffsets is None and np.array_equal(myRank, peerRank):
                # conformal pair cut consistently: every face pair stays on one rank (regionCouplePolyPatch::localParallel())
                for g in range(nRanks):
                    sel = np.nonzero(myRank == g)[0]
                    ranks[g].regions[ri].interfaces.append(Interface(
                        REGION_COUPLE, localIdx[ri][itf.faceCells[sel]], itf.bouCoeffs[sel].copy(),
                        itf.intCoeffs[sel].copy(), g, itf.peerRegion, itf.peerIface, name=itf.name))
                continue
            # general case (non-conformal GGI, or the two regions cut at different places: the shipped n (2 1 2)): the pair
            # is interpolated on the global zones.  The zone of a patch = its serial face list; rank g holds the faces
            # sel_g (its zoneAddressing); the GGI rows of those faces address shadow ZONE faces; the shadow zone is held in
            # pieces by the ranks with shadow faces.
            nZone = int(peer.faceCells.size)
            if itf.ggiOffsets is None:
                go, ga, gw = np.arange(itf.nFaces + 1, dtype=np.int32), np.arange(itf.nFaces, dtype=np.int32), np.ones(itf.nFaces)
            else:
                go, ga, gw = itf.ggiOffsets, itf.ggiAddr, itf.ggiWeights
            pieces = [(h, itf.peerRegion, itf.peerIface, np.nonzero(peerRank == h)[0].astype(np.int32)) for h in range(nRanks)]
            pieces = [p for p in pieces if p[3].size]
            for g in range(nRanks):
                sel = np.nonzero(myRank == g)[0]
                cnt = (go[sel + 1] - go[sel]).astype(np.int64)
                off = np.zeros(sel.size + 1, np.int32)
                off[1:] = np.cumsum(cnt)
                idx = (np.repeat(go[sel].astype(np.int64) - off[:-1], cnt) + np.arange(int(off[-1]))) if sel.size else np.zeros(0, np.int64)
                ranks[g].regions[ri].interfaces.append(Interface(
                    REGION_COUPLE, localIdx[ri][itf.faceCells[sel]], itf.bouCoeffs[sel].copy(), itf.intCoeffs[sel].copy(),
                    g, itf.peerRegion, itf.peerIface, off, np.asarray(ga)[idx].astype(np.int32), np.asarray(gw)[idx].copy(),
                    name=itf.name, nPeerFaces=nZone, pieces=list(pieces) if sel.size else []))

    # processor patches, appended last, ordered by neighbour rank
    for ri, reg in enumerate(serial):
        lo = reg.upper if reg.lower is None else reg.lower
        pf = procFaces[ri]
        nbrs: List[List[int]] = [[] for _ in range(nRanks)]
        for (a, b) in pf:
            nbrs[a].append(b)
            nbrs[b].append(a)
        ifaceIdx: Dict[Tuple[int, int], int] = {}
        for g in range(nRanks):
            base = len(ranks[g].regions[ri].interfaces)
            for k, nb in enumerate(sorted(set(nbrs[g]))):
                ifaceIdx[(g, nb)] = base + k
        for g in range(nRanks):
            for nb in sorted(set(nbrs[g])):
                # faces where g is on the lower side and nb on the upper side, and vice versa,
                # merged in global face order (identical sequence on both ranks)
                fa = pf.get((g, nb), np.empty(0, dtype=np.int64))
                fb = pf.get((nb, g), np.empty(0, dtype=np.int64))
                faces = np.concatenate([fa, fb])
                mineIsLower = np.concatenate([np.ones(fa.size, bool), np.zeros(fb.size, bool)])
                order = np.argsort(faces, kind="stable")
                faces, mineIsLower = faces[order], mineIsLower[order]
                myCell = np.where(mineIsLower, reg.lowerAddr[faces], reg.upperAddr[faces])
                # row l: coefficient of psi_u is upper[f]; row u: coefficient of psi_l is lower[f]
                aCoef = np.where(mineIsLower, reg.upper[faces], lo[faces])
                tCoef = np.where(mineIsLower, lo[faces], reg.upper[faces])
                ranks[g].regions[ri].interfaces.append(Interface(
                    PROCESSOR, localIdx[ri][myCell], -aCoef, -tCoef, nb, ri, ifaceIdx[(nb, g)],
                    name=f"procBoundary{g}to{nb}"))
    return Case(f"{case.name}_np{nRanks}", ranks)


def decompose_cht_zslabs(case: Case, fluid, solid, nRanks: int) -> Case:
    """z-slab decomposition (``simple; n (1 1 nRanks)``) of the CHT case: both regions are cut at the
    same layers, so every regionCouple face pair stays on one rank."""
    maps = []
    for reg in (fluid, solid):
        k = reg.cell_ijk_layer().astype(np.int64)
        maps.append((k * nRanks // reg.nz).astype(np.int32))
    return decompose(case, maps, nRanks)


def decompose_cht_simple(case: Case, fluid, solid, n: Sequence[int]) -> Case:
    """``method simple; n (nx ny nz)`` applied to each region on its own, as decomposePar does it
    (tutorials/conjugateHeatTransfer/flowOverHeatedPlate/system/{fluid,solid}/decomposeParDict:17-27, n (2 1 2)): the
    fluid (x in [-0.5, 3]) and the plate (x in [0, 1]) are cut at different x, so pieces of the regionCouple pair face
    other ranks (Interface.pieces)."""
    maps = [simple_cell_to_rank(reg.cell_xyz(), n) for reg in (fluid, solid)]
    return decompose(case, maps, int(np.prod(n)))


def flatten_ranks(case: Case) -> Case:
    """All sub-domains of a decomposed case as the regions of ONE rank (row = rank * nRegions + region, the oracle's row
    order): the same coupled system, solvable on one device; processor patches become interfaces between regions."""
    import copy
    nReg = case.nRegions
    regions = []
    for rk in case.ranks:
        for reg in rk.regions:
            q = copy.copy(reg)
            q.interfaces = []
            for itf in reg.interfaces:
                j = copy.copy(itf)
                j.peerRegion = itf.peerRank * nReg + itf.peerRegion
                j.peerRank = 0
                if getattr(itf, "pieces", None):
                    j.pieces = [(0, h * nReg + pr, pi, za) for (h, pr, pi, za) in itf.pieces]
                q.interfaces.append(j)
            regions.append(q)
    return Case(case.name + "_flat", [RankSystem(0, 1, regions)])
