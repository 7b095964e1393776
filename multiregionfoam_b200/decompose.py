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

    # regionCouple pieces (must pair rank-locally; see DESIGN.md multi-GPU section)
    for ri, reg in enumerate(serial):
        for ii, itf in enumerate(reg.interfaces):
            if itf.kind != REGION_COUPLE:
                raise ValueError("serial case may only hold regionCouple interfaces")
            peer = serial[itf.peerRegion].interfaces[itf.peerIface]
            myRank = cellToRank[ri][itf.faceCells]
            peerRank = cellToRank[itf.peerRegion][peer.faceCells]
            if itf.ggiOffsets is not None:
                raise NotImplementedError("decomposition of non-conformal GGI interfaces")
            if not np.array_equal(myRank, peerRank):
                raise NotImplementedError(
                    "regionCouple face pairs must live on the same rank (choose a decomposition that "
                    "cuts both regions consistently, e.g. z-slabs)")
            for g in range(nRanks):
                sel = np.nonzero(myRank == g)[0]
                ranks[g].regions[ri].interfaces.append(Interface(
                    REGION_COUPLE, localIdx[ri][itf.faceCells[sel]], itf.bouCoeffs[sel].copy(),
                    itf.intCoeffs[sel].copy(), g, itf.peerRegion, itf.peerIface, name=itf.name))

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
