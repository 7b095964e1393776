"""Host-side description of a (decomposed) coupled LDU system.

Vocabulary follows foam-extend / the reference: a *region* owns one ``lduMatrix``
(``lowerAddr``/``upperAddr``/``diag``/``upper``/``lower``), its coupled patches are
*interfaces* with ``faceCells``, ``boundaryCoeffs`` and ``internalCoeffs``; a monolithic
solve couples several regions (``coupledFvMatrix`` at
/root/reference/src/multiRegionSystem/multiRegionSystem.C:85,143-153); a parallel run holds
one sub-domain of every region per rank, joined by processor interfaces.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

REGION_COUPLE = 0  # regionCouple / ggi patch (non-processor coupled interface)
PROCESSOR = 1      # processor patch


@dataclass
class Interface:
    kind: int
    faceCells: np.ndarray            # int32 [P]
    bouCoeffs: np.ndarray            # float64 [P]  boundaryCoeffs (result[fc] -= bou*pnf)
    intCoeffs: np.ndarray            # float64 [P]  internalCoeffs (already inside diag; used by Tmul)
    peerRank: int
    peerRegion: int
    peerIface: int
    # GGI CSR mapping the peer patch's face values onto this patch's faces; None = identity
    ggiOffsets: Optional[np.ndarray] = None
    ggiAddr: Optional[np.ndarray] = None
    ggiWeights: Optional[np.ndarray] = None
    name: str = ""
    nPeerFaces: Optional[int] = None  # faces of the shadow patch (None: same as this patch)
    # shadow patch spread over ranks by the decomposition: nPeerFaces = faces of the shadow ZONE, ggiAddr = zone face
    # labels, pieces = [(rank, region, iface, zoneAddr of that piece's faces)]; peerRank/peerRegion/peerIface unused
    pieces: Optional[list] = None

    @property
    def nFaces(self) -> int:
        return int(self.faceCells.size)


@dataclass
class Region:
    name: str
    nCells: int
    lowerAddr: np.ndarray            # int32 [F]
    upperAddr: np.ndarray            # int32 [F]
    diag: np.ndarray                 # float64 [N]
    upper: np.ndarray                # float64 [F]
    lower: Optional[np.ndarray]      # float64 [F] or None (symmetric)
    source: np.ndarray               # float64 [N]
    psi: np.ndarray                  # float64 [N] initial guess / solution
    interfaces: List[Interface] = field(default_factory=list)
    # for decomposed regions: global cell index of every local cell (None when undecomposed)
    globalCells: Optional[np.ndarray] = None

    @property
    def nFaces(self) -> int:
        return int(self.lowerAddr.size)

    @property
    def symmetric(self) -> bool:
        return self.lower is None


@dataclass
class RankSystem:
    rank: int
    nRanks: int
    regions: List[Region]

    @property
    def nCells(self) -> int:
        return sum(r.nCells for r in self.regions)

    @property
    def nFaces(self) -> int:
        return sum(r.nFaces for r in self.regions)


@dataclass
class Case:
    """A coupled system on nRanks ranks (nRanks == 1: serial)."""
    name: str
    ranks: List[RankSystem]

    @property
    def nRanks(self) -> int:
        return len(self.ranks)

    @property
    def nRegions(self) -> int:
        return len(self.ranks[0].regions)

    @property
    def nCells(self) -> int:
        return sum(r.nCells for r in self.ranks)

    @property
    def nFaces(self) -> int:
        return sum(r.nFaces for r in self.ranks)

    def concat(self, what: str) -> np.ndarray:
        """Concatenate a per-region vector ('psi' or 'source') in (rank, region) row order."""
        return np.concatenate([getattr(reg, what) for rk in self.ranks for reg in rk.regions])

    def row_offsets(self) -> np.ndarray:
        n = [reg.nCells for rk in self.ranks for reg in rk.regions]
        return np.concatenate([[0], np.cumsum(n)]).astype(np.int64)

    def to_global(self, vec: np.ndarray, nGlobalPerRegion: List[int]) -> List[np.ndarray]:
        """Scatter a concatenated (rank, region) vector back to per-region global vectors."""
        out = [np.zeros(n) for n in nGlobalPerRegion]
        off = 0
        for rk in self.ranks:
            for ri, reg in enumerate(rk.regions):
                seg = vec[off:off + reg.nCells]
                if reg.globalCells is None:
                    out[ri][:] = seg
                else:
                    out[ri][reg.globalCells] = seg
                off += reg.nCells
        return out


def algorithmic_bytes_per_iteration(nCells: int, nFaces: int, solver: str = "BiCGStab") -> int:
    """compulsory-traffic model, asymmetric storage.
    Amul A = 24N+24F; DILU precondition P = 48N+32F; BiCGStab vector work V = 160N
    => 2A+2P+V = 304N+112F.  PCG+DIC: 184N+48F."""
    if solver in ("BiCGStab", "PBiCGStab"):
        return 304 * nCells + 112 * nFaces
    if solver in ("PCG", "CG"):
        return 184 * nCells + 48 * nFaces
    raise ValueError(solver)
