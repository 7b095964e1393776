"""General 2-D multi-block hex topologies (extruded in z) in OpenFOAM LDU form: what ``blockMesh`` produces for a
``blockMeshDict`` whose blocks are quadrilaterals in the x-y plane with arbitrary vertex connectivity (O-grids
around a cylinder, ...), refined by an integer factor and extruded into ``layers`` z-layers.

Cells are numbered block by block in dictionary order, i fastest, then j, then k (blockMesh's order); internal faces
are all adjacent cell pairs (min, max) sorted lexicographically (OpenFOAM's upper-triangular order).
Two blocks are neighbours when they share an edge of the 2-D vertex graph; the cells along the shared edge pair up in
the direction given by the vertex labels, so blocks may meet with any relative orientation.

The one topology shipped here is HronTurekFsi3 (BASELINE config 4):
``/root/reference/tutorials/fluidStructureInteraction/HronTurekFsi3/constant/fluid/polyMesh/blockMeshDict:105-128`` (24 blocks,
5 336 cells at refinement 1) and ``.../solid/polyMesh/blockMeshDict:34`` (one block ``(105 6 1)``), with the
``interfaceShadow`` / ``interface`` patch pair of ``:162-169`` / solid ``:47-52`` along the elastic flag.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np


@dataclass
class Block2D:
    v: Tuple[int, int, int, int]  # vertices of the z = 0 face: (i0 j0), (i1 j0), (i1 j1), (i0 j1)
    nx: int
    ny: int


# local edges of a block: name -> (vertex index a, vertex index b, axis the edge runs along, fixed index side)
_EDGES = {
    "bottom": (0, 1, "i", 0),   # j = 0, runs along i from v0 to v1
    "right": (1, 2, "j", 1),    # i = nx-1, runs along j from v1 to v2
    "top": (3, 2, "i", 1),      # j = ny-1, runs along i from v3 to v2
    "left": (0, 3, "j", 0),     # i = 0, runs along j from v0 to v3
}


@dataclass
class MultiBlock2D:
    name: str
    blocks: List[Block2D]
    r: int = 1
    layers: int = 1

    nCells: int = 0
    lowerAddr: np.ndarray = field(default=None, repr=False)
    upperAddr: np.ndarray = field(default=None, repr=False)
    offsets: List[int] = field(default_factory=list)

    def dims(self, b: int) -> Tuple[int, int, int]:
        return self.blocks[b].nx * self.r, self.blocks[b].ny * self.r, self.layers

    def cells(self, b: int) -> np.ndarray:
        nx, ny, nz = self.dims(b)
        return self.offsets[b] + np.arange(nx * ny * nz, dtype=np.int64).reshape(nz, ny, nx)  # [k, j, i]

    def edge_cells(self, b: int, edge: str) -> np.ndarray:
        """Cells of block b along one of its edges, [k, position along the edge from vertex a to vertex b]."""
        c = self.cells(b)
        _, _, axis, side = _EDGES[edge]
        if axis == "i":
            return c[:, -1 if side else 0, :]
        return c[:, :, -1 if side else 0]

    def find_edge(self, va: int, vb: int) -> Tuple[int, str, bool]:
        """(block, edge name, reversed) of the block edge between 2-D vertices va -> vb."""
        for b, blk in enumerate(self.blocks):
            for name, (a, c, _, _) in _EDGES.items():
                if (blk.v[a], blk.v[c]) == (va, vb):
                    return b, name, False
                if (blk.v[a], blk.v[c]) == (vb, va):
                    return b, name, True
        raise KeyError((va, vb))

    def patch_cells(self, edges: Sequence[Tuple[int, int]]) -> np.ndarray:
        """faceCells of a patch given as a list of 2-D vertex pairs (va, vb) (the blockMeshDict's patch faces, in order):
        per face block the cells [k, position from va to vb], flattened position-fastest, then k - blockMesh numbers
        the faces of a block face that way."""
        out = []
        for va, vb in edges:
            b, name, rev = self.find_edge(va, vb)
            c = self.edge_cells(b, name)
            out.append((c[:, ::-1] if rev else c).reshape(-1))
        return np.concatenate(out).astype(np.int32)

    def build(self) -> "MultiBlock2D":
        self.offsets = []
        n = 0
        for b in range(len(self.blocks)):
            self.offsets.append(n)
            nx, ny, nz = self.dims(b)
            n += nx * ny * nz
        self.nCells = n
        lo: List[np.ndarray] = []
        up: List[np.ndarray] = []

        def add(a: np.ndarray, c: np.ndarray):
            a, c = a.reshape(-1), c.reshape(-1)
            lo.append(np.minimum(a, c))
            up.append(np.maximum(a, c))

        for b in range(len(self.blocks)):
            c = self.cells(b)
            add(c[:, :, :-1], c[:, :, 1:])
            add(c[:, :-1, :], c[:, 1:, :])
            if self.layers > 1:
                add(c[:-1], c[1:])
        # block-to-block: edges of the vertex graph shared by two blocks
        seen: Dict[frozenset, Tuple[int, str]] = {}
        for b, blk in enumerate(self.blocks):
            for name, (a, c, _, _) in _EDGES.items():
                key = frozenset((blk.v[a], blk.v[c]))
                if key in seen:
                    b2, name2 = seen[key]
                    a2, c2, _, _ = _EDGES[name2]
                    e1, e2 = self.edge_cells(b, name), self.edge_cells(b2, name2)
                    if e1.shape != e2.shape:
                        raise ValueError(f"blocks {b2} and {b} do not match along {sorted(key)}: {e2.shape[1]} against {e1.shape[1]} cells")
                    same = (blk.v[a], blk.v[c]) == (self.blocks[b2].v[a2], self.blocks[b2].v[c2])
                    add(e1, e2 if same else e2[:, ::-1])
                else:
                    seen[key] = (b, name)
        l = np.concatenate(lo)
        u = np.concatenate(up)
        order = np.lexsort((u, l))
        self.lowerAddr = l[order].astype(np.int32)
        self.upperAddr = u[order].astype(np.int32)
        return self


# HronTurekFsi3: the 24 fluid blocks (vertices of the z = 0 face, (nx ny)) and the flag
_HT_FLUID = [
    ((0, 1, 6, 5), 12, 12), ((1, 2, 7, 6), 15, 12), ((2, 3, 8, 7), 30, 12), ((3, 64, 65, 8), 35, 12), ((64, 4, 9, 65), 25, 12),
    ((5, 6, 23, 22), 12, 15), ((6, 10, 16, 23), 11, 15), ((6, 7, 11, 10), 15, 11), ((7, 13, 12, 11), 8, 11), ((7, 8, 14, 13), 30, 8),
    ((8, 65, 66, 14), 35, 8), ((65, 9, 15, 66), 25, 8), ((16, 17, 24, 23), 15, 11), ((17, 18, 19, 24), 8, 11), ((19, 20, 25, 24), 30, 8),
    ((20, 67, 68, 25), 35, 8), ((67, 21, 26, 68), 25, 8), ((22, 23, 28, 27), 12, 13), ((23, 24, 29, 28), 15, 13), ((24, 25, 30, 29), 30, 13),
    ((25, 68, 69, 30), 35, 13), ((68, 26, 31, 69), 25, 13), ((14, 66, 67, 20), 35, 2), ((66, 15, 21, 67), 25, 2),
]
# patch interfaceShadow of the fluid (blockMeshDict:162-169) as 2-D vertex pairs, and the x (or y) coordinates [mm] of the
# straight pieces of the flag they lie on: top y = 210 from the cylinder (x = 248.9898) to the tip (x = 600), the tip
# x = 600 from y = 210 to 190, the bottom back to the cylinder
HT_FLUID_INTERFACE = [(19, 18), (20, 19), (20, 14), (13, 14), (12, 13)]
HT_SOLID_INTERFACE = [(3, 2), (2, 1), (1, 0)]   # solid vertices 3=18, 2=20, 1=14, 0=12 (solid blockMeshDict:47-52)


def hron_turek(r: int = 1, layers: int = 1) -> Tuple[MultiBlock2D, MultiBlock2D]:
    fluid = MultiBlock2D("fluid", [Block2D(v, nx, ny) for v, nx, ny in _HT_FLUID], r, layers).build()
    solid = MultiBlock2D("solid", [Block2D((0, 1, 2, 3), 105, 6)], r, layers).build()
    return fluid, solid


def hron_turek_interface_intervals(fluid: MultiBlock2D, solid: MultiBlock2D):
    """The faces of the two interface patches as intervals [s0, s1] x [layer k] of one arc-length coordinate s [mm] that
    runs along the flag: top edge from the cylinder to the tip (0 .. 351.01), the tip (.. 371.01), the bottom edge back
    (.. 722.02); uniform spacing inside every block face (the flag-side gradings of blocks 8 and 13 are ignored: only
    the overlap structure matters for the transfer).  -> (s0, s1, k) arrays of the fluid and of the solid patch, in
    patch face order (position along the block face fastest, then the layer: the order of patch_cells)."""
    Ltop, Ltip, x19 = 600.0 - 248.9898, 20.0, 299.5733 - 248.9898
    s_fluid = {(19, 18): (x19, 0.0), (20, 19): (Ltop, x19), (20, 14): (Ltop, Ltop + Ltip),
               (13, 14): (2 * Ltop + Ltip - x19, Ltop + Ltip), (12, 13): (2 * Ltop + Ltip, 2 * Ltop + Ltip - x19)}
    s_solid = {(3, 2): (0.0, Ltop), (2, 1): (Ltop, Ltop + Ltip), (1, 0): (Ltop + Ltip, 2 * Ltop + Ltip)}

    def pieces(mesh, edges, table):
        s0, s1, kk = [], [], []
        for va, vb in edges:
            b, name, _ = mesh.find_edge(va, vb)
            n = mesh.edge_cells(b, name).shape[1]
            e = np.linspace(table[(va, vb)][0], table[(va, vb)][1], n + 1)   # from va to vb, like patch_cells
            lo_, hi = np.minimum(e[:-1], e[1:]), np.maximum(e[:-1], e[1:])
            for k in range(mesh.layers):
                s0.append(lo_)
                s1.append(hi)
                kk.append(np.full(n, k))
        return np.concatenate(s0), np.concatenate(s1), np.concatenate(kk)

    return pieces(fluid, HT_FLUID_INTERFACE, s_fluid), pieces(solid, HT_SOLID_INTERFACE, s_solid)


def interval_ggi(to, frm):
    """GGI weights between two patches given as intervals (s0, s1, k): for every receiving face the overlapping faces of
    the other patch in ascending face order and overlap / own length, rows rescaled to sum to one (what
    GGIInterpolation computes for faces that are rectangles [s0, s1] x layer k in one plane).  -> CSR offsets, addr, weights."""
    t0, t1, tk = to
    f0, f1, fk = frm
    order = np.argsort(f0, kind="stable")
    offsets, addr, w = [0], [], []
    for i in range(t0.size):
        cand = order[(f1[order] > t0[i] + 1e-12) & (f0[order] < t1[i] - 1e-12) & (fk[order] == tk[i])]
        cand = np.sort(cand)
        ov = np.minimum(f1[cand], t1[i]) - np.maximum(f0[cand], t0[i])
        ww = ov / (t1[i] - t0[i])
        ww = ww / ww.sum() if ww.size else ww
        addr.append(cand)
        w.append(ww)
        offsets.append(offsets[-1] + cand.size)
    return (np.asarray(offsets, np.int32), np.concatenate(addr).astype(np.int32) if addr else np.zeros(0, np.int32),
            np.concatenate(w) if w else np.zeros(0))
