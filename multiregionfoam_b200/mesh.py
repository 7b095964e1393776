"""Structured multi-block meshes in OpenFOAM LDU (owner/neighbour) form, and a reader
for the shipped ``polyMesh`` addressing files.

The generator reproduces blockMesh's cell and face ordering for blocks joined along x
(cells i-fastest inside a block, blocks concatenated in dictionary order, internal faces
in upper-triangular order = all adjacent cell pairs (min,max) sorted lexicographically).
At refinement r=1, L=1 it reproduces ``owner[:nInternal]``/``neighbour`` of

    /root/reference/tutorials/conjugateHeatTransfer/flowOverHeatedPlate/constant/fluid/polyMesh
    /root/reference/tutorials/conjugateHeatTransfer/flowOverHeatedPlate/constant/solid/polyMesh

exactly (fluid 13 612 cells / 26 851 internal faces, solid 8 200 / 16 159); the block
dimensions and gradings are those of ``blockMeshDictMonolithic`` (fluid :45-50, solid :35-38).
That identity is the golden test of this module (tests/golden/flowOverHeatedPlate_addr.npz).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------- geometry

def graded_widths(n: int, length: float, expansion: float) -> np.ndarray:
    """Cell widths of a blockMesh edge with ``simpleGrading`` expansion ratio
    (= last cell width / first cell width), geometric progression."""
    if n == 1 or abs(expansion - 1.0) < 1e-14:
        return np.full(n, length / n)
    r = expansion ** (1.0 / (n - 1))
    w = r ** np.arange(n, dtype=np.float64)
    return w * (length / w.sum())


@dataclass
class Block:
    nx: int
    x0: float
    x1: float
    gradx: float = 1.0


@dataclass
class StructuredRegion:
    """Blocks joined along x sharing the same (ny, nz) cross-section."""

    name: str
    blocks: List[Block]
    ny: int
    nz: int
    y0: float
    y1: float
    grady: float
    z0: float = 0.0
    z1: float = 0.4

    # filled by build()
    nCells: int = 0
    nFaces: int = 0
    lowerAddr: np.ndarray = field(default=None, repr=False)
    upperAddr: np.ndarray = field(default=None, repr=False)
    faceArea: np.ndarray = field(default=None, repr=False)   # |S_f| of internal faces
    faceDelta: np.ndarray = field(default=None, repr=False)  # |d| between cell centres
    faceDir: np.ndarray = field(default=None, repr=False)    # 0=x,1=y,2=z (int8)
    volume: np.ndarray = field(default=None, repr=False)
    blockOffsets: np.ndarray = field(default=None, repr=False)

    def dims(self) -> Tuple[int, int, int]:
        return (sum(b.nx for b in self.blocks), self.ny, self.nz)

    def build(self) -> "StructuredRegion":
        ny, nz = self.ny, self.nz
        dy = graded_widths(ny, self.y1 - self.y0, self.grady)
        dz = graded_widths(nz, self.z1 - self.z0, 1.0)
        dxs = [graded_widths(b.nx, b.x1 - b.x0, b.gradx) for b in self.blocks]
        offs = np.zeros(len(self.blocks) + 1, dtype=np.int64)
        for b, blk in enumerate(self.blocks):
            offs[b + 1] = offs[b] + blk.nx * ny * nz
        N = int(offs[-1])
        if N >= 2**31:
            raise ValueError("region too large for int32 labels")
        self.blockOffsets = offs
        self.nCells = N

        # per-cell number of upper neighbours -> face start offsets
        cnt = np.empty(N, dtype=np.int8)
        for b, blk in enumerate(self.blocks):
            nx = blk.nx
            c = np.zeros((nz, ny, nx), dtype=np.int8)
            c[:, :, : nx - 1] += 1                      # +x in block
            if b < len(self.blocks) - 1:
                c[:, :, nx - 1] += 1                    # +x across the block seam
            c[:, : ny - 1, :] += 1                      # +y
            c[: nz - 1, :, :] += 1                      # +z
            cnt[offs[b]:offs[b + 1]] = c.reshape(-1)
        start = np.zeros(N + 1, dtype=np.int64)
        np.cumsum(cnt, out=start[1:])
        F = int(start[-1])
        self.nFaces = F
        l = np.empty(F, dtype=np.int32)
        u = np.empty(F, dtype=np.int32)
        area = np.empty(F, dtype=np.float64)
        delta = np.empty(F, dtype=np.float64)
        fdir = np.empty(F, dtype=np.int8)
        vol = np.empty(N, dtype=np.float64)

        for b, blk in enumerate(self.blocks):
            nx = blk.nx
            dx = dxs[b]
            o = int(offs[b])
            nb = nx * ny * nz
            cell = (o + np.arange(nb, dtype=np.int64)).reshape(nz, ny, nx)
            pos = start[o:o + nb].reshape(nz, ny, nx).copy()  # next free face slot of each cell
            DX = dx[None, None, :]
            DY = dy[None, :, None]
            DZ = dz[:, None, None]
            vol[o:o + nb] = (DX * DY * DZ).reshape(-1)

            def put(mask_slices, upper_cells, a, d, direction):
                idx = pos[mask_slices].reshape(-1)
                l[idx] = cell[mask_slices].reshape(-1)
                u[idx] = upper_cells.reshape(-1)
                area[idx] = np.broadcast_to(a, cell[mask_slices].shape).reshape(-1)
                delta[idx] = np.broadcast_to(d, cell[mask_slices].shape).reshape(-1)
                fdir[idx] = direction
                pos[mask_slices] += 1

            # slot 0: +x inside the block (u = c+1)
            if nx > 1:
                sl = (slice(None), slice(None), slice(0, nx - 1))
                put(sl, cell[:, :, 1:], DY * DZ, (0.5 * (dx[:-1] + dx[1:]))[None, None, :], 0)
            # slot 1: +y (u = c+nx)
            if ny > 1:
                sl = (slice(None), slice(0, ny - 1), slice(None))
                put(sl, cell[:, 1:, :], DX * DZ, (0.5 * (dy[:-1] + dy[1:]))[None, :, None], 1)
            # slot 2: +z (u = c+nx*ny)
            if nz > 1:
                sl = (slice(0, nz - 1), slice(None), slice(None))
                put(sl, cell[1:, :, :], DX * DY, (0.5 * (dz[:-1] + dz[1:]))[:, None, None], 2)
            # slot 3: +x across the seam to block b+1 (largest upper index of the cell)
            if b < len(self.blocks) - 1:
                nxn = self.blocks[b + 1].nx
                on = int(offs[b + 1])
                jj = np.arange(ny, dtype=np.int64)[None, :, None]
                kk = np.arange(nz, dtype=np.int64)[:, None, None]
                nbr = on + nxn * (jj + ny * kk)          # cell (0,j,k) of block b+1
                sl = (slice(None), slice(None), slice(nx - 1, nx))
                put(sl, np.broadcast_to(nbr, (nz, ny, 1)), DY * DZ,
                    0.5 * (dx[-1] + dxs[b + 1][0]), 0)
        self.lowerAddr, self.upperAddr = l, u
        self.faceArea, self.faceDelta, self.faceDir, self.volume = area, delta, fdir, vol
        self._dx, self._dy, self._dz = dxs, dy, dz
        return self

    # ---- boundary helpers (cells adjacent to a side + face area + centre-to-face distance)
    def _cells(self, b: int) -> np.ndarray:
        blk = self.blocks[b]
        o = int(self.blockOffsets[b])
        return (o + np.arange(blk.nx * self.ny * self.nz, dtype=np.int64)).reshape(self.nz, self.ny, blk.nx)

    def side_xmin(self):
        c = self._cells(0)[:, :, 0]
        a = self._dy[None, :] * self._dz[:, None]
        return c.reshape(-1).astype(np.int32), a.reshape(-1), np.full(c.size, 0.5 * self._dx[0][0])

    def side_xmax(self):
        c = self._cells(len(self.blocks) - 1)[:, :, -1]
        a = self._dy[None, :] * self._dz[:, None]
        return c.reshape(-1).astype(np.int32), a.reshape(-1), np.full(c.size, 0.5 * self._dx[-1][-1])

    def side_y(self, b: int, top: bool):
        """Cells of block b adjacent to y-min (top=False) or y-max, ordered (k, i) i-fastest --
        the order blockMesh gives the patch faces of one block side."""
        j = self.ny - 1 if top else 0
        c = self._cells(b)[:, j, :]
        a = self._dx[b][None, :] * self._dz[:, None]
        return c.reshape(-1).astype(np.int32), a.reshape(-1), np.full(c.size, 0.5 * self._dy[j])

    def side_z(self, top: bool):
        """Cells adjacent to z-min (top=False) or z-max in ascending cell order (block by block, (j, i)
        i-fastest): the order of the faces of a z-cut processor patch on both sides of the cut."""
        k = self.nz - 1 if top else 0
        cs, as_ = [], []
        for b in range(len(self.blocks)):
            cs.append(self._cells(b)[k, :, :].reshape(-1))
            as_.append((self._dx[b][None, :] * self._dy[:, None]).reshape(-1))
        c = np.concatenate(cs)
        return c.astype(np.int32), np.concatenate(as_), np.full(c.size, 0.5 * self._dz[k])

    def cell_centres_x(self) -> np.ndarray:
        xs = []
        for b, blk in enumerate(self.blocks):
            dx = self._dx[b]
            xc = blk.x0 + np.cumsum(dx) - 0.5 * dx
            xs.append(np.broadcast_to(xc[None, None, :], (self.nz, self.ny, blk.nx)).reshape(-1))
        return np.concatenate(xs)

    def cell_xyz(self):
        """Cell-centre coordinates as simpleGeomDecomp sorts them (x exact; y and z through the row / layer index, which
        is monotone in the coordinate)."""
        return self.cell_centres_x(), self.cell_j().astype(np.float64), self.cell_ijk_layer().astype(np.float64)

    def cell_ijk_layer(self) -> np.ndarray:
        """z-layer index k of every cell (used by the z-slab decomposition)."""
        ks = []
        for b, blk in enumerate(self.blocks):
            kk = np.arange(self.nz, dtype=np.int32)[:, None, None]
            ks.append(np.broadcast_to(kk, (self.nz, self.ny, blk.nx)).reshape(-1))
        return np.concatenate(ks)

    def cell_j(self) -> np.ndarray:
        js = []
        for b, blk in enumerate(self.blocks):
            jj = np.arange(self.ny, dtype=np.int32)[None, :, None]
            js.append(np.broadcast_to(jj, (self.nz, self.ny, blk.nx)).reshape(-1))
        return np.concatenate(js)


def flow_over_heated_plate(r: int = 1, layers: int = 1, z1: float = 0.4) -> Tuple[StructuredRegion, StructuredRegion]:
    """The two regions of tutorials/conjugateHeatTransfer/flowOverHeatedPlate refined by the
    integer factor r in x and y and extruded to ``layers`` cells in z
    (blockMeshDictMonolithic: fluid blocks (81 41 1)(200 41 1)(51 41 1) simpleGrading
    (.2 16 1)(5 16 1)(1 16 1); solid (200 41 1) simpleGrading (5 0.0625 1))."""
    fluid = StructuredRegion(
        "fluid",
        [Block(81 * r, -0.5, 0.0, 0.2), Block(200 * r, 0.0, 1.0, 5.0), Block(51 * r, 1.0, 3.0, 1.0)],
        ny=41 * r, nz=layers, y0=0.0, y1=0.5, grady=16.0, z1=z1).build()
    solid = StructuredRegion(
        "solid", [Block(200 * r, 0.0, 1.0, 5.0)],
        ny=41 * r, nz=layers, y0=-0.25, y1=0.0, grady=0.0625, z1=z1).build()
    return fluid, solid


# --------------------------------------------------------------------------- polyMesh reader

_LIST_RE = re.compile(rb"\n\s*(\d+)\s*\n?\(")


def read_label_list(path: str) -> np.ndarray:
    """Read an OpenFOAM ascii ``labelList`` (owner / neighbour)."""
    data = open(path, "rb").read()
    hdr_end = data.find(b"}")  # end of FoamFile header
    m = _LIST_RE.search(data, hdr_end)
    if not m:
        raise ValueError(f"no list found in {path}")
    n = int(m.group(1))
    body = data[m.end():data.rfind(b")")]
    arr = np.array(body.split(), dtype=np.int64)
    if arr.size != n:
        raise ValueError(f"{path}: expected {n} labels, found {arr.size}")
    return arr.astype(np.int32)


def read_boundary(path: str) -> List[dict]:
    """Read a ``polyBoundaryMesh`` file into a list of dicts (name, type, nFaces, startFace, ...)."""
    txt = open(path, "r").read()
    txt = re.sub(r"//.*", "", txt)
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    body = txt[txt.find("}") + 1:]
    m = re.search(r"(\d+)\s*\(", body)
    inner = body[m.end():body.rfind(")")]
    patches = []
    for pm in re.finditer(r"([\w\-]+)\s*\{([^}]*)\}", inner):
        d = {"name": pm.group(1)}
        for em in re.finditer(r"([\w]+)\s+([^;]+);", pm.group(2)):
            v = em.group(2).strip()
            d[em.group(1)] = int(v) if re.fullmatch(r"-?\d+", v) else v
        patches.append(d)
    return patches


def read_polymesh_addressing(polymesh_dir: str):
    """Return (nCells, lowerAddr, upperAddr, patches) with patch faceCells attached."""
    owner = read_label_list(f"{polymesh_dir}/owner")
    nbr = read_label_list(f"{polymesh_dir}/neighbour")
    nInt = nbr.size
    nCells = int(max(owner.max(), nbr.max())) + 1
    patches = read_boundary(f"{polymesh_dir}/boundary")
    for p in patches:
        s, n = p["startFace"], p["nFaces"]
        p["faceCells"] = owner[s:s + n].copy()
    return nCells, owner[:nInt].copy(), nbr, patches


def is_upper_triangular(l: np.ndarray, u: np.ndarray) -> bool:
    """owner < neighbour and faces sorted lexicographically by (owner, neighbour)."""
    if l.size == 0:
        return True
    if not np.all(l < u):
        return False
    key = l.astype(np.int64) * (int(u.max()) + 1) + u.astype(np.int64)
    return bool(np.all(np.diff(key) > 0))


def level_stats(nCells: int, l: np.ndarray, u: np.ndarray):
    """Wavefront (level-set) statistics of the strict lower-triangular dependency graph."""
    lev = np.zeros(nCells, dtype=np.int32)
    # faces sorted by l then u: processing in face order visits l ascending, so lev[l] is final
    # by the time face (l,u) is seen only if all faces into l precede; true in upper-tri order.
    order = np.argsort(u, kind="stable")
    # simple (slow) pass; only for small meshes in tests
    for f in range(l.size):
        a, b = l[f], u[f]
        if lev[b] < lev[a] + 1:
            lev[b] = lev[a] + 1
    del order
    nlev = int(lev.max()) + 1 if nCells else 0
    counts = np.bincount(lev, minlength=nlev)
    return nlev, counts
