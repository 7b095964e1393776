"""LDU dump format "B200LDU1" (INTEGRATION.md): one rank's coupled LDU system exactly as the solver
receives it -- addressing, coefficients, interface tables, source, initial guess -- plus the solver
controls and the residual history the reference printed.  A foam-extend host writes it from the
adapter (``b200Dump yes;``); here it is read back to replay the solve on the GPU and in the oracle,
which turns "parity unpinned" into parity against foam-extend's own numbers.

Layout (little endian):
  char[8]  "B200LDU1"
  i32      rank, nRanks, nRegions
  per region:  i32 nCells, nFaces, symmetric, nIfaces
               i32 lowerAddr[nFaces], upperAddr[nFaces]
               f64 diag[nCells], upper[nFaces], (lower[nFaces] if !symmetric), source[nCells], psi0[nCells]
    per iface: i32 kind, nFaces, peerRank, peerRegion, peerIface, nPeerFaces, hasGgi
               i32 faceCells[nFaces];  f64 bouCoeffs[nFaces], intCoeffs[nFaces]
               if hasGgi: i32 offsets[nFaces+1], addr[nnz];  f64 weights[nnz]
  char[32] solver, char[32] preconditioner;  f64 tolerance, relTol;  i32 minIter, maxIter
  i32 nHistory;  f64 history[nHistory]      (normalised residual after 0, 1, ... iterations; may be 0)
"""
from __future__ import annotations

import struct
from typing import Optional, Tuple

import numpy as np

from .case import Interface, RankSystem, Region

MAGIC = b"B200LDU1"


def write_dump(path: str, rs: RankSystem, controls: dict, history: Optional[np.ndarray] = None) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<3i", rs.rank, rs.nRanks, len(rs.regions)))
        for reg in rs.regions:
            f.write(struct.pack("<4i", reg.nCells, reg.nFaces, int(reg.lower is None), len(reg.interfaces)))
            for a, dt in ((reg.lowerAddr, "<i4"), (reg.upperAddr, "<i4"), (reg.diag, "<f8"), (reg.upper, "<f8")):
                f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
            if reg.lower is not None:
                f.write(np.ascontiguousarray(reg.lower, dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(reg.source, dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(reg.psi, dtype="<f8").tobytes())
            for itf in reg.interfaces:
                has = itf.ggiOffsets is not None
                nPeer = itf.nFaces if itf.nPeerFaces is None else itf.nPeerFaces
                f.write(struct.pack("<7i", itf.kind, itf.nFaces, itf.peerRank, itf.peerRegion, itf.peerIface, nPeer, int(has)))
                f.write(np.ascontiguousarray(itf.faceCells, dtype="<i4").tobytes())
                f.write(np.ascontiguousarray(itf.bouCoeffs, dtype="<f8").tobytes())
                f.write(np.ascontiguousarray(itf.intCoeffs, dtype="<f8").tobytes())
                if has:
                    f.write(np.ascontiguousarray(itf.ggiOffsets, dtype="<i4").tobytes())
                    f.write(np.ascontiguousarray(itf.ggiAddr, dtype="<i4").tobytes())
                    f.write(np.ascontiguousarray(itf.ggiWeights, dtype="<f8").tobytes())
        f.write(struct.pack("<32s32s", str(controls.get("solver", "")).encode(), str(controls.get("preconditioner", "")).encode()))
        f.write(struct.pack("<2d2i", float(controls.get("tolerance", 1e-6)), float(controls.get("relTol", 0.0)),
                            int(controls.get("minIter", 0)), int(controls.get("maxIter", 1000))))
        h = np.zeros(0) if history is None else np.ascontiguousarray(history, dtype="<f8")
        f.write(struct.pack("<i", h.size))
        f.write(h.tobytes())


def read_dump(path: str) -> Tuple[RankSystem, dict, np.ndarray]:
    data = open(path, "rb").read()
    if data[:8] != MAGIC:
        raise ValueError(f"{path}: not a B200LDU1 dump")
    pos = 8

    def ints(n):
        nonlocal pos
        v = struct.unpack_from(f"<{n}i", data, pos)
        pos += 4 * n
        return v

    def arr(n, dt):
        nonlocal pos
        a = np.frombuffer(data, dtype=dt, count=n, offset=pos).copy()
        pos += a.nbytes
        return a

    rank, nRanks, nReg = ints(3)
    regions = []
    for r in range(nReg):
        nCells, nFaces, sym, nIf = ints(4)
        l, u = arr(nFaces, "<i4"), arr(nFaces, "<i4")
        diag, upper = arr(nCells, "<f8"), arr(nFaces, "<f8")
        lower = None if sym else arr(nFaces, "<f8")
        source, psi = arr(nCells, "<f8"), arr(nCells, "<f8")
        reg = Region(f"region{r}", nCells, l, u, diag, upper, lower, source, psi)
        for _ in range(nIf):
            kind, nF, pr, pg, pi, nPeer, has = ints(7)
            fc, bou, inc = arr(nF, "<i4"), arr(nF, "<f8"), arr(nF, "<f8")
            go = ga = gw = None
            if has:
                go = arr(nF + 1, "<i4")
                ga = arr(int(go[-1]), "<i4")
                gw = arr(int(go[-1]), "<f8")
            reg.interfaces.append(Interface(kind, fc, bou, inc, pr, pg, pi, go, ga, gw, nPeerFaces=nPeer))
        regions.append(reg)
    solver, precond = struct.unpack_from("<32s32s", data, pos)
    pos += 64
    tol, rel, mn, mx = struct.unpack_from("<2d2i", data, pos)
    pos += 24
    (nh,) = ints(1)
    hist = arr(nh, "<f8")
    controls = dict(solver=solver.rstrip(b"\0").decode(), preconditioner=precond.rstrip(b"\0").decode(), tolerance=tol,
                    relTol=rel, minIter=mn, maxIter=mx)
    return RankSystem(rank, nRanks, regions), controls, hist
