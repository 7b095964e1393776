"""CPU ORACLE (TEST INFRASTRUCTURE ONLY) of the GGI weight construction: GGIInterpolation's published algorithm
(foam-extend 4.1 GGIInterpolationWeights.C, not in /root/reference -- PARITY UNPINNED; anchored on the reference's call
site ggiInterfaceToInterfaceMapping.C:62-77: tolerances SMALL, rescale = true) restated independently of the product
code: no broad phase (all pairs), the MASTER polygon clipped by the SLAVE's half-planes (the product clips the slave by the
master's), basis from the normal and a coordinate axis (the product uses the first edge).  Pure Python: small cases only.
Also the closed form for two rectangular grids in one plane (products of interval overlaps): the known answers the
restatement is pinned to.  Only tests/ may import this module."""
from __future__ import annotations

import numpy as np

SMALL = 1e-15
FEATURE_COS = 0.8


def _centre_normal(P):
    n = len(P)
    if n == 3:
        return P.mean(0), 0.5 * np.cross(P[1] - P[0], P[2] - P[0])
    avg = P.mean(0)
    sumN, sumA, sumAc = np.zeros(3), 0.0, np.zeros(3)
    for i in range(n):
        p, q = P[i], P[(i + 1) % n]
        tn = np.cross(q - p, avg - p)
        a = np.linalg.norm(tn)
        sumN += tn
        sumA += a
        sumAc += a * (p + q + avg)
    return (sumAc / (3.0 * sumA) if sumA > 0 else avg), 0.5 * sumN


def _signed_area(Q):
    x, y = Q[:, 0], Q[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _clip(subject, clipper):
    """subject (any orientation) clipped by the convex polygon clipper (made counter-clockwise first)."""
    if _signed_area(clipper) < 0:
        clipper = clipper[::-1]
    out = [tuple(p) for p in subject]
    m = len(clipper)
    for k in range(m):
        a, b = clipper[k], clipper[(k + 1) % m]
        inp, out = out, []
        if not inp:
            break
        side = lambda p: (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
        for i in range(len(inp)):
            p, q = inp[i], inp[(i + 1) % len(inp)]
            dp, dq = side(p), side(q)
            if dp >= 0:
                out.append(p)
            if (dp >= 0) != (dq >= 0):
                t = dp / (dp - dq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return np.array(out) if len(out) >= 3 else None


def ggi_weights(mFaces, mPts, sFaces, sPts, tol=SMALL, rescale=True):
    """-> offsets, addr, weights (CSR over the master faces); faces are lists of point labels."""
    mPts, sPts = np.asarray(mPts, float), np.asarray(sPts, float)
    off, addr, wts = [0], [], []
    sInfo = [_centre_normal(sPts[list(f)]) for f in sFaces]
    for f in mFaces:
        P = mPts[list(f)]
        c, nrm = _centre_normal(P)
        nhat = nrm / np.linalg.norm(nrm)
        ax = np.eye(3)[int(np.argmin(np.abs(nhat)))]
        u = np.cross(nhat, ax)
        u /= np.linalg.norm(u)
        v = np.cross(nhat, u)
        M2 = np.stack([(P - c) @ u, (P - c) @ v], 1)
        mA = abs(_signed_area(M2))
        row = []
        for j, g in enumerate(sFaces):
            sn = sInfo[j][1]
            if abs(nhat @ sn) / np.linalg.norm(sn) < FEATURE_COS:
                continue
            S = sPts[list(g)]
            S2 = np.stack([(S - c) @ u, (S - c) @ v], 1)
            if S2[:, 0].max() < M2[:, 0].min() or S2[:, 0].min() > M2[:, 0].max() or \
               S2[:, 1].max() < M2[:, 1].min() or S2[:, 1].min() > M2[:, 1].max():
                continue
            R = _clip(M2, S2)
            if R is None:
                continue
            w = abs(_signed_area(R)) / mA
            if w > tol:
                row.append((j, w))
        tot = sum(w for _, w in row)
        for j, w in row:
            addr.append(j)
            wts.append(w / tot if rescale and tot > 0 else w)
        off.append(len(addr))
    return np.array(off, np.int32), np.array(addr, np.int32), np.array(wts)


def rect_grid_weights(xm, ym, xs, ys, rescale=True):
    """Known answer: master grid xm x ym (faces i-fastest), slave grid xs x ys in the same plane."""
    def overlaps(a, b):
        return [[(j, min(a[i + 1], b[j + 1]) - max(a[i], b[j])) for j in range(len(b) - 1)
                 if min(a[i + 1], b[j + 1]) - max(a[i], b[j]) > 0] for i in range(len(a) - 1)]
    ox, oy = overlaps(xm, xs), overlaps(ym, ys)
    nxs = len(xs) - 1
    off, addr, wts = [0], [], []
    for j in range(len(ym) - 1):
        for i in range(len(xm) - 1):
            A = (xm[i + 1] - xm[i]) * (ym[j + 1] - ym[j])
            row = sorted((js * nxs + is_, lx * ly / A) for is_, lx in ox[i] for js, ly in oy[j])
            tot = sum(w for _, w in row)
            for a, w in row:
                addr.append(a)
                wts.append(w / tot if rescale else w)
            off.append(len(addr))
    return np.array(off, np.int32), np.array(addr, np.int32), np.array(wts)
