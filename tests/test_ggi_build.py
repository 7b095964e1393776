"""CPU tests of the GGI weight construction: the kernels' arithmetic and the host broad phase
(multiregionfoam_b200/csrc/ggi_build.hpp compiled into tests/_build/libggi_emu.so) against
 * the closed form for rectangular grids in one plane (known answers), tolerance 1e-12;
 * the independent pure-Python restatement oracle/pyggi.py (all pairs, roles of the polygons swapped, another basis) on
   planar / warped / triangulated / rotated patches, tolerance 1e-11, identical addressing;
 * properties: conformal pairs give the identity exactly, rows sum to one after rescaling, un-rescaled weights x areas are
   symmetric between the two sides (the intersection area belongs to both), a constant field is interpolated exactly.
The GPU leg (tests/test_gpu_zggi_build.py) runs the same patches through the C ABI and must reproduce the emulator bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

from ggi_helpers import grid_patch, split_triangles, to_csr
from multiregionfoam_b200 import build as b200build
from oracle import pyggi


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(b200build.build_ggi_emulator())
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.emu_ggi_build.argtypes = [C.c_int32, ip, ip, dp, C.c_int32, ip, ip, dp, C.c_double, C.c_int, ip, C.c_int32, ip, dp]
    return L


def emu_build(L, mFaces, mPts, sFaces, sPts, tol=1e-15, rescale=True):
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    mo, ml = to_csr(mFaces)
    so, sl = to_csr(sFaces)
    mp, sp = np.ascontiguousarray(mPts, np.float64), np.ascontiguousarray(sPts, np.float64)
    off = np.zeros(len(mFaces) + 1, np.int32)
    cap = 64 * max(len(mFaces), 1)
    addr, w = np.zeros(cap, np.int32), np.zeros(cap)
    P = lambda a, t: a.ctypes.data_as(t)
    nnz = L.emu_ggi_build(len(mFaces), P(mo, ip), P(ml, ip), P(mp, dp), len(sFaces), P(so, ip), P(sl, ip), P(sp, dp), tol,
                          int(rescale), P(off, ip), cap, P(addr, ip), P(w, dp))
    assert 0 <= nnz <= cap
    return off, addr[:nnz].copy(), w[:nnz].copy()


def graded(n, lo, hi, g, seed):
    rng = np.random.default_rng(seed)
    d = 1.0 + g * rng.random(n)
    return lo + (hi - lo) * np.concatenate([[0.0], np.cumsum(d) / d.sum()])


def patches(kind):
    if kind == "conformal":
        x, y = graded(9, 0, 1, 1.0, 1), graded(7, 0, 0.5, 1.0, 2)
        return grid_patch(x, y), grid_patch(x, y, flip=True)
    if kind == "refined":       # 2:1 and 3:2 non-matching, same extent
        return grid_patch(np.linspace(0, 1, 9), np.linspace(0, 1, 7)), grid_patch(np.linspace(0, 1, 17), np.linspace(0, 1, 10), flip=True)
    if kind == "graded":        # unrelated gradings
        return grid_patch(graded(11, 0, 2, 2.0, 3), graded(6, 0, 1, 2.0, 4)), grid_patch(graded(14, 0, 2, 3.0, 5), graded(9, 0, 1, 1.0, 6), flip=True)
    if kind == "partial":       # the slave covers part of the master only
        return grid_patch(np.linspace(0, 1, 8), np.linspace(0, 1, 8)), grid_patch(np.linspace(0.33, 0.9, 6), np.linspace(-0.2, 0.61, 5), flip=True)
    raise KeyError(kind)


def as_grid_args(kind):
    (mf, mp), (sf, sp) = patches(kind)
    xs = lambda p: np.unique(p[:, 0])
    ys = lambda p: np.unique(p[:, 1])
    return xs(mp), ys(mp), xs(sp), ys(sp)


@pytest.mark.parametrize("kind", ["conformal", "refined", "graded", "partial"])
def test_known_answers_rectangular_grids(emu, kind):
    (mf, mp), (sf, sp) = patches(kind)
    off, addr, w = emu_build(emu, mf, mp, sf, sp)
    roff, raddr, rw = pyggi.rect_grid_weights(*as_grid_args(kind))
    assert np.array_equal(off, roff) and np.array_equal(addr, raddr)
    np.testing.assert_allclose(w, rw, rtol=1e-12, atol=1e-14)
    if kind == "conformal":
        assert np.array_equal(addr, np.arange(len(mf))) and np.array_equal(w, np.ones(len(mf)))


def rotation(seed):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    return Q * np.sign(np.linalg.det(Q))


@pytest.mark.parametrize("variant", ["planar", "rotated", "warped", "triangles", "mixed"])
def test_against_independent_restatement(emu, variant):
    (mf, mp), (sf, sp) = patches("graded")
    if variant == "rotated":
        R = rotation(8)
        mp, sp = mp @ R.T + 3.0, sp @ R.T + 3.0
    if variant == "warped":     # both sides on the same gently curved surface
        warp = lambda X, Y: 0.05 * np.sin(2 * X) * np.cos(3 * Y)
        mf, mp = grid_patch(graded(11, 0, 2, 2.0, 3), graded(6, 0, 1, 2.0, 4), warp=warp)
        sf, sp = grid_patch(graded(14, 0, 2, 3.0, 5), graded(9, 0, 1, 1.0, 6), flip=True, warp=warp)
    if variant == "triangles":
        mf, sf = split_triangles(mf), split_triangles(sf)
    if variant == "mixed":
        sf = sf[: len(sf) // 2] + split_triangles(sf[len(sf) // 2:])
    off, addr, w = emu_build(emu, mf, mp, sf, sp)
    roff, raddr, rw = pyggi.ggi_weights(mf, mp, sf, sp)
    if variant == "warped":
        # slivers at the tolerance may be kept by one and not by the other: compare as dense rows
        dense = lambda o, a, x: np.array([[x[o[i]:o[i + 1]][a[o[i]:o[i + 1]] == j].sum() for j in range(len(sf))] for i in range(len(mf))])
        np.testing.assert_allclose(dense(off, addr, w), dense(roff, raddr, rw), atol=1e-11)
    else:
        assert np.array_equal(off, roff) and np.array_equal(addr, raddr)
        np.testing.assert_allclose(w, rw, rtol=1e-11, atol=1e-14)
    rows = np.add.reduceat(w, off[:-1])
    np.testing.assert_allclose(rows, 1.0, rtol=1e-14)


def test_area_symmetry_and_constant_field(emu):
    (mf, mp), (sf, sp) = patches("graded")
    off, addr, w = emu_build(emu, mf, mp, sf, sp, rescale=False)
    soff, saddr, sw = emu_build(emu, sf, sp, mf, mp, rescale=False)
    area = lambda f, p: np.array([0.5 * abs(np.sum(p[q, 0] * np.roll(p[q, 1], -1) - np.roll(p[q, 0], -1) * p[q, 1])) for q in f])
    mA, sA = area(mf, mp), area(sf, sp)
    A = np.zeros((len(mf), len(sf)))
    B = np.zeros((len(sf), len(mf)))
    for i in range(len(mf)):
        A[i, addr[off[i]:off[i + 1]]] = w[off[i]:off[i + 1]] * mA[i]
    for j in range(len(sf)):
        B[j, saddr[soff[j]:soff[j + 1]]] = sw[soff[j]:soff[j + 1]] * sA[j]
    np.testing.assert_allclose(A, B.T, atol=1e-14)                      # one intersection area, seen from both sides
    np.testing.assert_allclose(A.sum(), mA.sum(), rtol=1e-12)           # full cover: areas add up
    off, addr, w = emu_build(emu, mf, mp, sf, sp)
    const = np.add.reduceat(w * 7.25, off[:-1])
    np.testing.assert_allclose(const, 7.25, rtol=1e-14)


def test_empty_and_invalid(emu):
    (mf, mp), (sf, sp) = patches("refined")
    off, addr, w = emu_build(emu, mf, mp, [], np.zeros((0, 3)))
    assert addr.size == 0 and not off.any()
    far = sp + np.array([10.0, 0, 0])
    off, addr, w = emu_build(emu, mf, mp, sf, far)                         # uncovered master faces: empty rows
    assert addr.size == 0
    perp = sp[:, [0, 2, 1]]                                                # slave patch turned by 90 degrees: feature angle
    off, addr, w = emu_build(emu, mf, mp, sf, perp)
    assert addr.size == 0
