"""GPU tests of the GGI weight construction (b200_ggi_build / b200_ggi_fetch) through the C ABI.

 * against the CPU emulator built from the kernels' own arithmetic header (tests/test_ggi_build.py pins that emulator to
   the closed form and to the independent restatement oracle/pyggi.py): identical addressing (integer work: exact),
   weights within 1e-13 relative (FP64: same operations in the same order, IEEE division and square root on both sides;
   the tolerance only allows for the last bit);
 * directly against the closed form for rectangular grids, 1e-12;
 * conformal pair -> identity exactly; the built tables drive b200_ggi_interpolate: a constant field is reproduced,
   a linear field is reproduced to the mesh's second-order error on the non-matching pair.
"""
import numpy as np
import pytest

from ggi_helpers import grid_patch, split_triangles, to_csr
from oracle import pyggi
from test_ggi_build import as_grid_args, emu, emu_build, patches, rotation  # noqa: F401  (emu is a fixture)

pytestmark = pytest.mark.gpu

W_RTOL = 1e-13


def gpu_build(ctx, mf, mp, sf, sp, **kw):
    mo, ml = to_csr(mf)
    so, sl = to_csr(sf)
    return ctx.ggi_build(mo, ml, mp, so, sl, sp, **kw)


@pytest.mark.parametrize("kind", ["conformal", "refined", "graded", "partial"])
def test_rectangular_grids(gpu_ctx, emu, kind):
    (mf, mp), (sf, sp) = patches(kind)
    n0 = gpu_ctx.launches
    off, addr, w = gpu_build(gpu_ctx, mf, mp, sf, sp)
    assert gpu_ctx.launches == n0 + 2
    eoff, eaddr, ew = emu_build(emu, mf, mp, sf, sp)
    assert np.array_equal(off, eoff) and np.array_equal(addr, eaddr)
    np.testing.assert_allclose(w, ew, rtol=W_RTOL, atol=0)
    roff, raddr, rw = pyggi.rect_grid_weights(*as_grid_args(kind))
    assert np.array_equal(off, roff) and np.array_equal(addr, raddr)
    np.testing.assert_allclose(w, rw, rtol=1e-12, atol=1e-14)
    if kind == "conformal":
        assert np.array_equal(addr, np.arange(len(mf))) and np.array_equal(w, np.ones(len(mf)))


@pytest.mark.parametrize("variant", ["rotated", "warped", "triangles", "unscaled"])
def test_general_patches_match_emulator(gpu_ctx, emu, variant):
    (mf, mp), (sf, sp) = patches("graded")
    kw = {}
    if variant == "rotated":
        R = rotation(8)
        mp, sp = mp @ R.T + 3.0, sp @ R.T + 3.0
    if variant == "warped":
        warp = lambda X, Y: 0.05 * np.sin(2 * X) * np.cos(3 * Y)
        x = lambda n, s: np.sort(np.concatenate([[0.0, 2.0], 2.0 * np.random.default_rng(s).random(n)]))
        y = lambda n, s: np.sort(np.concatenate([[0.0, 1.0], np.random.default_rng(s).random(n)]))
        mf, mp = grid_patch(x(40, 1), y(30, 2), warp=warp)
        sf, sp = grid_patch(x(55, 3), y(21, 4), flip=True, warp=warp)
    if variant == "triangles":
        mf, sf = split_triangles(mf), split_triangles(sf)
    if variant == "unscaled":
        kw = dict(rescale=False, nonOverlapTol=1e-3)
    off, addr, w = gpu_build(gpu_ctx, mf, mp, sf, sp, **kw)
    eoff, eaddr, ew = emu_build(emu, mf, mp, sf, sp, tol=kw.get("nonOverlapTol", 1e-15), rescale=kw.get("rescale", True))
    assert np.array_equal(off, eoff) and np.array_equal(addr, eaddr)
    np.testing.assert_allclose(w, ew, rtol=W_RTOL, atol=0)
    if not kw:
        np.testing.assert_allclose(np.add.reduceat(w, off[:-1]), 1.0, rtol=1e-14)


def test_built_tables_drive_the_face_transfer(gpu_ctx):
    (mf, mp), (sf, sp) = patches("graded")
    off, addr, w = gpu_build(gpu_ctx, mf, mp, sf, sp)
    centre = lambda f, p: np.array([p[q].mean(0) for q in f])
    cm, cs = centre(mf, mp), centre(sf, sp)
    const = gpu_ctx.ggi_interpolate(off, addr, w, np.full(len(sf), 7.25))
    np.testing.assert_allclose(const, 7.25, rtol=1e-14)
    lin = gpu_ctx.ggi_interpolate(off, addr, w, 2.0 * cs[:, 0] - 3.0 * cs[:, 1])
    # area-weighted averages of a linear field over the overlaps: exact up to the offset between a face centre and the
    # centroid of what covers it, i.e. bounded by the face size times the gradient
    assert np.max(np.abs(lin - (2.0 * cm[:, 0] - 3.0 * cm[:, 1]))) < 0.5


def test_empty_and_errors(gpu_ctx):
    from multiregionfoam_b200 import ldu
    (mf, mp), (sf, sp) = patches("refined")
    off, addr, w = gpu_build(gpu_ctx, mf, mp, [], np.zeros((0, 3)))
    assert addr.size == 0 and not off.any()
    off, addr, w = gpu_build(gpu_ctx, mf, mp, sf, sp + np.array([10.0, 0, 0]))
    assert addr.size == 0 and off.size == len(mf) + 1
    with pytest.raises(ldu.B200Error) as e:      # a 9-point face
        gpu_build(gpu_ctx, [list(range(9))], np.random.default_rng(0).random((9, 3)), sf, sp)
    assert e.value.code == -6
    with pytest.raises(ldu.B200Error) as e:      # label out of range
        gpu_build(gpu_ctx, [[0, 1, 99]], mp[:3], sf, sp)
    assert e.value.code == -1
