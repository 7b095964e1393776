"""CPU tests of the device-side T-equation assembly.

 * oracle/fv_oracle.c (the operators restated one fvMatrix at a time) against the independent numpy assembler of the
   CHT case (multiregionfoam_b200/assembly.py) -- tolerance 1e-12: different summation order;
 * the kernels' arithmetic (multiregionfoam_b200/csrc/fv_assemble.hpp, compiled for the CPU into
   tests/_build/libassemble_emu.so) against the oracle: BIT-EXACT, on the CHT regions and on seeded unstructured
   addressings with mixed flux signs, face conductivities, several boundary faces per cell and a permuted slot order.
The GPU leg (tests/test_gpu_zassemble.py) runs the same cases through the C ABI.
"""
import ctypes as C

import numpy as np
import pytest

from multiregionfoam_b200 import build as b200build
from multiregionfoam_b200.assembly import assemble_cht, cht_fv_tables, cht_fv_tables_slab, cht_rank_slab
from multiregionfoam_b200.mesh import flow_over_heated_plate
from oracle import pyfv


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(b200build.build_assemble_emulator())
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.emu_assemble_T.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, C.c_double, C.c_double, C.c_double, dp, dp, dp, dp, dp,
                                 C.c_int, ip, dp, dp, ip, dp, dp, dp, dp, dp]
    return L


def _d(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def emulate(L, form, l, u, rhoC, rDeltaT, kappa, V, magSf, delta, Told, *, kappaFace=None, phi=None, bCells=None, bInt=None,
            bSrc=None, seed=0):
    """Runs the kernel arithmetic with the resident vectors in a random slot order (padding slots in between)."""
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
    i32 = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)
    l, u, V, magSf, delta, Told = i32(l), i32(u), f64(V), f64(magSf), f64(delta), f64(Told)
    kappaFace, phi, bCells, bInt, bSrc = f64(kappaFace), f64(phi), i32(bCells), f64(bInt), f64(bSrc)
    n, nf = V.size, l.size
    nSlots = n + n // 3 + 5
    slotOfCell = i32(np.random.default_rng(seed).permutation(nSlots)[:n])
    xSlots = np.full(nSlots, np.nan)
    xSlots[slotOfCell] = Told
    bSlots = np.zeros(nSlots)
    diag, upper, lower = np.empty(n), np.empty(nf), np.empty(nf)
    rc = L.emu_assemble_T(form, n, nf, _i(l), _i(u), rhoC, rDeltaT, kappa, _d(kappaFace), _d(V), _d(magSf), _d(delta), _d(phi),
                          0 if bCells is None else bCells.size, _i(bCells), _d(bInt), _d(bSrc), _i(slotOfCell), _d(xSlots),
                          _d(diag), _d(upper), _d(lower), _d(bSlots))
    assert rc == 0
    pad = np.ones(nSlots, bool)
    pad[slotOfCell] = False
    assert not bSlots[pad].any()
    return diag, upper, lower, bSlots[slotOfCell]


def random_fv_case(n, seed, transport):
    """Seeded unstructured upper-triangular addressing with FV-like tables."""
    rng = np.random.default_rng(seed)
    pairs = set()
    for c in range(n - 1):
        for nb in rng.choice(np.arange(c + 1, min(n, c + 40)), size=min(3, n - 1 - c), replace=False):
            pairs.add((c, int(nb)))
    pairs = sorted(pairs)
    l = np.array([p[0] for p in pairs], np.int32)
    u = np.array([p[1] for p in pairs], np.int32)
    nf = l.size
    nB = n // 2
    t = dict(l=l, u=u, V=0.5 + rng.random(n), magSf=0.1 + rng.random(nf), delta=1.0 + 3.0 * rng.random(nf),
             Told=300.0 + 10.0 * rng.random(n), kappaFace=1.0 + rng.random(nf),
             phi=(rng.standard_normal(nf) * (rng.random(nf) > 0.2)) if transport else None,   # both signs and exact zeros
             bCells=rng.integers(0, max(n // 4, 1), nB).astype(np.int32),                          # several faces per cell
             bInt=rng.random(nB), bSrc=300.0 * rng.random(nB))
    return t


def _all_equal(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("r,layers", [(1, 1), (1, 3)])
def test_oracle_matches_numpy_assembler(r, layers):
    fluid, solid = flow_over_heated_plate(r, layers)
    case = assemble_cht(fluid, solid)
    for reg, mesh, t in zip(case.ranks[0].regions, (fluid, solid), cht_fv_tables(fluid, solid)):
        d, up, lo, src = pyfv.assemble_T(t["form"], mesh.lowerAddr, mesh.upperAddr, t["rhoC"], t["rDeltaT"], t["kappa"], t["V"],
                                         t["magSf"], t["deltaCoeffs"], t["T0"], phi=t["phi"], bCells=t["bCells"],
                                         bInt=t["bInt"], bSrc=t["bSrc"])
        np.testing.assert_allclose(d, reg.diag, rtol=1e-12)
        np.testing.assert_allclose(up, reg.upper, rtol=1e-12)
        np.testing.assert_allclose(lo, reg.upper if reg.lower is None else reg.lower, rtol=1e-12)
        np.testing.assert_allclose(src, reg.source, rtol=1e-12)
        if reg.lower is None:
            assert np.array_equal(up, lo)  # the conduction equation stays symmetric


@pytest.mark.parametrize("r,layers", [(1, 1), (1, 4)])
def test_kernel_arithmetic_bit_exact_cht(emu, r, layers):
    fluid, solid = flow_over_heated_plate(r, layers)
    for mesh, t in zip((fluid, solid), cht_fv_tables(fluid, solid)):
        kw = dict(phi=t["phi"], bCells=t["bCells"], bInt=t["bInt"], bSrc=t["bSrc"])
        Told = t["T0"] + np.random.default_rng(5).random(mesh.nCells)
        args = (t["form"], mesh.lowerAddr, mesh.upperAddr, t["rhoC"], t["rDeltaT"], t["kappa"], t["V"], t["magSf"],
                t["deltaCoeffs"], Told)
        assert _all_equal(emulate(emu, *args, **kw), pyfv.assemble_T(*args, **kw))


@pytest.mark.parametrize("n", [1, 2, 57, 4000])
@pytest.mark.parametrize("transport", [False, True])
@pytest.mark.parametrize("faceKappa", [False, True])
def test_kernel_arithmetic_bit_exact_unstructured(emu, n, transport, faceKappa):
    t = random_fv_case(n, 100 + n, transport)
    kw = dict(kappaFace=t["kappaFace"] if faceKappa else None, phi=t["phi"], bCells=t["bCells"], bInt=t["bInt"], bSrc=t["bSrc"])
    args = (int(transport), t["l"], t["u"], 250.0, 100.0, 5.0, t["V"], t["magSf"], t["delta"], t["Told"])
    assert _all_equal(emulate(emu, *args, seed=n, **kw), pyfv.assemble_T(*args, **kw))


def test_no_boundary_faces_and_no_flux(emu):
    t = random_fv_case(300, 7, False)
    args = (1, t["l"], t["u"], 2.0, 10.0, 0.3, t["V"], t["magSf"], t["delta"], t["Told"])  # transport form without phi
    got, ref = emulate(emu, *args), pyfv.assemble_T(*args)
    assert _all_equal(got, ref)
    assert np.array_equal(got[1], got[2])


def test_assembled_row_sums(emu):
    """Size-independent property: without boundary faces, the row sum of the conduction matrix is the ddt diagonal
    (the laplacian is conservative), and the column sums of the convection part vanish."""
    t = random_fv_case(2000, 11, True)
    rhoC, rdt, kappa = 3.0, 50.0, 0.7
    d, up, lo, _ = emulate(emu, 0, t["l"], t["u"], rhoC, rdt, kappa, t["V"], t["magSf"], t["delta"], t["Told"])
    rows = d + np.bincount(t["l"], weights=up, minlength=2000) + np.bincount(t["u"], weights=lo, minlength=2000)
    np.testing.assert_allclose(rows, rdt * rhoC * t["V"], rtol=1e-10)
    d1, up1, lo1, _ = emulate(emu, 1, t["l"], t["u"], rhoC, rdt, kappa, t["V"], t["magSf"], t["delta"], t["Told"], phi=t["phi"])
    # columns: diag[c] + sum of the coefficients multiplying x[c] in other rows = ddt diagonal (div and lap are conservative)
    cols = d1 + np.bincount(t["u"], weights=up1, minlength=2000) + np.bincount(t["l"], weights=lo1, minlength=2000)
    np.testing.assert_allclose(cols, rhoC * rdt * t["V"], rtol=1e-9)


@pytest.mark.parametrize("rank,nranks", [(0, 2), (1, 2), (1, 3)])
def test_decomposed_slab_tables(emu, rank, nranks):
    """N > 1: the tables of one z-slab rank (processor patches in the boundary-face list) reproduce the matrices
    cht_rank_slab hands to the multi-GPU solve (1e-12), and the kernel arithmetic stays bit-exact against the oracle."""
    rs = cht_rank_slab(1, 2, rank, nranks)
    tables, meshes = cht_fv_tables_slab(1, 2, rank, nranks)
    for reg, mesh, t in zip(rs.regions, meshes, tables):
        kw = dict(phi=t["phi"], bCells=t["bCells"], bInt=t["bInt"], bSrc=t["bSrc"])
        args = (t["form"], mesh.lowerAddr, mesh.upperAddr, t["rhoC"], t["rDeltaT"], t["kappa"], t["V"], t["magSf"],
                t["deltaCoeffs"], t["T0"])
        d, up, lo, src = pyfv.assemble_T(*args, **kw)
        np.testing.assert_allclose(d, reg.diag, rtol=1e-12)
        np.testing.assert_allclose(up, reg.upper, rtol=1e-12)
        np.testing.assert_allclose(src, reg.source, rtol=1e-12)
        nProc = sum(i.nFaces for i in reg.interfaces[1:])      # patch list: [regionCouple, processor patches...]
        assert nProc > 0 and np.array_equal(t["bCells"][-nProc:], np.concatenate([i.faceCells for i in reg.interfaces[1:]]))
        assert _all_equal(emulate(emu, *args, **kw), (d, up, lo, src))
