"""CPU emulation of the device tables (csrc/schedule.hpp): walking the SELL layout and the
chain-pipelined sweep schedules exactly as the kernels do must reproduce the oracle BIT FOR BIT
(row-internal arithmetic order of lduMatrix::Amul and DIC/DILU precondition is preserved)."""
import ctypes as C

import numpy as np
import pytest

from helpers import chain_region, golden_region, random_vec
from multiregionfoam_b200.assembly import cht_case, single_region_case, synthetic_coeffs
from multiregionfoam_b200.build import build_schedule_emulator
from multiregionfoam_b200.mesh import flow_over_heated_plate
from oracle import pyoracle


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(build_schedule_emulator())
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.emu_run.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, C.c_int, dp, dp, dp, dp, dp, ip]
    return L


def run_emu(L, reg, dilu, x, r):
    n = reg.nCells
    l, u = np.ascontiguousarray(reg.lowerAddr, np.int32), np.ascontiguousarray(reg.upperAddr, np.int32)
    y, w, rD = np.empty(n), np.empty(n), np.empty(n)
    stats = np.zeros(16, np.int32)
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    I = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    lower = None if reg.lower is None else P(np.ascontiguousarray(reg.lower))
    rc = L.emu_run(n, reg.nFaces, I(l), I(u), P(reg.diag), P(reg.upper), lower, int(dilu), P(x), P(r), P(y), P(w), P(rD), I(stats))
    assert rc == 0, f"emu_run rc={rc}"
    return y, w, rD, stats


def regions(golden_addr):
    fluid, solid = flow_over_heated_plate(1, 3)
    yield synthetic_coeffs(fluid.nCells, fluid.lowerAddr, fluid.upperAddr, symmetric=False, name="fluid3d")
    yield synthetic_coeffs(solid.nCells, solid.lowerAddr, solid.upperAddr, symmetric=True, name="solid3d")
    yield golden_region(golden_addr, "bubbleA", symmetric=False)
    yield golden_region(golden_addr, "duineveld0", symmetric=True)
    yield golden_region(golden_addr, "duineveld1", symmetric=False)
    yield chain_region(1000, symmetric=False)
    yield synthetic_coeffs(1, np.empty(0, np.int32), np.empty(0, np.int32), symmetric=True, name="one-cell")


def test_tables_reproduce_oracle_bitwise(emu, golden_addr):
    for reg in regions(golden_addr):
        O = pyoracle.OracleSystem(single_region_case(reg))
        x, r = random_vec(reg.nCells, 1), random_vec(reg.nCells, 2)
        dilu = reg.lower is not None
        y, w, rD, stats = run_emu(emu, reg, dilu, x, r)
        O.precond_setup("DILU" if dilu else "DIC")
        assert np.array_equal(y, O.amul(x)), reg.name
        assert np.array_equal(rD, O.rD()), reg.name
        assert np.array_equal(w, O.precondition(r)), reg.name


def test_line_structure_on_structured_mesh(emu, monkeypatch):
    # x-lines continued across the two block seams are the paths: ny*nz of them; lines linked along j are
    # cut into warps of 32 lanes (41 = 32 + 9 per z-layer); every in-warp dependency travels by shuffle
    monkeypatch.setenv("B200_MERGE_SEQ", "0")   # the warps of one k-plane on their own
    fluid, _ = flow_over_heated_plate(1, 3)
    reg = synthetic_coeffs(fluid.nCells, fluid.lowerAddr, fluid.upperAddr, symmetric=True)
    _, _, _, stats = run_emu(emu, reg, False, random_vec(reg.nCells, 1), random_vec(reg.nCells, 2))
    nx, ny, nz = fluid.dims()
    assert stats[8] == 1                 # LINE mode
    assert stats[3] == ny * nz           # paths
    assert stats[4] == 2 * nz and stats[0] == 2 * nz   # warps
    assert stats[1] == nz + 1            # group levels: k + (second j-group)
    assert stats[11] == (nx - 1) * ny * nz           # own-lane terms: i-1 neighbours (incl. seams)
    assert stats[10] == nx * (ny - 2) * nz           # shuffle terms: j-1 neighbours inside a warp
    assert stats[9] == nx * ny * (nz - 1) + nx * nz  # memory terms: k-1 neighbours + the j-1 of each warp's lane 0
    assert stats[12] == stats[0] and stats[13] == stats[0]   # every group takes the fast split-stream path
    # 41 lines = 32 + 9 lanes per layer; (332 + kSkew * (lanes - 1)) steps rounded up to 16 (lanes are skewed by kSkew steps)
    import os, re
    hpp = open(os.path.join(os.path.dirname(__file__), "..", "multiregionfoam_b200", "csrc", "schedule.hpp")).read()
    skew = int(re.search(r"#define B200_SKEW (\d+)", hpp).group(1))  # the default the emulator is built with
    up16 = lambda v: (v + 15) // 16 * 16
    assert stats[7] == nz * 32 * (up16(nx + skew * 31) + up16(nx + skew * 8))
    # only the steps in which some lane crosses one of the two block seams leave the canonical (shuffle, own) form
    assert 0 < stats[14] <= 2 * 2 * 41 * nz and 0 < stats[15] <= 2 * 2 * 41 * nz
    unmerged_slots = int(stats[7])

    # merged: the last warp of a plane is filled with the first lines of a later one (here, gap 0.5: the next one; 41 * 3 = 123
    # lines = 3 full warps + 27 lanes instead of 3 x (32 + 9)): fewer padding slots, the same terms, every group on the fast path
    monkeypatch.delenv("B200_MERGE_SEQ")
    monkeypatch.setenv("B200_MERGE_GAP", "0.5")
    monkeypatch.setenv("B200_MERGE_MIN_GROUPS", "0")   # merging is reserved for systems with many more warps than resident CTAs
    _, _, _, stats = run_emu(emu, reg, False, random_vec(reg.nCells, 1), random_vec(reg.nCells, 2))
    assert stats[8] == 1 and stats[3] == ny * nz
    assert stats[0] == -(-ny * nz // 32)
    assert stats[11] == (nx - 1) * ny * nz
    assert stats[9] + stats[10] == nx * ny * (nz - 1) + nx * (ny - 1) * nz      # k-1 and j-1 neighbours, by memory or by shuffle
    cuts = sum(1 for b in range(32, ny * nz, 32) if b % ny)                      # warp boundaries inside a plane
    assert stats[10] == nx * ((ny - 1) * nz - cuts)                              # j-1 links cut there go through memory
    assert stats[12] == stats[0] and stats[13] == stats[0]
    assert stats[7] < 0.7 * unmerged_slots


def test_default_merge_gap_pairs_planes_further_apart(emu, monkeypatch):
    """Default gap (1.5 levels per warp of the sequence): with 2 warps per plane the partner is 3 planes on; the tables still
    reproduce the oracle bit by bit.  Small systems (fewer warps than 1.5 x the resident CTAs) are not merged at all."""
    fluid, _ = flow_over_heated_plate(1, 12)
    reg = synthetic_coeffs(fluid.nCells, fluid.lowerAddr, fluid.upperAddr, symmetric=False)
    O = pyoracle.OracleSystem(single_region_case(reg))
    x, r = random_vec(reg.nCells, 1), random_vec(reg.nCells, 2)
    monkeypatch.setenv("B200_MERGE_SEQ", "0")
    _, _, _, plain = run_emu(emu, reg, True, x, r)
    monkeypatch.delenv("B200_MERGE_SEQ")
    _, _, _, small = run_emu(emu, reg, True, x, r)
    assert small[0] == plain[0] == 24 and small[7] == plain[7]
    monkeypatch.setenv("B200_MERGE_MIN_GROUPS", "0")
    y, w, rD, stats = run_emu(emu, reg, True, x, r)
    assert stats[0] < plain[0]
    assert stats[7] < plain[7]
    O.precond_setup("DILU")
    assert np.array_equal(y, O.amul(x)) and np.array_equal(rD, O.rD()) and np.array_equal(w, O.precondition(r))


def test_block_mode_on_unstructured_mesh(emu, golden_addr):
    reg = golden_region(golden_addr, "duineveld1", symmetric=True)
    _, _, _, stats = run_emu(emu, reg, False, random_vec(reg.nCells, 1), random_vec(reg.nCells, 2))
    assert stats[8] == 0 and stats[2] == 9
