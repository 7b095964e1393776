"""CPU tests of the conformal-interface maps (directMap variant): oracle (the reference's N^2 loop,
directMapInterfaceToInterfaceMapping.C:155-168, restated in oracle/ldu_oracle.c) against known answers, and the kernel's
arithmetic + tiling / early-exit control flow (csrc/direct_map.hpp via tests/cpp/direct_map_emulate.cpp) against the oracle:
index work, so everything is exact.  GPU leg: tests/test_gpu_zz_direct_map.py."""
import ctypes as C

import numpy as np
import pytest

from ggi_helpers import direct_map_cases
from multiregionfoam_b200 import build as b200build
from oracle import pyoracle

CASES = direct_map_cases()


@pytest.fixture(scope="module")
def emu_map():
    L = C.CDLL(b200build.build_direct_map_emulator())
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.emu_direct_map_build.argtypes = [C.c_int32, dp, C.c_int32, dp, C.c_double, ip]

    def run(to, frm, tol):
        t, f = np.ascontiguousarray(to, np.float64), np.ascontiguousarray(frm, np.float64)
        m = np.full(t.shape[0], -7, np.int32)
        n = L.emu_direct_map_build(t.shape[0], t.ctypes.data_as(dp), f.shape[0], f.ctypes.data_as(dp), tol, m.ctypes.data_as(ip))
        return m, n
    return run


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_known_answers_and_emulator(emu_map, name):
    to, frm, tol, expected = CASES[name]
    mo, no = pyoracle.direct_map_build(to, frm, tol)
    if expected is not None:
        assert np.array_equal(mo, expected)
    else:   # a permutation of the same points: a bijection that reproduces the locations
        assert no == 0 and np.array_equal(np.sort(mo), np.arange(len(frm))) and np.array_equal(np.asarray(frm)[mo], to)
    assert no == int((mo < 0).sum())
    me, ne = emu_map(to, frm, tol)
    assert np.array_equal(me, mo) and ne == no


def test_transfer_round_trip():
    to, frm, tol, perm = CASES["permuted_faces"]
    m, _ = pyoracle.direct_map_build(to, frm, tol)
    back, _ = pyoracle.direct_map_build(frm, to, tol)          # the B-to-A map (:276-289)
    f = np.random.default_rng(0).random((len(frm), 3))
    assert np.array_equal(pyoracle.direct_map(back, pyoracle.direct_map(m, f, 3), 3), f)
