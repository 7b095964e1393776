"""GPU tests of b200_direct_map_build / b200_direct_map_transfer through the C ABI against the oracle's restatement of the
reference's N^2 search (directMapInterfaceToInterfaceMapping.C:155-168): index work -> EXACT."""
import numpy as np
import pytest

from ggi_helpers import direct_map_cases, face_centres, grid_patch, min_edge_length
from multiregionfoam_b200 import ldu
from oracle import pyoracle

pytestmark = pytest.mark.gpu

CASES = direct_map_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_map_matches_oracle(gpu_ctx, name):
    to, frm, tol, expected = CASES[name]
    mo, no = pyoracle.direct_map_build(to, frm, tol)
    n0 = gpu_ctx.launches
    mg = gpu_ctx.direct_map_build(to, frm, tol, require_conformal=False)
    assert np.array_equal(mg, mo)
    if expected is not None:
        assert np.array_equal(mg, expected)
    if len(to):
        assert gpu_ctx.launches == n0 + 1
    if no:
        with pytest.raises(ldu.B200Error):          # the reference's "not conformal" FatalError
            gpu_ctx.direct_map_build(to, frm, tol)


def test_large_conformal_interface_and_transfer(gpu_ctx):
    """An interface of the size of C2's (13 200 faces would do; 40 000 here): the reference's search is 1.6e9 distance
    tests on one core; known answer = the permutation; face and point fields go across and back unchanged."""
    faces, pts = grid_patch(np.linspace(0, 4, 201), np.linspace(0, 1, 201))
    cA = face_centres(faces, pts)
    tol = 0.001 * min_edge_length(faces[:50], pts)     # uniform grid: every edge has the same length
    perm = np.random.default_rng(1).permutation(len(faces))
    cB = cA[perm]
    aToB = gpu_ctx.direct_map_build(cB, cA, tol)
    bToA = gpu_ctx.direct_map_build(cA, cB, tol)
    assert np.array_equal(aToB, perm) and np.array_equal(bToA, np.argsort(perm))
    f = np.random.default_rng(2).random((len(faces), 3))
    onB = gpu_ctx.direct_map_transfer(aToB, f)
    assert np.array_equal(onB, pyoracle.direct_map(aToB, f, 3)) and np.array_equal(onB, f[perm])
    assert np.array_equal(gpu_ctx.direct_map_transfer(bToA, onB), f)
    with pytest.raises(ldu.B200Error):                  # a map with a -1 cannot be used
        gpu_ctx.direct_map_transfer(np.array([0, -1], np.int32), f)
