"""The sweep kernels wait for values of other groups by polling; a dependency that never arrives must end in an error
code, not in a hang ("there is no test that the error path works").  Fault injection: with
B200_SWEEP_DEBUG=4 the first group of every sweep does nothing, so each group that depends on it runs into the polling
limit (B200_SWEEP_SPIN_LIMIT, lowered here so that the test takes milliseconds); the call has to return B200_EDEVICE
with the library's message, and the context has to stay usable for a system built without the fault."""
import os

import numpy as np
import pytest

from helpers import random_vec
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case
from oracle import pyoracle

pytestmark = pytest.mark.gpu
B200_EDEVICE = -7  # include/b200_ldu.h


def test_missing_dependency_times_out_with_an_error(gpu_ctx):
    case = cht_case(1, 5)[0]
    r = random_vec(case.nCells, 3)
    os.environ["B200_SWEEP_DEBUG"] = "4"
    os.environ["B200_SWEEP_SPIN_LIMIT"] = "300"
    try:
        S = ldu.LduSystem(gpu_ctx, case.ranks[0])
    finally:
        del os.environ["B200_SWEEP_DEBUG"], os.environ["B200_SWEEP_SPIN_LIMIT"]
    try:
        with pytest.raises(ldu.B200Error) as ei:
            S.precondition(ldu.PRECOND_DILU, r)
        assert ei.value.code == B200_EDEVICE and "timed out" in str(ei.value)
        # the error word is cleared: the same faulty system reports the time-out again instead of a stale state
        with pytest.raises(ldu.B200Error):
            S.precondition(ldu.PRECOND_DILU, r)
    finally:
        S.close()
    # a system built without the fault on the same context works and is bit-exact
    S2 = ldu.LduSystem(gpu_ctx, case.ranks[0])
    try:
        O = pyoracle.OracleSystem(case)
        O.precond_setup("DILU")
        assert np.array_equal(S2.precondition(ldu.PRECOND_DILU, r), O.precondition(r))
    finally:
        S2.close()
