"""ctypes binding of oracle/fv_oracle.c (TEST INFRASTRUCTURE ONLY): the T-equation assembly restated operator by
operator.  Only tests/ may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfv_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "fv_oracle.c")
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _lib.fvo_assemble_T.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, C.c_double, C.c_double, C.c_double, dp, dp, dp,
                                        dp, dp, dp, C.c_int, ip, dp, dp, dp, dp, dp, dp]
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def assemble_T(form, l, u, rhoC, rDeltaT, kappa, V, magSf, deltaCoeffs, Told, *, kappaFace=None, phi=None,
               bCells=None, bInt=None, bSrc=None):
    """-> diag, upper, lower, source of the assembled T equation (boundary contributions included)."""
    f64 = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
    i32 = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)
    l, u, V, magSf, deltaCoeffs, Told = i32(l), i32(u), f64(V), f64(magSf), f64(deltaCoeffs), f64(Told)
    kappaFace, phi, bCells, bInt, bSrc = f64(kappaFace), f64(phi), i32(bCells), f64(bInt), f64(bSrc)
    n, nf = V.size, l.size
    nB = 0 if bCells is None else bCells.size
    diag, source = np.empty(n), np.empty(n)
    upper, lower = np.empty(nf), np.empty(nf)
    rc = lib().fvo_assemble_T(int(form), n, nf, _i(l), _i(u), float(rhoC), float(rDeltaT), float(kappa), _d(kappaFace),
                              _d(V), _d(magSf), _d(deltaCoeffs), _d(phi), _d(Told), nB, _i(bCells), _d(bInt), _d(bSrc),
                              _d(diag), _d(upper), _d(lower), _d(source))
    assert rc == 0
    return diag, upper, lower, source
