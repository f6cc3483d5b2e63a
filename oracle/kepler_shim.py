"""ORACLE — test infrastructure.  ctypes front end of `libemp_oracle.so`.

`solve(M, ecc)` has the signature of `kepler.solve` from kepler.py 0.0.7 (the
call the reference's templates make, e.g. support/models/kep00.model:6), so
this module can be registered as `sys.modules['kepler']` when a script emitted
by the real reference generator is executed (tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libemp_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "kepler_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libemp_oracle.so"])
    return _LIB_PATH


def _load():
    build()
    lib = ctypes.CDLL(_LIB_PATH)
    lib.emp_oracle_kepler_solve.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    lib.emp_oracle_kepler_solve.restype = None
    lib.emp_oracle_kepler_solve1.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.emp_oracle_kepler_solve1.restype = ctypes.c_double
    lib.emp_oracle_kepler_starter.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.emp_oracle_kepler_starter.restype = ctypes.c_double
    return lib


_lib = _load()


def solve(M, ecc):
    """E = kepler.solve(M, ecc): eccentric anomaly, elementwise, float64."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    ecc = np.ascontiguousarray(np.broadcast_to(np.asarray(ecc, dtype=np.float64), M.shape))
    E = np.empty_like(M)
    _lib.emp_oracle_kepler_solve(M.ctypes.data, ecc.ctypes.data, E.ctypes.data, M.size)
    return E


def starter(M, ecc):
    return _lib.emp_oracle_kepler_starter(float(M), float(ecc))
