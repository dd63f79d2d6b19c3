"""ctypes front end of oracle/native_ref.cpp ("oracle-native": the reference's call structure with a blocked OpenMP Gemm,
stand-in for the reference's default pure-Go BLAS backend).  TEST / MEASUREMENT INFRASTRUCTURE: used by tests/ and by
bench.py's CPU legs only."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_native_ref())
        _lib.native_ref_threads.restype = C.c_int
        _lib.native_mlp_step_r.restype = C.c_int
    return _lib


def threads() -> int:
    return int(lib().native_ref_threads())


def set_threads(n: int) -> None:
    lib().native_ref_set_threads(C.c_int(int(n)))


def mlp_step_r(sizes, n, params, rvecs, x):
    """One bench/mlp.go:43-57 iteration (fresh accumulators, outGrad = 1, outRGrad = 0).  params / rvecs: 2L float64
    arrays in the order W0, b0, W1, b1, ...  Returns (out, rout, grads, rgrads)."""
    L = len(sizes) - 1
    P = 2 * L
    params = [np.ascontiguousarray(p, dtype=np.float64) for p in params]
    rvecs = [np.ascontiguousarray(r, dtype=np.float64) for r in rvecs]
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.size == n * sizes[0] and len(params) == P and len(rvecs) == P
    out, rout = np.empty(n * sizes[-1]), np.empty(n * sizes[-1])
    grads = [np.zeros(p.size) for p in params]
    rgrads = [np.zeros(p.size) for p in params]
    PP = C.c_void_p * P
    csizes = (C.c_int64 * len(sizes))(*sizes)
    rc = lib().native_mlp_step_r(csizes, C.c_int(L), C.c_int64(n), PP(*[p.ctypes.data for p in params]),
                                 PP(*[r.ctypes.data for r in rvecs]), C.c_void_p(x.ctypes.data), C.c_void_p(out.ctypes.data),
                                 C.c_void_p(rout.ctypes.data), PP(*[g.ctypes.data for g in grads]),
                                 PP(*[g.ctypes.data for g in rgrads]))
    assert rc == 0
    return out, rout, grads, rgrads
