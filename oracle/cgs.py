"""ORACLE / TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/gs.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboraclegs.so')
_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, 'gs.c')
        if (not os.path.exists(_SO)
                or os.path.getmtime(_SO) < os.path.getmtime(src)):
            build()
        _lib = ctypes.CDLL(_SO)
        i32p = np.ctypeslib.ndpointer(np.int32, flags='C')
        f64p = np.ctypeslib.ndpointer(np.float64, flags='C')
        _lib.gs_sweeps.argtypes = [
            ctypes.c_int, i32p, i32p, f64p, f64p, f64p, f64p, ctypes.c_int,
            ctypes.c_int
        ]
        _lib.gs_sweeps_multi.argtypes = [
            ctypes.c_int, ctypes.c_int, i32p, i32p, f64p, f64p, f64p, f64p,
            ctypes.c_int, ctypes.c_int
        ]
        _lib.csr_matvec.argtypes = [ctypes.c_int, i32p, i32p, f64p, f64p, f64p]
    return _lib


def _invdiag(indptr, indices, data):
    n = len(indptr) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    d = np.zeros(n)
    mask = rows == indices
    d[rows[mask]] = data[mask]
    return d**-1


def gauss_seidel(indptr, indices, data, f, u, its, backward=False,
                 invdiag=None):
    """In-place `its` lexicographic GS sweeps on u (n,) or u (nvec, n)."""
    if invdiag is None:
        invdiag = _invdiag(indptr, indices, data)
    n = len(indptr) - 1
    f = np.ascontiguousarray(f, dtype=np.float64)
    assert u.flags.c_contiguous and u.dtype == np.float64
    if u.ndim == 1:
        lib().gs_sweeps(n, indptr, indices, data, invdiag, f, u, its,
                        int(backward))
    else:
        assert u.shape[1] == n and f.shape == u.shape
        lib().gs_sweeps_multi(u.shape[0], n, indptr, indices, data, invdiag,
                              f, u, its, int(backward))
    return u
