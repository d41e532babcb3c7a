"""Preconditioned conjugate gradients on device-resident vectors.

Drop-in for `PCG` of /root/reference/source/linalg.py:6-42: same arguments,
same recurrences and the same absolute stopping test r.z < eps^2, same return
value `(w, iters)` and callback `(w, r, k)`.  For KronVectorMPI operands the
two updates w += alpha p, r -= alpha t are one fused pass and p = z + beta p
another (the reference's `alpha * p` temporary does not exist); the dots are
single-pass device reductions followed by one scalar allreduce.
"""
import numpy as np

from ._lib import check, lib, ptr, stream
from .mpi_vector import KronVectorMPI


def PCG(T, P, b, w0=None, kmax=100000, eps=1e-6, callback=None):
    device = isinstance(b, KronVectorMPI)
    if w0 is None:
        w = KronVectorMPI(b.dofs_distr) if device else np.zeros(b.shape)
    else:
        w = w0
    iters = 0
    if b.dot(b) == 0:
        return w, iters
    # r = b - T w (linalg.py:20).  From the zero start the operator is linear in
    # w, so T w is exactly zero and the apply is skipped.
    r = b.copy() if w0 is None else b - T @ w
    p = P @ r
    abs_r = r.dot(p)
    if abs_r < eps * eps:
        return w, iters
    for k in range(1, kmax):
        iters += 1
        t = T @ p
        alpha = abs_r / p.dot(t)
        if device:
            w._invalidate()
            r._invalidate()
            check(lib().stk_pcg_update(float(alpha), ptr(p.data), ptr(t.data),
                                       ptr(w.data), ptr(r.data), w.numel,
                                       stream()))
        else:
            w += alpha * p
            r -= alpha * t
        del t
        if callback is not None:
            callback(w, r, k)
        z = P @ r
        abs_r_old = abs_r
        abs_r = r.dot(z)
        if abs_r < eps * eps:
            break
        beta = abs_r / abs_r_old
        if device:
            p._invalidate()
            check(lib().stk_xpay(ptr(z.data), float(beta), ptr(p.data),
                                 p.numel, stream()))
        else:
            p *= beta
            p += z
        del z
    return w, iters
