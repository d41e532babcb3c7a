"""Preconditioned conjugate gradients on device-resident vectors.

Drop-in for `PCG` of /root/reference/source/linalg.py:6-42: same arguments,
same recurrences and the same absolute stopping test r.z < eps^2, same return
value `(w, iters)` and callback `(w, r, k)`.  The operands are `KronVectorMPI`s
(the solve path has no host form: NumPy operands are refused).  The two
updates w += alpha p, r -= alpha t are one fused pass and p = z + beta p another
(the reference's `alpha * p` temporary does not exist); the dots are
single-pass device reductions followed by one scalar allreduce.  alpha =
r.z / p.t and beta = r.z / (r.z)_old stay on the device (the update kernels
read the two dot results and divide), so an iteration has ONE host read-back:
r.z for the stopping test, exactly where the reference tests it.
"""
from ._lib import check, lib, ptr, stream
from .mpi_vector import KronVectorMPI


def PCG(T, P, b, w0=None, kmax=100000, eps=1e-6, callback=None):
    if not isinstance(b, KronVectorMPI):
        raise TypeError('PCG: the right-hand side must be a KronVectorMPI (the '
                        'B200 solve path has no NumPy form)')
    w = KronVectorMPI(b.dofs_distr) if w0 is None else w0
    iters = 0
    if b.dot(b) == 0:
        return w, iters
    # r = b - T w (linalg.py:20).  From the zero start the operator is linear in
    # w, so T w is exactly zero and the apply is skipped.
    r = b.copy() if w0 is None else b - T @ w
    p = P @ r
    rz = r.dot_global_device(p)
    if float(rz.item()) < eps * eps:
        return w, iters
    for k in range(1, kmax):
        iters += 1
        t = T @ p
        pt = p.dot_global_device(t)
        w._invalidate()
        r._invalidate()
        # w += alpha p, r -= alpha t with alpha = rz / pt (linalg.py:27-30)
        check(lib().stk_pcg_update_dev(ptr(rz), ptr(pt), ptr(p.data),
                                       ptr(t.data), ptr(w.data), ptr(r.data),
                                       w.numel, stream()))
        del t
        if callback is not None:
            callback(w, r, k)
        z = P @ r
        rz_old = rz
        rz = r.dot_global_device(z)
        if float(rz.item()) < eps * eps:  # the iteration's one read-back
            break
        # p = z + beta p with beta = rz / rz_old (linalg.py:38-40)
        p._invalidate()
        check(lib().stk_xpay_dev(ptr(z.data), ptr(rz), ptr(rz_old),
                                 ptr(p.data), p.numel, stream()))
        del z
    return w, iters
