"""Condition-number estimate of P A by preconditioned Lanczos.

Drop-in for `Lanczos` of /root/reference/source/lanczos.py:9-171 (same
constructor, `.lmax/.lmin/.iterations/.converged/.cond()`).  The vector work
(`A @`, `P @`, dots, axpys) runs wherever the operands live -- on the GPU for
KronVectorMPI / MPI operators -- while the Sturm-sequence bisection on the
small tridiagonal matrix is host scalar code, as in the reference.
"""
import time
from math import sqrt

import numpy as np
import scipy.sparse as sp


class Lanczos:
    MAXLANCZOS = 2000
    TOLBISEC = 0.000001
    TOL = 0.0001

    def _charpoly(self, k, x):
        """Characteristic polynomial of the leading (k+1) x (k+1) Lanczos
        matrix at x by the three-term recurrence (lanczos.py:77-85)."""
        prev, cur = 1.0, self.alpha[0] - x
        for l in range(1, k + 1):
            prev, cur = cur, (self.alpha[l] -
                              x) * cur - self.beta[l - 1]**2 * prev
        return cur

    def bisec(self, k, ymax, zmin, tol):
        """Tighten [ymax, zmax] around the largest and [ymin, zmin] around the
        smallest eigenvalue of the (k+1) x (k+1) matrix (lanczos.py:20-75).
        The outer brackets come from Gershgorin discs."""
        a, b = self.alpha, self.beta
        zmax = a[0] + abs(b[0])
        ymin = a[0] - abs(b[0])
        for l in range(1, k):
            zmax = max(zmax, a[l] + abs(b[l - 1]) + abs(b[l]))
            ymin = min(ymin, a[l] - abs(b[l - 1]) - abs(b[l]))
        zmax = max(zmax, a[k] + abs(b[k - 1]))
        ymin = max(min(ymin, a[k] - abs(b[k - 1])), 0.0)

        neg = np.signbit
        pz = self._charpoly(k, zmax)
        while abs(zmax - ymax) > tol * min(abs(zmax), abs(ymax)):
            x = 0.5 * (ymax + zmax)
            px = self._charpoly(k, x)
            if neg(px) != neg(pz):
                ymax = x
            else:
                zmax, pz = x, px
        py = self._charpoly(k, ymax)
        if neg(pz) != neg(py) and py != 0:
            ymax = zmax

        py = self._charpoly(k, ymin)
        while abs(zmin - ymin) > tol * min(abs(zmin), abs(ymin)):
            x = 0.5 * (ymin + zmin)
            px = self._charpoly(k, x)
            if neg(px) != neg(py):
                zmin = x
            else:
                ymin, py = x, px
        pz = self._charpoly(k, zmin)
        if neg(pz) != neg(py) and pz != 0:
            zmin = ymin
        return ymax, zmin

    def __init__(self, A, P=None, w=None, maxIterations=MAXLANCZOS, tol=TOL,
                 tolBisec=TOLBISEC):
        self.alpha = np.zeros(maxIterations)
        self.beta = np.zeros(maxIterations - 1)
        self.converged = True
        if P is None:
            P = sp.identity(A.shape[0])
        start = time.process_time()
        if w is None:
            w = 2.0 * np.random.rand(A.shape[0]) - 1.0

        v = A @ w
        nrm = sqrt(v.dot(w))
        v /= nrm
        w /= nrm
        v = P @ v
        self.alpha[0] = (A @ v).dot(w)
        lmax = lmin = self.alpha[0]
        k = 0
        while True:
            if k == maxIterations - 1:
                self.converged = False
                break
            v -= self.alpha[k] * w
            self.beta[k] = sqrt((A @ v).dot(v))
            w, v = v / self.beta[k], -self.beta[k] * w
            v += P @ (A @ w)
            k += 1
            self.alpha[k] = (A @ v).dot(w)
            lmax_old, lmin_old = lmax, lmin
            lmax, lmin = self.bisec(k, lmax, lmin, tolBisec)
            if (lmax - lmax_old) < tol * lmax_old and (lmin_old -
                                                       lmin) < tol * lmin:
                break
        self.iterations = k + 1
        self.time = time.process_time() - start
        self.lmax, self.lmin = lmax, lmin
        self.alpha = np.resize(self.alpha, k)
        self.beta = np.resize(self.beta, k - 1)

    def cond(self):
        return self.lmax / self.lmin

    def __str__(self):
        return '{}\tits={}\tlmax={}\tlmin={}\tkappa={}\ttime={} s'.format(
            'converged' if self.converged else 'NOT converged',
            self.iterations, self.lmax, self.lmin, self.cond(), self.time)
