"""ORACLE / TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's
Kronecker space-time solve path.

Not product code: only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import this module.  The product path never does, and fails
loudly when its CUDA library is missing.

Parity status: PINNED.  The functions below are checked (tests/test_oracle.py)
against
  * the known answers of the reference's own tests (the 5x5 interleaved W for
    J=2, the level-0/1/2 wavelet shapes, prod(I+split_j) == W, the 5x5
    tridiagonal np.kron fixture, permute == reshape.T), and
  * outputs of the UNMODIFIED reference classes run in the build container
    through oracle/ref_harness.py, committed as tests/golden/*.npz by
    oracle/gen_golden.py (operator applies, PCG iteration counts, residual
    histories, Lanczos condition numbers).
The only unpinned boundary is NGSolve's assembly (absent from this image), so
matrices come from the product's host assembler and are fed identically to the
oracle and to the GPU path.

Conventions: a space-time vector is a dense (N, M) float64 array, row t = time
node t (the global view of the reference's X_loc slabs).  Operators act on it
without regard to the MPI decomposition, which does not change the algebra.
"""
from math import sqrt

import numpy as np
import scipy.sparse as sp

from . import cgs


# ---------------------------------------------------------------------------
# distribution  (source/mpi_vector.py:5-38)
# ---------------------------------------------------------------------------
def slab_bounds(N, P):
    """[t_begin, t_end) per rank: N//P slices each, the N%P leftover slices go
    to the LAST ranks (mpi_vector.py:18-31)."""
    assert N >= P  # mpi_vector.py:15
    base, rest = divmod(N, P)
    bounds, t = [], 0
    for p in range(P):
        n = base + (1 if P - p - 1 < rest else 0)
        bounds.append((t, t + n))
        t += n
    assert t == N
    return bounds


def dot(X, Y, P=1):
    """Global dot = sum over ranks of the local np.dot (mpi_vector.py:205-210)."""
    N = X.shape[0]
    return sum(
        float(np.dot(X[a:b].reshape(-1), Y[a:b].reshape(-1)))
        for a, b in slab_bounds(N, P))


def permute(X):
    """Kron-order swap (N, M) -> (M, N) (mpi_vector.py:212-240)."""
    return np.ascontiguousarray(X.T)


# ---------------------------------------------------------------------------
# Kronecker applies  (source/mpi_kron.py)
# ---------------------------------------------------------------------------
def apply_space(op, X):
    """(I (x) op) X : `op` acts on every time slice (mpi_kron.py:143-150).
    `op` is anything supporting `op @ (M, k) array` (CSR, ndarray, callable)."""
    if callable(op):
        return op(X)
    return np.ascontiguousarray((op @ X.T).T)


def apply_time(T, X):
    """(T (x) I) X (mpi_kron.py:186-201, 240-256, 285-317)."""
    return np.asarray(T @ X)


def kron_apply(T, op, X):
    """(T (x) op) X as TridiagKronMatMPI does it: time first, then space in
    place (mpi_kron.py:214-219)."""
    return apply_space(op, apply_time(T, X))


def chain(*ops):
    """x -> ops[0] ops[1] ... ops[-1] x on slice blocks (linop.py:68-79)."""
    def run(X):
        for op in reversed(ops):
            X = apply_space(op, X)
        return X
    return run


# ---------------------------------------------------------------------------
# wavelets  (source/wavelets.py:45-169)
# ---------------------------------------------------------------------------
def wavelet_levels(J, interleaved=True):
    """Level of the wavelet attached to each index (wavelets.py:70-79)."""
    if interleaved:
        lv = np.zeros(2**J + 1, dtype=int)
        for j in range(J, -1, -1):
            lv[::2**(J - j)] = j
        return lv
    return np.array([0, 0] + [j for j in range(1, J + 1)
                              for _ in range(2**(j - 1))])


def _synth_level(c, d, j):
    """[p(j) q(j)] (c; d): coarse hats c (2^(j-1)+1 rows) and level-j wavelets
    d (2^(j-1) rows) -> hats of level j (wavelets.py:81-104)."""
    s = 2.0**(j / 2)
    fine = np.empty((c.shape[0] + d.shape[0], ) + c.shape[1:])
    ev = c.copy()
    ev[:-1] -= 0.5 * s * d
    ev[1:] -= 0.5 * s * d
    ev[0] -= 0.5 * s * d[0]  # boundary wavelets: -1 instead of -1/2
    ev[-1] -= 0.5 * s * d[-1]
    fine[0::2] = ev
    fine[1::2] = 0.5 * (c[:-1] + c[1:]) + s * d
    return fine


def _analysis_level(y, j):
    """[p(j) q(j)]^T y (wavelets.py:67-68,120-134)."""
    s = 2.0**(j / 2)
    ev, od = y[0::2], y[1::2]
    c = ev.copy()
    c[:-1] += 0.5 * od
    c[1:] += 0.5 * od
    d = od - 0.5 * ev[:-1] - 0.5 * ev[1:]
    d[0] -= 0.5 * ev[0]
    d[-1] -= 0.5 * ev[-1]
    return c, s * d


def wavelet_synthesis(X, J, interleaved=True):
    """W X along axis 0: wavelet -> hat coordinates (wavelets.py:106-118)."""
    Y = np.array(X, dtype=np.float64, copy=True)
    for j in range(1, J + 1):
        if interleaved:
            S = 2**(J - j)
            Y[::S] = _synth_level(Y[::2 * S], Y[S::2 * S], j)
        else:
            nc, nf = 2**(j - 1) + 1, 2**j + 1
            Y[:nf] = _synth_level(Y[:nc], Y[nc:nf], j)
    return Y


def wavelet_analysis(X, J, interleaved=True):
    """W^T X along axis 0 (wavelets.py:120-134)."""
    Y = np.array(X, dtype=np.float64, copy=True)
    for j in range(J, 0, -1):
        if interleaved:
            S = 2**(J - j)
            c, d = _analysis_level(Y[::S], j)
            Y[::2 * S], Y[S::2 * S] = c, d
        else:
            nc, nf = 2**(j - 1) + 1, 2**j + 1
            c, d = _analysis_level(Y[:nf], j)
            Y[:nc], Y[nc:nf] = c, d
    return Y


def wavelet_split(J, j):
    """The level-j step of the interleaved transform minus the identity on the
    nodes it touches, so that W = prod_j (I + split_j) (wavelets.py:136-169)."""
    n = 2**J + 1
    E = np.eye(n)
    S = 2**(J - j)
    out = E.copy()
    out[::S] = _synth_level(E[::2 * S], E[S::2 * S], j)
    return sp.csr_matrix(out - E)


# ---------------------------------------------------------------------------
# multigrid  (source/multigrid.py:130-197)
# ---------------------------------------------------------------------------
class MultiGridOracle:
    """V-cycle preconditioner on slice blocks X of shape (k, M)."""
    def __init__(self, mat, P_mats, smoothsteps=2, vcycles=1, threads=1):
        self.nu, self.vcycles = smoothsteps, vcycles
        self.threads = threads
        self.P = [sp.csr_matrix(P) for P in P_mats]
        self.R = [P.T.tocsr() for P in self.P]  # multigrid.py:60
        self.mats = [sp.csr_matrix(mat)]
        for j in reversed(range(len(self.P))):  # multigrid.py:140-145
            self.mats.insert(0, (self.R[j] @ self.mats[0] @ self.P[j]).tocsr())
        for A in self.mats:
            A.sort_indices()
        self.csr = [(A.indptr.astype(np.int32), A.indices.astype(np.int32),
                     A.data.astype(np.float64)) for A in self.mats]
        self.invdiag = [A.diagonal()**-1 for A in self.mats]
        # coarsest level: exact solve (multigrid.py:161-165,170)
        self.coarse_inv = np.linalg.inv(self.mats[0].toarray())
        self.shape = self.mats[-1].shape

    def _smooth(self, j, U, F, backward):
        cgs.gauss_seidel(*self.csr[j], F, U, self.nu, backward=backward,
                         invdiag=self.invdiag[j])

    def _cycle(self, j, U, F):
        """MGM (multigrid.py:168-182), all slices of the block at once."""
        if j == 0:
            U[:] = F @ self.coarse_inv.T
            return
        self._smooth(j, U, F, backward=False)
        D = apply_space(self.R[j - 1], apply_space(self.mats[j], U) - F)
        Uc = np.zeros_like(D)
        self._cycle(j - 1, Uc, D)
        U -= apply_space(self.P[j - 1], Uc)
        self._smooth(j, U, F, backward=True)

    def __call__(self, X):
        """(k, M) -> (k, M): `vcycles` V-cycles from a zero initial guess
        (multigrid.py:184-193)."""
        F = np.ascontiguousarray(X, dtype=np.float64)
        if self.threads > 1 and F.shape[0] > 1:
            # the slices are independent (the reference's ranks each own a
            # slab of them): one chunk of slices per host thread
            from concurrent.futures import ThreadPoolExecutor
            chunks = np.array_split(np.arange(F.shape[0]),
                                    min(self.threads, F.shape[0]))
            with ThreadPoolExecutor(len(chunks)) as pool:
                parts = list(pool.map(lambda c: self._solve(F[c]), chunks))
            return np.concatenate(parts, axis=0)
        return self._solve(F)

    def _solve(self, F):
        F = np.ascontiguousarray(F)
        U = np.zeros_like(F)
        for _ in range(self.vcycles):
            self._cycle(len(self.mats) - 1, U, F)
        return U

    def __matmul__(self, B):
        """SciPy-style (M,) or (M, k) interface."""
        B = np.asarray(B, dtype=np.float64)
        if B.ndim == 1:
            return self(B[None, :])[0]
        return np.ascontiguousarray(self(np.ascontiguousarray(B.T)).T)


class InvOracle:
    """Direct inverse on slice blocks X of shape (k, M): SuperLU with the
    options of source/linop.py:18-26 (`InvLinOp`, precond='direct')."""
    def __init__(self, mat):
        import scipy.sparse.linalg as spla
        self.lu = spla.splu(sp.csc_matrix(mat),
                            options={"SymmetricMode": True},
                            permc_spec="MMD_AT_PLUS_A")
        self.shape = mat.shape

    def __call__(self, X):
        return np.ascontiguousarray(
            self.lu.solve(np.ascontiguousarray(np.asarray(X).T)).T)

    def __matmul__(self, B):
        return self.lu.solve(np.asarray(B, dtype=np.float64))


# ---------------------------------------------------------------------------
# the operator graph of heateq_mpi.py:126-191
# ---------------------------------------------------------------------------
class HeatEqOracle:
    def __init__(self, prob, smoothsteps=3, vcycles=2, interleaved=True,
                 threads=1, precond='multigrid'):
        p = self.prob = prob
        P_mats = p.hierarchy.P_mats
        self.J = p.J_time
        self.interleaved = interleaved
        # `threads` > 1: the time slices are split into slabs, one per host
        # thread, exactly like the reference's MPI ranks (mpi_vector.py:18-31);
        # the C kernels and SciPy's sparse products release the GIL.
        self.threads = threads
        if precond == 'multigrid':  # heateq_mpi.py:142-153
            self.K = MultiGridOracle(p.A_x, P_mats, smoothsteps, vcycles)
            self.C = [MultiGridOracle(m, P_mats, smoothsteps, vcycles)
                      for m in p.Cinv_j]
        else:  # heateq_mpi.py:154-157
            assert precond == 'direct'
            self.K = InvOracle(p.A_x)
            self.C = [InvOracle(m) for m in p.Cinv_j]
        self.levels = wavelet_levels(self.J, interleaved)
        K, Mx, Ax = self.K, p.M_x, p.A_x
        self.terms = [  # heateq_mpi.py:166-178
            (p.A_t, chain(Mx, K, Mx)),
            (p.L_t, chain(Mx, K, Ax)),
            (p.L_t.T.tocsr(), chain(Ax, K, Mx)),
            (p.M_t, chain(Ax, K, Ax)),
            (p.G_t, chain(Mx)),
        ]
        self.rhs = np.outer(p.u0_t, p.u0_x)  # heateq_mpi.py:189-191

    def W(self, X):
        return wavelet_synthesis(X, self.J, self.interleaved)

    def WT(self, X):
        return wavelet_analysis(X, self.J, self.interleaved)

    def _slabs(self, n):
        P = max(1, min(self.threads, n))
        return [(a, b) for a, b in slab_bounds(n, P)]

    def _parallel(self, fn, n):
        """fn(a, b) on every slab [a, b) of n slices, results concatenated."""
        slabs = self._slabs(n)
        if len(slabs) == 1:
            return fn(0, n)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(slabs)) as pool:
            parts = list(pool.map(lambda ab: fn(*ab), slabs))
        return np.concatenate(parts, axis=0)

    def S_slab(self, TX_slab):
        """The space chains of the five terms on one slab of slices, given the
        slab of each time-applied vector (what one MPI rank does after the
        halo exchange, mpi_kron.py:214-219)."""
        out = np.zeros_like(TX_slab[0])
        for (T, op), tx in zip(self.terms, TX_slab):
            out += apply_space(op, np.ascontiguousarray(tx))
        return out

    def P_slab(self, X_slab, levels_slab):
        """Block-diagonal preconditioner on one slab (mpi_kron.py:122-132)."""
        out = np.empty_like(X_slab)
        for j in np.unique(levels_slab):
            sel = np.nonzero(levels_slab == j)[0]
            C = self.C[j]
            out[sel] = C(apply_space(self.prob.A_x, C(X_slab[sel])))
        return out

    def S(self, X):
        """SumMPI over the five Kronecker terms (mpi_kron.py:77-90): the time
        stencils on the whole vector, the space chains slab by slab."""
        TX = [apply_time(T, X) for T, _ in self.terms]
        return self._parallel(
            lambda a, b: self.S_slab([tx[a:b] for tx in TX]), X.shape[0])

    def WT_S_W(self, X):
        return self.WT(self.S(self.W(X)))  # mpi_kron.py:101-110

    def P(self, X):
        """Block diagonal in time: slice t gets C_j A_x C_j with j = level of
        wavelet t (mpi_kron.py:122-132, heateq_mpi.py:159-162,183-184)."""
        return self._parallel(
            lambda a, b: self.P_slab(X[a:b], self.levels[a:b]), X.shape[0])

    def solve(self, **kw):
        return pcg(self.WT_S_W, self.P, self.rhs, **kw)


# ---------------------------------------------------------------------------
# Krylov  (source/linalg.py:6-42, source/lanczos.py)
# ---------------------------------------------------------------------------
def pcg(T, P, b, w0=None, kmax=100000, eps=1e-6, callback=None, nranks=1):
    """PCG with the reference's absolute stopping test r.z < eps^2.
    T, P are callables on (N, M) arrays.  Returns (w, iters)."""
    ip = lambda x, y: dot(x, y, nranks)
    w = np.zeros_like(b) if w0 is None else w0
    iters = 0
    if ip(b, b) == 0:
        return w, iters
    r = b - T(w)
    p = P(r)
    rz = ip(r, p)
    if rz < eps * eps:
        return w, iters
    for k in range(1, kmax):
        iters += 1
        t = T(p)
        a = rz / ip(p, t)
        w = w + a * p
        r = r - a * t
        if callback is not None:
            callback(w, r, k)
        z = P(r)
        rz_old, rz = rz, ip(r, z)
        if rz < eps * eps:
            break
        p = (rz / rz_old) * p + z
    return w, iters


def _sturm(alpha, beta, k, x):
    """Characteristic polynomial of the leading (k+1)x(k+1) Lanczos
    tridiagonal at x (lanczos.py:77-85)."""
    prev, cur = 1.0, alpha[0] - x
    for l in range(1, k + 1):
        prev, cur = cur, (alpha[l] - x) * cur - beta[l - 1]**2 * prev
    return cur


def _bisect_extremes(alpha, beta, k, ymax, zmin, tol):
    """Refine the brackets of the extreme eigenvalues (lanczos.py:20-75)."""
    hi = alpha[0] + abs(beta[0])
    lo = alpha[0] - abs(beta[0])
    for l in range(1, k):
        hi = max(hi, alpha[l] + abs(beta[l - 1]) + abs(beta[l]))
        lo = min(lo, alpha[l] - abs(beta[l - 1]) - abs(beta[l]))
    hi = max(hi, alpha[k] + abs(beta[k - 1]))
    lo = max(min(lo, alpha[k] - abs(beta[k - 1])), 0.0)
    zmax, ymin = hi, lo
    sgn = np.signbit
    pz = _sturm(alpha, beta, k, zmax)
    while abs(zmax - ymax) > tol * min(abs(zmax), abs(ymax)):
        x = 0.5 * (ymax + zmax)
        px = _sturm(alpha, beta, k, x)
        if sgn(px) != sgn(pz):
            ymax = x
        else:
            zmax, pz = x, px
    py = _sturm(alpha, beta, k, ymax)
    if sgn(pz) != sgn(py) and py != 0:
        ymax = zmax
    py = _sturm(alpha, beta, k, ymin)
    while abs(zmin - ymin) > tol * min(abs(zmin), abs(ymin)):
        x = 0.5 * (ymin + zmin)
        px = _sturm(alpha, beta, k, x)
        if sgn(px) != sgn(py):
            zmin = x
        else:
            ymin, py = x, px
    pz = _sturm(alpha, beta, k, zmin)
    if sgn(pz) != sgn(py) and pz != 0:
        zmin = ymin
    return ymax, zmin


def lanczos(A, P, w, max_iterations=2000, tol=1e-4, tol_bisec=1e-6):
    """Extreme eigenvalues of P A by preconditioned Lanczos
    (lanczos.py:87-159).  A, P callables; w the start vector (copied).
    Returns (lmax, lmin, iterations)."""
    ip = lambda x, y: float(np.dot(x.reshape(-1), y.reshape(-1)))
    alpha = np.zeros(max_iterations)
    beta = np.zeros(max_iterations - 1)
    w = np.array(w, dtype=np.float64, copy=True)
    v = A(w)
    nrm = sqrt(ip(v, w))
    v = v / nrm
    w = w / nrm
    v = P(v)
    alpha[0] = ip(A(v), w)
    lmax = lmin = alpha[0]
    k = 0
    while k < max_iterations - 1:
        v = v - alpha[k] * w
        beta[k] = sqrt(ip(A(v), v))
        w, v = v / beta[k], -beta[k] * w
        v = v + P(A(w))
        k += 1
        alpha[k] = ip(A(v), w)
        lmax_old, lmin_old = lmax, lmin
        lmax, lmin = _bisect_extremes(alpha, beta, k, lmax, lmin, tol_bisec)
        if (lmax - lmax_old) < tol * lmax_old and (lmin_old -
                                                    lmin) < tol * lmin:
            break
    return lmax, lmin, k + 1
