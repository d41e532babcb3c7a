"""Distributed space-time vector, device resident.

Drop-in for /root/reference/source/mpi_vector.py: `DofDistributionMPI`
(:5-38) and `KronVectorMPI` (:41-240) keep their names, constructor
arguments, attributes and operators.  The storage is not a NumPy `(n_t, M)`
array but a block in B200 HBM in the time-fastest layout of include/stk.h;
`X_loc` is a transposing host view of it kept for API parity and tests.
All arithmetic runs in libstk (no CPU fallback).
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream


def pitch(n_t):
    """Row pitch (in doubles) of a block with n_t local time slices."""
    return max(4, (int(n_t) + 3) // 4 * 4)


class DofDistributionMPI:
    """Contiguous time-slab partition (mpi_vector.py:5-38): N//size slices
    per rank, the N%size leftover slices go to the LAST ranks."""
    def __init__(self, comm, N, M):
        self.N = int(N)
        self.M = int(M)
        self.comm = comm
        self.rank = comm.Get_rank()
        self.size = comm.Get_size()
        assert self.N >= self.size
        block, rest = divmod(self.N, self.size)
        self.dof_distribution = []
        self.displs = np.empty(self.size)
        self.counts = np.empty(self.size)
        t = 0
        for p in range(self.size):
            n = block + (1 if self.size - p - 1 < rest else 0)
            self.displs[p] = t * self.M
            self.counts[p] = n * self.M
            self.dof_distribution.append([t, t + n])
            t += n
        assert t == self.N
        self.t_begin, self.t_end = self.dof_distribution[self.rank]
        self.dof2proc = np.zeros(self.N)
        for p, (a, b) in enumerate(self.dof_distribution):
            self.dof2proc[a:b] = p

    @property
    def n_loc(self):
        return self.t_end - self.t_begin


_scratch = {}


def _dot_buffers(device):
    key = ('dot', device)
    if key not in _scratch:
        _scratch[key] = (torch.zeros(2048, dtype=torch.float64, device=device),
                         torch.zeros(1, dtype=torch.float64, device=device))
    return _scratch[key]


def _device():
    if not torch.cuda.is_available():
        raise _lib.StkError('no CUDA device: the space-time solve path has no '
                            'CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


class _XLocView:
    """Host-side `(n_t, M)` float64 view of the device block; reading
    downloads, writing uploads (tests do `vec.X_loc[:] = ...`)."""
    def __init__(self, vec):
        self._vec = vec

    @property
    def shape(self):
        return (self._vec.n_loc, self._vec.M)

    dtype = np.dtype(np.float64)
    ndim = 2

    def __array__(self, dtype=None, copy=None):
        return self._vec.to_host()

    def __getitem__(self, key):
        return self._vec.to_host()[key]

    def __setitem__(self, key, value):
        full = (isinstance(key, slice) and key == slice(None)) or key is Ellipsis
        if full:
            host = np.empty(self.shape)
            host[...] = np.asarray(value, dtype=np.float64).reshape(
                self.shape) if np.ndim(value) else value
        else:
            host = self._vec.to_host()
            host[key] = value
        self._vec.from_host(host)

    def copy(self):
        return self._vec.to_host()

    def reshape(self, *shape):
        return self._vec.to_host().reshape(*shape)

    def __len__(self):
        return self._vec.n_loc


class KronVectorMPI:
    """Vector of N x M space-time dofs, sharded by time slab
    (mpi_vector.py:41-240), resident on this rank's GPU."""
    def __init__(self, dofs_distr, initial_data=None, _data=None):
        self.dofs_distr = dofs_distr
        self.t_begin = dofs_distr.t_begin
        self.t_end = dofs_distr.t_end
        self.N = dofs_distr.N
        self.M = dofs_distr.M
        self.rank = dofs_distr.rank
        self.n_loc = self.t_end - self.t_begin
        self.ld = pitch(self.n_loc)
        self._halo = {}
        if _data is not None:
            self.data = _data
        else:
            self.data = None
            self.reset(initial_data)

    # -- storage ----------------------------------------------------------
    @property
    def numel(self):
        return self.M * self.ld

    def reset(self, initial_data=None):
        """Zero, or a copy of `initial_data` (n_t, M) (mpi_vector.py:62-71)."""
        self._invalidate()
        if self.data is None:
            self.data = torch.zeros((self.M, self.ld), dtype=torch.float64,
                                    device=_device())
            if initial_data is None:
                return
        if initial_data is None:
            self.data.zero_()
        else:
            self.from_host(initial_data)

    def _invalidate(self):
        """Any write drops the cached halo slices (mpi_vector.py:77-82)."""
        self._halo = {}

    @property
    def X_loc(self):
        return _XLocView(self)

    @X_loc.setter
    def X_loc(self, value):
        self.from_host(value)

    def from_host(self, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.shape == (self.n_loc, self.M), (host.shape, self.n_loc,
                                                    self.M)
        self._invalidate()
        staging = torch.empty(self.n_loc * self.M, dtype=torch.float64,
                              device=self.data.device)
        check(lib().stk_block_upload_host(host.ctypes.data, self.n_loc, self.M,
                                          ptr(self.data), self.ld,
                                          ptr(staging), stream()))

    def to_host(self, out=None):
        host = np.empty((self.n_loc, self.M)) if out is None else out
        assert host.flags.c_contiguous and host.shape == (self.n_loc, self.M)
        staging = torch.empty(self.n_loc * self.M, dtype=torch.float64,
                              device=self.data.device)
        check(lib().stk_block_download_host(ptr(self.data), self.ld,
                                            self.n_loc, self.M,
                                            host.ctypes.data, ptr(staging),
                                            stream()))
        return host

    def set_kron(self, u_t_loc, u_x):
        """self = u_t_loc (x) u_x, i.e. X_loc = np.kron(u_t_loc, u_x) reshaped
        (heateq_mpi.py:189-191), formed on the device from the two factors."""
        assert len(u_t_loc) == self.n_loc and len(u_x) == self.M
        self._invalidate()
        dev = self.data.device
        ut = torch.zeros(self.ld, dtype=torch.float64, device=dev)
        ut[:self.n_loc] = torch.as_tensor(np.asarray(u_t_loc, dtype=np.float64))
        ux = torch.as_tensor(np.asarray(u_x, dtype=np.float64)).to(dev)
        check(lib().stk_outer(self.M, self.ld, ptr(ux), ptr(ut), ptr(self.data),
                              stream()))
        return self

    def copy(self):
        return KronVectorMPI(self.dofs_distr, _data=self.data.clone())

    def empty_like(self):
        """Uninitialised vector of the same shape (every libstk kernel writes
        the pads itself)."""
        return KronVectorMPI(self.dofs_distr,
                             _data=torch.empty_like(self.data))

    # -- BLAS-1 (mpi_vector.py:84-122) -----------------------------------
    def axpy(self, a, other):
        """self += a * other without the reference's temporary."""
        self._invalidate()
        check(lib().stk_axpy(float(a), ptr(other.data), ptr(self.data),
                             self.numel, stream()))
        return self

    def __iadd__(self, other):
        return self.axpy(1.0, other)

    def __isub__(self, other):
        return self.axpy(-1.0, other)

    def __imul__(self, a):
        self._invalidate()
        check(lib().stk_scale(float(a), ptr(self.data), self.numel, stream()))
        return self

    def __itruediv__(self, a):
        return self.__imul__(1.0 / float(a))

    def __add__(self, other):
        out = self.copy()
        out += other
        return out

    def __sub__(self, other):
        out = self.copy()
        out -= other
        return out

    def __rmul__(self, a):
        out = self.copy()
        out *= a
        return out

    __mul__ = __rmul__

    def __truediv__(self, a):
        out = self.copy()
        out /= a
        return out

    def __neg__(self):
        return self.__rmul__(-1.0)

    def dot_device(self, other):
        """Local dot product as a device scalar (no host sync)."""
        ws, out = _dot_buffers(self.data.device)
        check(lib().stk_dot(ptr(self.data), ptr(other.data), self.numel,
                            ptr(ws), ptr(out), stream()))
        return out

    def dot_global_device(self, other):
        """Global dot product as a one-element device tensor of its own: the
        local reduction, then the scalar allreduce, nothing read back."""
        out = self.dot_device(other).clone()
        if self.dofs_distr.size > 1:
            out = self.dofs_distr.comm.allreduce_sum(out)
        return out

    def dot(self, other):
        """Global dot product (mpi_vector.py:205-210): fused single-pass local
        reduction, then one allreduce of the device scalar over NCCL."""
        return float(self.dot_global_device(other).item())

    # -- root <-> slabs (mpi_vector.py:124-138; tests/as_global_matrix) ---
    def scatter(self, X_glob):
        comm = self.dofs_distr.comm
        X_glob = comm.bcast(
            None if X_glob is None else np.asarray(X_glob, dtype=np.float64))
        self.from_host(
            X_glob.reshape(self.N, self.M)[self.t_begin:self.t_end])

    def gather(self, X_glob):
        parts = self.dofs_distr.comm.gather(self.to_host())
        if self.rank == 0:
            np.asarray(X_glob).reshape(self.N, self.M)[...] = np.concatenate(
                parts, axis=0)

    # -- time <-> space transpose (mpi_vector.py:212-240) -----------------
    def permute(self):
        """The same vector with the Kronecker factors swapped: an (M x N)
        vector sharded over the FIRST (space) index.  One all-to-all over
        NVLink; see permute.py."""
        from .permute import permute_vector
        return permute_vector(self)

    def communicate_dofs(self, dofs):
        """Fetch the time slices with the given GLOBAL indices that live on
        other ranks (mpi_vector.py:189-203; every rank must call it with the
        needs of its own rows of a symmetric-pattern time matrix, passed as
        (row, col) pairs or as a plain list of columns).  Returns
        {global index: device row of M doubles}."""
        import scipy.sparse as sp
        from .timeop import TimeOpPlan
        cols = sorted({int(d[1]) if np.ndim(d) else int(d) for d in dofs})
        cols = [c for c in cols if not self.t_begin <= c < self.t_end]
        # the exchange pattern must be known to both sides: gather the needs
        all_needs = self._gather_needs(cols)
        rows, cs = [], []
        for p, need in enumerate(all_needs):
            a, _ = self.dofs_distr.dof_distribution[p]
            rows += [a] * len(need)
            cs += list(need)
        T = sp.csr_matrix((np.ones(len(rows)), (rows, cs)),
                          shape=(self.N, self.N))
        plan = TimeOpPlan(self.dofs_distr, T)
        halo = plan.fetch(self)
        return {int(c): halo[k] for k, c in enumerate(plan.halo_cols)}

    def _gather_needs(self, cols):
        comm = self.dofs_distr.comm
        if self.dofs_distr.size == 1:
            return [cols]
        gathered = comm.gather(cols)
        return comm.bcast(gathered)

    # -- halo (mpi_vector.py:140-203) -------------------------------------
    def communicate_bdr(self, callback=None):
        """Fetch the last slice of the previous rank and the first slice of the
        next one (cached until the next write); `callback` runs while the
        transfer is in flight.  Returns (prev, next) device rows or None."""
        from .timeop import neighbour_plan
        plan = neighbour_plan(self.dofs_distr)
        halo = plan.fetch(self, callback)
        prev = nxt = None
        k = 0
        if self.t_begin > 0:
            prev, k = halo[0], 1
        if self.t_end < self.N:
            nxt = halo[k]
        return prev, nxt
