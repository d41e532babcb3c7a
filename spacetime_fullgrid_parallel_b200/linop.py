"""Space operators on device blocks.

Mirrors the serial operator layer of /root/reference/source/linop.py that the
MPI operators consume: `CompositeLinOp` (:68-79, the M.K.A chains of
heateq_mpi.py:166-178) and the plain CSR matrices passed as `mat_space`.
Everything acts on (M, ld) device blocks -- all local time slices at once --
through libstk; there is no host path.

Protocol of a space operator:
    op.shape                          (M, M)
    op.apply_block(x, out, ctx=None)  out <- op x   (x, out: (M, ld) device
                                      tensors, out is not x)
`ctx` carries per-slice coefficients for operator families (multigrid.py);
plain operators ignore it.
"""
import numpy as np
import scipy.sparse as sp
import torch

from ._lib import check, lib, ptr, stream


def locality_order(pattern):
    """Row order in which neighbouring rows of a (square, structurally
    symmetrised) sparsity pattern are visited close together: reverse
    Cuthill-McKee.  Setup only; any permutation is valid for the kernels."""
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    pat = sp.csr_matrix(pattern)
    n = pat.shape[0]
    if n < 2 or pat.shape[0] != pat.shape[1]:
        return None
    g = sp.csr_matrix((np.ones(pat.nnz, dtype=np.int8), pat.indices,
                       pat.indptr), shape=pat.shape)
    return np.ascontiguousarray(
        reverse_cuthill_mckee((g + g.T).tocsr(), symmetric_mode=True),
        dtype=np.int32)


def row_order_enabled():
    import os
    return os.environ.get('STK_ROW_ORDER', '1') != '0'


class _RowOrder:
    """Keeps a device row schedule registered for a CSR row-pointer array
    (stk_csr_set_row_order) for as long as the owner lives."""
    def __init__(self, indptr_dev, nrows, order, device):
        self.indptr = indptr_dev
        self.order = None
        if order is None or not row_order_enabled():
            return
        order = np.ascontiguousarray(order, dtype=np.int32)
        assert len(order) == nrows
        self.order = torch.from_numpy(order).to(device)
        check(lib().stk_csr_set_row_order(ptr(indptr_dev), int(nrows),
                                          ptr(self.order)))

    def __del__(self):
        try:
            if self.order is not None:
                lib().stk_csr_set_row_order(ptr(self.indptr), 0, None)
        except Exception:
            pass


# below this many rows the whole block stays in L2 whatever the order
ROW_ORDER_MIN_ROWS = 16384


class DeviceCSR:
    """A CSR matrix resident in HBM (fp64 values, int32 indices, as
    mpi_shared_mem.py:46-48 stores them)."""
    def __init__(self, mat, device=None, row_order='auto'):
        mat = sp.csr_matrix(mat, dtype=np.float64)
        mat.sort_indices()
        if device is None:
            from .mpi_vector import _device
            device = _device()
        self.shape = mat.shape
        self.nnz = mat.nnz
        self.host = mat
        self.indptr = torch.from_numpy(mat.indptr.astype(np.int32)).to(device)
        self.indices = torch.from_numpy(mat.indices.astype(np.int32)).to(device)
        self.data = torch.from_numpy(mat.data.astype(np.float64)).to(device)
        self.num_applies = 0
        # row schedule of the SpMM kernels: 'auto' = a locality order for
        # large square matrices, None = index order, or an explicit permutation
        if isinstance(row_order, str):
            row_order = (locality_order(mat)
                         if mat.shape[0] >= ROW_ORDER_MIN_ROWS
                         and row_order_enabled() else None)
        self._row_order = _RowOrder(self.indptr, mat.shape[0], row_order,
                                    device)

    def spmm(self, x, out, alpha=1.0, beta=0.0, z=None):
        """out = alpha * A x + beta * z  on blocks; z may be out."""
        ld = x.shape[1]
        check(lib().stk_space_spmm(self.shape[0], ptr(self.indptr),
                                   ptr(self.indices), 1, ptr(self.data), None,
                                   None, None, ptr(x), float(alpha),
                                   float(beta), ptr(z), ptr(out), ld,
                                   stream()))
        self.num_applies += 1

    def apply_block(self, x, out, ctx=None):
        self.spmm(x, out)

    def toarray(self):
        return self.host.toarray()

    def tocsr(self):
        return self.host

    def __matmul__(self, B):
        """Host arrays (M,) / (M, k), computed on the device."""
        return host_apply(self, B)


class DeviceCSRPair:
    """Two matrices on their union sparsity pattern (M_x and A_x): one index
    structure, two value arrays, so that products with both read the pattern
    and the input block once (stk_space_spmm_split / _pair)."""
    def __init__(self, mat0, mat1, device=None):
        if device is None:
            from .mpi_vector import _device
            device = _device()
        m0 = sp.csr_matrix(mat0.host if isinstance(mat0, DeviceCSR) else mat0,
                           dtype=np.float64)
        m1 = sp.csr_matrix(mat1.host if isinstance(mat1, DeviceCSR) else mat1,
                           dtype=np.float64)
        assert m0.shape == m1.shape
        pat = (abs(m0) + abs(m1)).tocsr()
        pat.sort_indices()
        n = pat.shape[1]
        rows = np.repeat(np.arange(pat.shape[0], dtype=np.int64),
                         np.diff(pat.indptr))
        keys = rows * n + pat.indices

        def project(m):
            m.sort_indices()
            r = np.repeat(np.arange(m.shape[0], dtype=np.int64),
                          np.diff(m.indptr))
            vals = np.zeros(len(keys))
            vals[np.searchsorted(keys, r * n + m.indices)] = m.data
            return torch.from_numpy(vals).to(device)

        self.shape = pat.shape
        self.indptr = torch.from_numpy(pat.indptr.astype(np.int32)).to(device)
        self.indices = torch.from_numpy(pat.indices.astype(np.int32)).to(device)
        self.vals0, self.vals1 = project(m0), project(m1)
        self._row_order = _RowOrder(
            self.indptr, pat.shape[0],
            locality_order(pat) if pat.shape[0] >= ROW_ORDER_MIN_ROWS
            and row_order_enabled() else None, device)

    def split(self, x, y0, y1):
        """y0 = mat0 x, y1 = mat1 x."""
        check(lib().stk_space_spmm_split(self.shape[0], ptr(self.indptr),
                                         ptr(self.indices), ptr(self.vals0),
                                         ptr(self.vals1), ptr(x), ptr(y0),
                                         ptr(y1), x.shape[1], stream()))

    def pair(self, x0, x1, out, alpha=1.0, beta=0.0, z=None, ldx=None):
        """out = alpha (mat0 x0 + mat1 x1) + beta z; x0 / x1 may be device
        addresses of sub-blocks with pitch ldx."""
        ld = out.shape[1]
        check(lib().stk_space_spmm_pair(self.shape[0], ptr(self.indptr),
                                        ptr(self.indices), ptr(self.vals0),
                                        ptr(self.vals1), _addr(x0), _addr(x1),
                                        ld if ldx is None else ldx,
                                        float(alpha), float(beta), ptr(z),
                                        ptr(out), ld, stream()))


def _addr(t):
    return t if isinstance(t, int) else ptr(t)


_csr_cache = {}


def as_space_op(op):
    """Device form of whatever the reference accepts as `mat_space`: a scipy
    sparse matrix (uploaded once, cached), or an object that already follows
    the protocol (DeviceCSR, MultiGrid, CompositeLinOp)."""
    if hasattr(op, 'apply_block'):
        return op
    if sp.issparse(op) or isinstance(op, np.ndarray):
        key = id(op)
        if key not in _csr_cache:
            _csr_cache[key] = (op, DeviceCSR(op))
        return _csr_cache[key][1]
    raise TypeError(
        'space operator %r has no device form: pass a scipy sparse matrix, '
        'shared_sparse_matrix(...), MultiGrid or CompositeLinOp (the B200 path '
        'has no CPU fallback for generic LinearOperators)' % type(op))


class CompositeLinOp:
    """x -> linops[0] (linops[1] (... linops[-1] x)) (linop.py:68-79)."""
    def __init__(self, linops):
        self.linops = [as_space_op(l) for l in linops]
        for a, b in zip(self.linops[:-1], self.linops[1:]):
            assert a.shape[1] == b.shape[0]
        self.shape = (self.linops[0].shape[0], self.linops[-1].shape[1])

    def apply_block(self, x, out, ctx=None):
        ops = self.linops[::-1]
        cur = x
        for k, op in enumerate(ops):
            last = k == len(ops) - 1
            if last:
                dst = out
            else:
                rows = op.shape[0]
                dst = torch.empty((rows, x.shape[1]), dtype=x.dtype,
                                  device=x.device)
            op.apply_block(cur, dst)
            cur = dst

    # scipy-style host interface (tests): (M,) or (M, k) arrays
    def __matmul__(self, B):
        return host_apply(self, B)


def lu_factors(mat):
    """Host factorisation behind `InvLinOp`: SuperLU with exactly the options
    of linop.py:20-24, returned as sparse factors
        Pr mat Pc = L U   (L unit lower, U upper triangular, CSR)
    so that  mat^{-1} b = Pc U^{-1} L^{-1} Pr b."""
    import scipy.sparse.linalg as spla
    mat = sp.csc_matrix(mat, dtype=np.float64)
    n = mat.shape[0]
    lu = spla.splu(mat, options={'SymmetricMode': True},
                   permc_spec='MMD_AT_PLUS_A')
    one, idx = np.ones(n), np.arange(n)
    Pr = sp.csr_matrix((one, (lu.perm_r, idx)), shape=(n, n))
    Pc = sp.csr_matrix((one, (idx, lu.perm_c)), shape=(n, n))
    return Pr, sp.csr_matrix(lu.L), sp.csr_matrix(lu.U), Pc


class InvLinOp:
    """Direct inverse of a sparse matrix as a space operator (linop.py:18-26,
    the `precond='direct'` branch of heateq_mpi.py:154-157 and the exact-inverse
    cross-check of heateq_mpi_test.py:66-135).

    The factorisation is SuperLU's, once, on the host (as the reference); the
    SOLVES run on the device for all time slices of a block at once.  A
    triangular solve is the lexicographic Gauss-Seidel sweep of its factor --
    forward for L, backward for U, exact after one sweep -- so it runs on the
    smoother's wavefront kernels (`k_gs_phase`, one launch per level of the
    factor's dependency DAG); the row/column permutations are one-entry-per-row
    SpMMs."""
    def __init__(self, mat, device=None):
        from .multigrid import gauss_seidel_schedule
        if device is None:
            from .mpi_vector import _device
            device = _device()
        if isinstance(mat, DeviceCSR):
            mat = mat.host
        Pr, L, U, Pc = lu_factors(mat)
        self.shape = Pr.shape
        self.dtype = np.dtype(np.float64)
        self.device = device
        self.Pr, self.Pc = DeviceCSR(Pr, device), DeviceCSR(Pc, device)
        self.num_applies = 0
        self._keep = []
        self._tri = []
        for T in (L, U):
            T.sort_indices()
            order, phase_ptr = gauss_seidel_schedule(T.indptr, T.indices)
            dev = [torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)
                   for a, dt in ((T.indptr, np.int32), (T.indices, np.int32),
                                 (T.data, np.float64),
                                 (T.diagonal(), np.float64), (order, np.int32))]
            import ctypes
            h = ctypes.c_void_p(lib().stk_mg_create(2, 1, 1, 1))
            assert h.value, 'stk_mg_create failed'
            ip, ix, dv, dd, od = dev
            check(lib().stk_mg_set_level(h, 1, T.shape[0], int(T.nnz), ptr(ip),
                                         ptr(ix), ptr(dv), ptr(dd), ptr(od),
                                         phase_ptr.ctypes.data,
                                         len(phase_ptr) - 1))
            self._keep.append(dev)
            self._tri.append(h)
        self.depth = None

    def __del__(self):
        try:
            for h in self._tri:
                lib().stk_mg_destroy(h)
            self._tri = []
        except Exception:
            pass

    def apply_block(self, x, out, ctx=None):
        ld = x.shape[1]
        a = torch.empty_like(x)
        b = torch.zeros_like(x)
        self.Pr.spmm(x, a)
        # L b = a (forward sweep from b = 0), U a = b (backward sweep from a = 0)
        check(lib().stk_mg_smooth(self._tri[0], 1, 1, 0, None, ptr(a), ptr(b),
                                  ld, stream()))
        a.zero_()
        check(lib().stk_mg_smooth(self._tri[1], 1, 1, 1, None, ptr(b), ptr(a),
                                  ld, stream()))
        self.Pc.spmm(a, out)
        self.num_applies += 1

    def __matmul__(self, B):
        return host_apply(self, B)

    matvec = matmat = dot = __matmul__


class KronLinOp:
    """Serial (mat_time (x) mat_space) x on a host vector of N*M entries
    (linop.py:6-15), computed on the device.  `mat_time` is any square sparse
    time matrix or an operator with `.as_matrix()` (the wavelet transform); the
    reference's rectangular factors have no use on this path and are refused."""
    def __init__(self, mat_time, mat_space):
        if hasattr(mat_time, 'as_matrix'):
            mat_time = mat_time.as_matrix()
        self.mat_time = sp.csr_matrix(mat_time, dtype=np.float64)
        self.mat_space = as_space_op(mat_space)
        if (self.mat_time.shape[0] != self.mat_time.shape[1]
                or self.mat_space.shape[0] != self.mat_space.shape[1]):
            raise ValueError('KronLinOp: square factors only on the device path')
        self.N, self.M = self.mat_time.shape[0], self.mat_space.shape[0]
        self.shape = (self.N * self.M, self.N * self.M)

    def __matmul__(self, x):
        from .comm import SerialComm
        from .mpi_kron import IdentityKronMatMPI, SparseKronIdentityMPI
        from .mpi_vector import DofDistributionMPI, KronVectorMPI
        d = DofDistributionMPI(SerialComm(), self.N, self.M)
        v = KronVectorMPI(d, np.asarray(x, dtype=np.float64).reshape(
            self.N, self.M))
        out = SparseKronIdentityMPI(d, self.mat_time) @ v
        IdentityKronMatMPI(d, self.mat_space)._matvec(out, out)  # alias-safe
        return out.to_host().reshape(-1)


def host_apply(op, B):
    """`op @ B` for host arrays (M,) or (M, k): columns are the batch, which
    is exactly the time-fastest block layout."""
    from .mpi_vector import _device, pitch
    B = np.asarray(B, dtype=np.float64)
    one = B.ndim == 1
    B2 = B.reshape(B.shape[0], -1)
    k = B2.shape[1]
    ld = pitch(k)
    dev = _device()
    x = torch.zeros((B2.shape[0], ld), dtype=torch.float64, device=dev)
    x[:, :k] = torch.from_numpy(np.ascontiguousarray(B2)).to(dev)
    out = torch.empty((op.shape[0], ld), dtype=torch.float64, device=dev)
    op.apply_block(x, out)
    res = out[:, :k].cpu().numpy()
    return res[:, 0].copy() if one else res
