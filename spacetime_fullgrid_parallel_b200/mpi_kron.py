"""Space-time operators on time-sharded device vectors.

Drop-in for /root/reference/source/mpi_kron.py: every class keeps its name,
constructor and `_matvec(vec_in, vec_out)` / `@` contract, the counters
`num_applies / time_applies / time_communication` and `as_global_matrix`.
Underneath, a `_matvec` is a handful of libstk launches on the whole local
block instead of NumPy/SciPy calls slice by slice.
"""
import numpy as np
import scipy.sparse as sp
import torch

from ._lib import check, lib, ptr, stream
from . import comm as _comm
from .comm import Wtime_device as Wtime
from .linop import CompositeLinOp, as_space_op
from .mpi_vector import KronVectorMPI, pitch
from .timeop import TimeOpPlan


def as_matrix(operator):
    """Dense matrix of a host-interface operator (mpi_kron.py:8-10)."""
    return operator @ np.eye(operator.shape[1])


class LinearOperatorMPI:
    """Base class (mpi_kron.py:13-59).

    Kernels are enqueued asynchronously: `time_applies` and
    `time_communication` have the reference's meaning (wall time until the
    result exists) only while `comm.SYNC_TIMING` is on, which the drivers that
    report them (heateq_mpi.main, heateq_mpi_timing) do; otherwise they count
    host enqueue time."""

    def __init__(self, dofs_distr):
        self.dofs_distr = dofs_distr
        self.N = dofs_distr.N
        self.M = dofs_distr.M
        self.num_applies = 0
        self.time_applies = 0
        self.time_communication = 0

    def __matmul__(self, x):
        assert isinstance(x, KronVectorMPI)
        start = Wtime()
        y = self._matvec(x, x.empty_like())
        self.num_applies += 1
        self.time_applies += Wtime() - start
        return y

    def time_per_apply(self):
        assert self.time_applies
        return (self.time_applies / self.num_applies,
                self.time_communication / self.num_applies)

    def as_global_matrix(self):
        """Dense (NM x NM) matrix on rank 0, column by column (tests only;
        mpi_kron.py:38-59)."""
        n = self.N * self.M
        rank = self.dofs_distr.rank
        result = np.zeros((n, n)) if rank == 0 else None
        x_glob = np.empty(n) if rank == 0 else None
        for k in range(n):
            e = None
            if rank == 0:
                e = np.zeros(n)
                e[k] = 1.0
            x = KronVectorMPI(self.dofs_distr)
            x.scatter(e)
            y = self @ x
            y.gather(x_glob)
            if rank == 0:
                result[:, k] = x_glob
        return result


class IdentityMPI(LinearOperatorMPI):
    def _matvec(self, vec_in, vec_out):
        vec_out._invalidate()
        vec_out.data.copy_(vec_in.data)
        return vec_out


class SumMPI(LinearOperatorMPI):
    """sum_k linops[k] (mpi_kron.py:71-90)."""
    def __init__(self, dofs_distr, linops):
        assert all(isinstance(l, LinearOperatorMPI) for l in linops)
        self.linops = linops
        super().__init__(dofs_distr)

    def _matvec(self, vec_in, vec_out):
        assert vec_in is not vec_out
        self.time_communication = 0
        tmp = None
        for k, linop in enumerate(self.linops):
            c0 = linop.time_communication
            if k == 0:
                linop._matvec(vec_in, vec_out)
            else:
                tmp = vec_in.empty_like() if tmp is None else tmp
                linop._matvec(vec_in, tmp)
                vec_out += tmp
            self.time_communication += linop.time_communication - c0
        return vec_out


class CompositeMPI(LinearOperatorMPI):
    """linops[0] o linops[1] o ... (mpi_kron.py:93-110)."""
    def __init__(self, dofs_distr, linops):
        assert all(isinstance(l, LinearOperatorMPI) for l in linops)
        N, M = linops[0].N, linops[0].M
        assert all(l.N == N and l.M == M for l in linops)
        self.linops = linops
        super().__init__(dofs_distr)

    def _matvec(self, vec_in, vec_out):
        assert vec_in is not vec_out
        self.time_communication = 0
        Y = vec_in
        ops = list(reversed(self.linops))
        for k, linop in enumerate(ops):
            c0 = linop.time_communication
            start = Wtime()
            dst = vec_out if k == len(ops) - 1 else Y.empty_like()
            Y = linop._matvec(Y, dst)
            linop.num_applies += 1
            linop.time_applies += Wtime() - start
            self.time_communication += linop.time_communication - c0
        return vec_out


class IdentityKronMatMPI(LinearOperatorMPI):
    """I (x) mat_space: the space operator on every local slice at once
    (mpi_kron.py:135-150).  Alias-safe (vec_out may be vec_in)."""
    def __init__(self, dofs_distr, mat_space):
        assert mat_space.shape == (dofs_distr.M, dofs_distr.M)
        super().__init__(dofs_distr)
        self.mat_space = as_space_op(mat_space)

    def _matvec(self, vec_in, vec_out):
        src = vec_in.data
        if vec_out is vec_in:
            dst = torch.empty_like(src)
            self.mat_space.apply_block(src, dst)
            vec_out._invalidate()
            vec_out.data = dst
        else:
            vec_out._invalidate()
            self.mat_space.apply_block(src, vec_out.data)
        return vec_out


class _TimeOpMPI(LinearOperatorMPI):
    """(T (x) I) for a sparse time matrix T with halo exchange."""
    def __init__(self, dofs_distr, mat_time):
        super().__init__(dofs_distr)
        self.mat_time = sp.csr_matrix(mat_time, dtype=np.float64)
        assert self.mat_time.shape == (self.N, self.N)
        self.plan = TimeOpPlan(dofs_distr, self.mat_time)

    def _apply_time(self, vec_in, vec_out):
        assert vec_in is not vec_out
        vec_out._invalidate()
        c0 = getattr(self.plan, 'time_communication', 0.0)
        self.plan.apply(vec_in, vec_out.data)
        self.time_communication += getattr(self.plan, 'time_communication',
                                           0.0) - c0
        return vec_out


class TridiagKronIdentityMPI(_TimeOpMPI):
    """T (x) I for tridiagonal T: +-1-slice halo (mpi_kron.py:153-201)."""
    def __init__(self, dofs_distr, mat_time):
        T = sp.csr_matrix(mat_time)
        coo = T.tocoo()
        assert np.all(np.abs(coo.row - coo.col) <= 1), 'not tridiagonal'
        super().__init__(dofs_distr, T)

    def _matvec(self, vec_in, vec_out):
        return self._apply_time(vec_in, vec_out)


class TridiagKronMatMPI(LinearOperatorMPI):
    """mat_time (x) mat_space (mpi_kron.py:204-222): time stencil first, then
    the space operator in place."""
    def __init__(self, dofs_distr, mat_time, mat_space):
        super().__init__(dofs_distr)
        self.N, self.M = mat_time.shape[0], mat_space.shape[0]
        self.T_I = TridiagKronIdentityMPI(dofs_distr, mat_time)
        self.I_M = IdentityKronMatMPI(dofs_distr, mat_space)
        self.mat_time, self.mat_space = mat_time, mat_space

    def _matvec(self, vec_in, vec_out):
        c0 = self.T_I.time_communication
        self.T_I._matvec(vec_in, vec_out)
        self.I_M._matvec(vec_out, vec_out)
        self.time_communication += self.T_I.time_communication - c0
        return vec_out

    def as_matrix(self):
        return np.kron(self.mat_time.toarray(),
                       self.mat_space.toarray()
                       if hasattr(self.mat_space, 'toarray') else
                       as_matrix(self.mat_space))


class SparseKronIdentityMPI(_TimeOpMPI):
    """(mat_time [+ I]) (x) I for a sparse time matrix with symmetric pattern
    (mpi_kron.py:259-317); the needed remote slices come in one exchange."""
    def __init__(self, dofs_distr, mat_time, add_identity=False):
        T = sp.csr_matrix(mat_time, dtype=np.float64)
        if add_identity:
            T = (T + sp.identity(T.shape[0], format='csr')).tocsr()
        self.add_identity = add_identity
        super().__init__(dofs_distr, T)

    def _matvec(self, vec_in, vec_out):
        return self._apply_time(vec_in, vec_out)


class MatKronIdentityMPI(LinearOperatorMPI):
    """mat_time (x) I for an arbitrary time operator via the time<->space
    all-to-all (mpi_kron.py:225-256): re-shard by space, apply mat_time along
    the now-local, contiguous time axis, re-shard back."""
    def __init__(self, dofs_distr, mat_time):
        super().__init__(dofs_distr)
        from .permute import plan_for
        if hasattr(mat_time, 'levels'):
            self.levels = mat_time.levels
        self.mat_time = mat_time
        T = mat_time.as_matrix() if hasattr(mat_time, 'as_matrix') else mat_time
        T = sp.csr_matrix(T, dtype=np.float64)
        T.sort_indices()
        assert T.shape == (self.N, self.N)
        self.T = T
        self.pplan = plan_for(dofs_distr)
        self._dev = None
        self._pos = None

    def _apply_time(self, sblock):
        """mat_time along the (complete, contiguous) time axis of a
        space-sharded block.  A wavelet transform runs as in-place lifting
        (level-wise ordering: plus the column permutation to/from node order);
        any other matrix goes through the sparse time-operator kernel."""
        from .wavelets import WaveletTransformOp, levelwise_positions
        op = self.mat_time
        if isinstance(op, WaveletTransformOp):
            N = self.N
            if op.interleaved:  # out of place, one pass
                res = torch.empty_like(sblock)
                check(lib().stk_wavelet_lift(res.shape[0], op.J,
                                             int(op.transposed), ptr(sblock),
                                             ptr(res), res.shape[1],
                                             stream()))
                return res
            else:
                # level-wise ordering (wavelets.py:113-117): coefficient k
                # lives at node pos[k]; gather along the time axis with the
                # permutation or its inverse, lifting in between
                if self._pos is None:
                    pos = levelwise_positions(op.J)
                    inv = np.empty_like(pos)
                    inv[pos] = np.arange(len(pos))
                    self._pos = (torch.from_numpy(pos.astype(np.int32)).to(sblock.device),
                                 torch.from_numpy(inv.astype(np.int32)).to(sblock.device))
                pos, inv = self._pos
                ldb = sblock.shape[1]
                res = torch.zeros_like(sblock)
                if not op.transposed:  # coefficients to their nodes, then lift
                    check(lib().stk_copy_cols(sblock.shape[0], N, ptr(sblock), ldb, 0,
                                              ptr(inv), ptr(res), ldb, 0, stream()))
                    check(lib().stk_wavelet_lift(res.shape[0], op.J, 0, ptr(res),
                                                 ptr(res), ldb, stream()))
                    return res
                tmp = torch.empty_like(sblock)  # lift, then nodes to coefficients
                check(lib().stk_wavelet_lift(res.shape[0], op.J, 1, ptr(sblock),
                                             ptr(tmp), ldb, stream()))
                check(lib().stk_copy_cols(sblock.shape[0], N, ptr(tmp), ldb, 0, ptr(pos),
                                          ptr(res), ldb, 0, stream()))
                return res
        if self._dev is None:
            dev = sblock.device
            self._dev = tuple(
                torch.from_numpy(a).to(dev)
                for a in (self.T.indptr.astype(np.int32),
                          self.T.indices.astype(np.int32),
                          self.T.data.astype(np.float64)))
        indptr, indices, vals = self._dev
        res = torch.empty_like(sblock)
        check(lib().stk_time_apply(sblock.shape[0], self.N, self.T.nnz,
                                   ptr(indptr), ptr(indices), ptr(vals),
                                   ptr(sblock), sblock.shape[1], self.N, None,
                                   0, 1.0, 0.0, ptr(res), res.shape[1],
                                   stream()))
        return res

    def _matvec(self, vec_in, vec_out):
        vec_out._invalidate()
        if self.dofs_distr.size == 1:
            # one rank: the block already holds the whole time axis per dof
            vec_out.data = self._apply_time(vec_in.data)
            return vec_out
        start = Wtime()
        sblock = self.pplan.forward(vec_in.data, vec_in.n_loc, vec_in.ld)
        self.time_communication += Wtime() - start
        res = self._apply_time(sblock)
        start = Wtime()
        vec_out.data = self.pplan.backward(res, vec_in.n_loc, vec_in.ld)
        self.time_communication += Wtime() - start
        return vec_out


class BlockDiagMPI(LinearOperatorMPI):
    """Block diagonal in time: slice t gets matrices_space[t]
    (mpi_kron.py:113-132).

    The reference loops over slices.  Here the local slices are batched:
      * operators that differ only in the coefficients of one
        `MultiGridFamily` (the C_j A_x C_j of heateq_mpi.py:159-162,183-184)
        run as ONE block apply with per-slice coefficients;
      * anything else is grouped by operator object, each group gathered into
        a compact block, applied, and scattered back.
    """
    def __init__(self, dofs_distr, matrices_space):
        M = matrices_space[0].shape[0]
        for mat in matrices_space:
            assert mat.shape == (M, M)
        assert len(matrices_space) == dofs_distr.N
        self.matrices_space = matrices_space
        super().__init__(dofs_distr)
        a, b = dofs_distr.t_begin, dofs_distr.t_end
        self._local = [as_space_op(m) for m in matrices_space[a:b]]
        self._chain = self._batched_chain(self._local, pitch(b - a))
        self._groups = None

    @staticmethod
    def _batched_chain(ops, ld):
        """[(operator, ctx)] applied right to left, or None."""
        from .multigrid import MultiGrid
        first = ops[0]
        if all(o is first for o in ops):
            return [(first, None)]
        chains = [o.linops if isinstance(o, CompositeLinOp) else [o]
                  for o in ops]
        if any(len(c) != len(chains[0]) for c in chains):
            return None
        out = []
        for pos in range(len(chains[0])):
            col = [c[pos] for c in chains]
            if all(o is col[0] for o in col):
                out.append((col[0], None))
            elif all(isinstance(o, MultiGrid) and o.family is col[0].family
                     for o in col):
                fam = col[0].family
                out.append((fam, fam.context([o.coefs for o in col], ld)))
            else:
                return None
        return out

    def _matvec(self, vec_in, vec_out):
        assert vec_out is not vec_in
        assert self.N == vec_in.N and self.M == vec_in.M
        vec_out._invalidate()
        if self._chain is not None:
            cur = vec_in.data
            stages = self._chain[::-1]
            for k, (op, ctx) in enumerate(stages):
                dst = vec_out.data if k == len(stages) - 1 else torch.empty_like(
                    vec_in.data)
                if ctx is None:
                    op.apply_block(cur, dst)
                else:
                    op.apply_block(cur, dst, ctx)
                cur = dst
            return vec_out
        return self._matvec_grouped(vec_in, vec_out)

    def _matvec_grouped(self, vec_in, vec_out):
        dev = vec_in.data.device
        if self._groups is None:
            groups = {}
            for t, op in enumerate(self._local):
                groups.setdefault(id(op), (op, []))[1].append(t)
            self._groups = [(op, torch.tensor(ts, dtype=torch.int32,
                                              device=dev))
                            for op, ts in groups.values()]
        M = vec_in.M
        for op, tidx in self._groups:
            n = len(tidx)
            ldg = pitch(n)
            packed = torch.empty((n, M), dtype=torch.float64, device=dev)
            check(lib().stk_pack_slices(ptr(vec_in.data), vec_in.ld, M,
                                        ptr(tidx), n, ptr(packed), stream()))
            blk = torch.empty((M, ldg), dtype=torch.float64, device=dev)
            check(lib().stk_block_from_rowmajor(ptr(packed), n, M, ptr(blk),
                                                ldg, stream()))
            res = torch.empty_like(blk)
            op.apply_block(blk, res)
            check(lib().stk_block_to_rowmajor(ptr(res), ldg, n, M, ptr(packed),
                                              stream()))
            check(lib().stk_unpack_slices(ptr(vec_out.data), vec_out.ld, M,
                                          ptr(tidx), n, ptr(packed), 1.0, 0.0,
                                          stream()))
        # pads of vec_out are not written by unpack
        if vec_out.ld > vec_out.n_loc:
            vec_out.data[:, vec_out.n_loc:].zero_()
        return vec_out
