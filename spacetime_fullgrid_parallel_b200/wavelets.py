"""Time wavelet transform (three-point wavelets -> hat functions).

Drop-in for /root/reference/source/wavelets.py: `WaveletTransformOp`
(:45-169, `.T`, `.levels`, `.split(j)`, interleaved or level-wise ordering),
`WaveletTransformKronIdentityMPI` (:172-183) and its transposed twin
(:186-198).

Device strategy: on one rank the whole time axis of a space dof is contiguous
in HBM, so W and W^T are single-pass in-place lifting kernels
(stk_wavelet_lift).  On P > 1 ranks the J levels are multiplied into one sparse
time matrix whose rows need at most 2J-1 remote slices, so one boundary
exchange replaces the reference's J rounds (timeop.py); W^T is its adjoint.
"""
import numpy as np
import scipy.sparse as sp
import torch

from ._lib import check, lib, ptr, stream
from .mpi_kron import LinearOperatorMPI
from .timeop import LevelChain, TimeOpPlan


def wavelet_levels(J, interleaved=True):
    """Level of the basis function attached to each index
    (wavelets.py:70-79)."""
    if interleaved:
        lv = np.zeros(2**J + 1, dtype=int)
        for j in range(1, J + 1):
            lv[2**(J - j)::2**(J - j + 1)] = j
        return lv
    return np.array([0, 0] + [j for j in range(1, J + 1)
                              for _ in range(2**(j - 1))], dtype=int)


def levelwise_positions(J):
    """Node (= interleaved index) of the k-th function in level-wise order."""
    pos = [0, 2**J]
    for j in range(1, J + 1):
        S = 2**(J - j)
        pos.extend((2 * q + 1) * S for q in range(2**(j - 1)))
    return np.array(pos, dtype=np.int64)


def _level_step(J, j):
    """Sparse matrix of level j of the interleaved transform: acts on the
    nodes m * 2^(J-j), identity elsewhere (wavelets.py:81-104,136-169)."""
    N = 2**J + 1
    S, nj = 2**(J - j), 2**j
    s = 2.0**(j / 2)
    rows, cols, vals = [], [], []
    touched = np.zeros(N, dtype=bool)
    for m in range(nj + 1):
        r = m * S
        touched[r] = True
        if m % 2:  # wavelet node: s on itself, 1/2 from both coarse hats
            rows += [r, r, r]
            cols += [r, r - S, r + S]
            vals += [s, 0.5, 0.5]
        else:  # coarse hat: itself minus s/2 of the adjacent wavelets
            rows.append(r)
            cols.append(r)
            vals.append(1.0)
            left = r - S if m > 0 else r + S
            right = r + S if m < nj else r - S
            rows += [r, r]
            cols += [left, right]
            vals += [-0.5 * s, -0.5 * s]
    idle = np.nonzero(~touched)[0]
    rows += list(idle)
    cols += list(idle)
    vals += [1.0] * len(idle)
    step = sp.coo_matrix((vals, (rows, cols)), shape=(N, N)).tocsr()
    step.sum_duplicates()
    return step


def wavelet_dependency_pattern(J):
    """Structural (0/1) pattern of the product of the J level steps: entry
    (t, s) is set when input s can reach output t through the lifting steps.
    It contains the pattern of W; the two differ where contributions cancel
    exactly in the product (W[1, 2] = 0 although node 2 is an intermediate of
    row 1), and a chain of level steps needs those intermediates too."""
    S = sp.identity(2**J + 1, format='csr')
    for j in range(1, J + 1):
        F = _level_step(J, j).copy()
        F.data = np.ones_like(F.data)
        S = (F @ S).tocsr()
        S.data = np.ones_like(S.data)
    S.sort_indices()
    return S


def WaveletTransformMat(J):
    """The level-wise transform as an explicit sparse matrix (the debugging
    aid of wavelets.py:9-42)."""
    return WaveletTransformOp(J, interleaved=False).as_matrix()


class WaveletTransformOp:
    """W: wavelet coordinates -> hat coordinates along axis 0 of an (N, k)
    array, N = 2^J + 1 (wavelets.py:45-169)."""
    def __init__(self, J, interleaved=False, _transposed=False):
        self.J = J
        self.interleaved = interleaved
        self.transposed = _transposed
        self.N = 2**J + 1
        self.shape = (self.N, self.N)
        self.dtype = np.dtype(np.float64)
        self.levels = wavelet_levels(J, interleaved)
        self._mat = None

    @property
    def T(self):
        return WaveletTransformOp(self.J, self.interleaved,
                                  not self.transposed)

    def split(self, j):
        """Level-j step minus the identity: W = prod_j (I + split(j)), j = 1
        applied first (wavelets.py:136-169).  Interleaved ordering only."""
        assert self.interleaved
        out = (_level_step(self.J, j) - sp.identity(self.N)).tocsr()
        out.eliminate_zeros()
        return out

    def as_matrix(self):
        """The transform as a sparse matrix (setup-time host product)."""
        if self._mat is None:
            W = sp.identity(self.N, format='csr')
            for j in range(1, self.J + 1):
                W = (_level_step(self.J, j) @ W).tocsr()
            if not self.interleaved:  # columns in level-wise order
                W = W[:, levelwise_positions(self.J)].tocsr()
            if self.transposed:
                W = W.T.tocsr()
            W.sort_indices()
            self._mat = W
        return self._mat

    def apply_block(self, x, out=None):
        """In-place on a device block (k, ld) holding the whole time axis of
        k columns; interleaved ordering."""
        assert self.interleaved
        dst = x if out is None else out
        check(lib().stk_wavelet_lift(x.shape[0], self.J, int(self.transposed),
                                     ptr(x), ptr(dst), x.shape[1], stream()))
        return dst

    def __matmul__(self, X):
        """Host arrays (N,) or (N, k), computed on the device."""
        from .mpi_vector import _device, pitch
        X = np.asarray(X, dtype=np.float64)
        one = X.ndim == 1
        X2 = X.reshape(self.N, -1)
        pos = levelwise_positions(self.J)
        if not self.interleaved and not self.transposed:
            Z = np.empty_like(X2)
            Z[pos] = X2  # level-wise coefficients to their nodes
            X2 = Z
        k = X2.shape[1]
        ld = pitch(self.N)
        blk = torch.zeros((k, ld), dtype=torch.float64, device=_device())
        blk[:, :self.N] = torch.from_numpy(np.ascontiguousarray(X2.T)).to(
            blk.device)
        check(lib().stk_wavelet_lift(k, self.J, int(self.transposed), ptr(blk),
                                     ptr(blk), ld, stream()))
        Y = blk[:, :self.N].cpu().numpy().T
        if not self.interleaved and self.transposed:
            Y = Y[pos]
        Y = np.ascontiguousarray(Y)
        return Y[:, 0].copy() if one else Y

    matvec = matmat = dot = __matmul__


class WaveletTransformKronIdentityMPI(LinearOperatorMPI):
    """W (x) I on a sharded vector (wavelets.py:172-183)."""
    transposed = False

    def __init__(self, dofs_distr, J):
        super().__init__(dofs_distr)
        assert dofs_distr.N == 2**J + 1
        self.J = J
        self.op = WaveletTransformOp(J, interleaved=True)
        self.levels = self.op.levels
        self.plan = self.chain = None
        if dofs_distr.size > 1:
            # exchange lists from the dependency pattern of the product of all
            # levels (the closure of the local rows under the lifting steps);
            # arithmetic as the chain of those steps on [local | halo] columns
            import os
            if os.environ.get('STK_WAVELET_CHAIN', '1') == '0':
                # the fused-matrix form (one sparse time matrix with ~2J+1
                # entries per row), kept as a cross-check of the chain
                self.plan = TimeOpPlan(dofs_distr, self.op.as_matrix())
            else:
                self.plan = TimeOpPlan(dofs_distr,
                                       wavelet_dependency_pattern(J))
                steps = [_level_step(J, j) for j in range(1, J + 1)]
                if self.transposed:
                    steps = [G.T.tocsr() for G in reversed(steps)]
                self.chain = LevelChain(self.plan, steps)

    def _matvec(self, vec_in, vec_out):
        if self.plan is None:
            vec_out._invalidate()
            check(lib().stk_wavelet_lift(vec_out.M, self.J,
                                         int(self.transposed),
                                         ptr(vec_in.data), ptr(vec_out.data),
                                         vec_out.ld, stream()))
            return vec_out
        assert vec_out is not vec_in
        t0 = self.plan.__dict__.get('time_communication', 0.0)
        vec_out._invalidate()
        pl = self.plan
        if self.chain is None:
            if self.transposed:
                pl.apply_adjoint(vec_in, vec_out)
            else:
                pl.apply(vec_in, vec_out.data)
        elif self.transposed:
            # local rows and the partial sums for remote slices in one pass,
            # then the adjoint of the halo exchange
            packed = torch.empty((pl.n_halo, vec_in.M), dtype=torch.float64,
                                 device=vec_in.data.device) if pl.n_halo else None
            self.chain.apply(vec_in.data, vec_in.ld, vec_in.M, vec_out.data,
                             vec_out.ld, None, packed)
            pl.scatter_add_halo(vec_out, packed)
        else:
            halo = pl.fetch(vec_in) if pl.n_halo else None
            self.chain.apply(vec_in.data, vec_in.ld, vec_in.M, vec_out.data,
                             vec_out.ld, halo, None)
        self.time_communication += self.plan.__dict__.get(
            'time_communication', 0.0) - t0
        return vec_out


class TransposedWaveletTransformKronIdentityMPI(WaveletTransformKronIdentityMPI
                                                ):
    """W^T (x) I (wavelets.py:186-198)."""
    transposed = True
