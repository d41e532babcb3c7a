"""CPU: host-side logic of the product (partition, exchange plans, wavelet
matrices, Gauss-Seidel schedule, host assembler) and the C ABI surface.
No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import scipy.sparse as sp

from conftest import ROOT, rand, rel
from oracle import restate as orc
from spacetime_fullgrid_parallel_b200 import _lib
from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
from spacetime_fullgrid_parallel_b200.comm import SerialComm
from spacetime_fullgrid_parallel_b200.mpi_vector import (DofDistributionMPI,
                                                         pitch)
from spacetime_fullgrid_parallel_b200.timeop import LevelChain, TimeOpPlan
from spacetime_fullgrid_parallel_b200.wavelets import (WaveletTransformOp,
                                                       _level_step,
                                                       wavelet_dependency_pattern,
                                                       levelwise_positions,
                                                       wavelet_levels)


class FakeComm(SerialComm):
    def __init__(self, rank, size):
        self.rank, self.size = rank, size

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size


def test_abi_exports_every_declared_symbol():
    """libstk.so loads and exports exactly what include/stk.h declares."""
    hdr = open(os.path.join(ROOT, 'include', 'stk.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(stk_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 25
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.stk_version() == 100
    # the ctypes signatures carry as many arguments as the C declarations
    for name, params in re.findall(r'\b(stk_[a-z0-9_]+)\s*\(([^)]*)\)', hdr):
        params = params.strip()
        n = 0 if params in ('', 'void') else params.count(',') + 1
        assert n == len(_lib.SIGNATURES[name][1]), (name, n)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB', '/nonexistent/libstk.so')
    try:
        _lib.lib()
    except _lib.StkError as e:
        assert 'no CPU fallback' in str(e)
    else:
        raise AssertionError('expected StkError')


def test_dof_distribution():
    """mpi_vector.py:18-38."""
    for N, P in ((9, 2), (257, 8), (1025, 8), (5, 5), (13, 3)):
        bounds = orc.slab_bounds(N, P)
        for r in range(P):
            d = DofDistributionMPI(FakeComm(r, P), N, 7)
            assert (d.t_begin, d.t_end) == bounds[r]
            assert [tuple(b) for b in d.dof_distribution] == bounds
            assert d.counts.sum() == N * 7 and d.displs[r] == bounds[r][0] * 7
            assert all(d.dof2proc[a:b].tolist() == [p] * (b - a)
                       for p, (a, b) in enumerate(bounds))
    assert pitch(1) == 4 and pitch(4) == 4 and pitch(33) == 36 and pitch(257) == 260


def _emulate(T, P, M=3, adjoint=False):
    """Run the plans of all P ranks with numpy standing in for the device."""
    N = T.shape[0]
    X = rand((N, M), seed=N + P)
    plans = [TimeOpPlan(DofDistributionMPI(FakeComm(r, P), N, M), T)
             for r in range(P)]
    bounds = plans[0].dofs_distr.dof_distribution
    out = np.zeros((N, M))
    if not adjoint:
        for r, pl in enumerate(plans):
            a, b = bounds[r]
            halo = np.zeros((pl.n_halo, M))
            for p, (off, cnt) in pl.recv_from.items():
                # what rank p packs for r must be what r expects, in order
                sent = X[bounds[p][0] + plans[p].send_to[r]]
                assert np.array_equal(bounds[p][0] + plans[p].send_to[r],
                                      pl.halo_cols[off:off + cnt])
                halo[off:off + cnt] = sent
            assert set(pl.recv_from) == {p for p in range(P)
                                         if r in plans[p].send_to}
            out[a:b] = pl.local @ np.concatenate([X[a:b], halo])
        return out, T @ X
    for r, pl in enumerate(plans):
        a, b = bounds[r]
        out[a:b] += pl.adj_local @ X[a:b]
        part = pl.adj_halo @ X[a:b]
        for p, (off, cnt) in pl.recv_from.items():
            out[bounds[p][0] + plans[p].send_to[r]] += part[off:off + cnt]
    return out, T.T @ X


def test_timeop_plans():
    N = 17
    tri = sp.diags([rand((N - 1, 1), 1)[:, 0], rand((N, 1), 2)[:, 0],
                    rand((N - 1, 1), 3)[:, 0]], [-1, 0, 1], format='csr')
    W = WaveletTransformOp(4, interleaved=True).as_matrix()
    G = sp.csr_matrix(([1.0], ([0], [0])), shape=(N, N))
    for T in (tri, W, G, sp.identity(N, format='csr')):
        for P in (1, 2, 3, 4, 8, 17):
            got, ref = _emulate(T, P)
            assert rel(got, ref) < 1e-14
            got, ref = _emulate(T, P, adjoint=True)
            assert rel(got, ref) < 1e-14


def test_wavelet_halo_is_small():
    """W needs at most 2J-1 remote slices per rank at the
    BASELINE decompositions (structural closure of the lifting steps)."""
    for J, P in ((8, 8), (10, 8), (6, 4)):
        W = wavelet_dependency_pattern(J)
        for r in range(P):
            pl = TimeOpPlan(DofDistributionMPI(FakeComm(r, P), 2**J + 1, 1), W)
            assert pl.n_halo <= 2 * J - 1


def test_wavelet_matrices(golden):
    g = golden['wavelets']
    for J in range(1, 7):
        X = rand((2**J + 1, 3), seed=J)
        for inter, tag in ((True, 'int'), (False, 'lvl')):
            op = WaveletTransformOp(J, interleaved=inter)
            assert np.array_equal(op.levels, g['levels_J%d_%s' % (J, tag)])
            assert np.array_equal(op.levels, orc.wavelet_levels(J, inter))
            assert rel(op.as_matrix() @ X, g['W_J%d_%s' % (J, tag)]) < 1e-14
            assert rel(op.T.as_matrix() @ X, g['WT_J%d_%s' % (J, tag)]) < 1e-14
    op = WaveletTransformOp(4, interleaved=True)
    for j in range(1, 5):
        assert rel(op.split(j).toarray(), g['split_J4_j%d' % j]) < 1e-15
    assert sorted(levelwise_positions(5)) == list(range(33))
    assert np.array_equal(wavelet_levels(3, True)[levelwise_positions(3)],
                          wavelet_levels(3, False))


def test_gauss_seidel_schedule():
    """Wavefronts respect the lexicographic dependencies; the class ordering
    gives <= 4 wavefronts on every level (assembly.py)."""
    from spacetime_fullgrid_parallel_b200.multigrid import gauss_seidel_schedule
    for order, max_depth in (('class', 4), ('lex', None), ('random', None)):
        prob = SquareProblem(4, 1, order=order, seed=3)
        A = prob.M_x
        rows, phase_ptr = gauss_seidel_schedule(A.indptr, A.indices)
        assert sorted(rows.tolist()) == list(range(A.shape[0]))
        wave = np.empty(A.shape[0], dtype=int)
        for ph in range(len(phase_ptr) - 1):
            wave[rows[phase_ptr[ph]:phase_ptr[ph + 1]]] = ph
        coo = A.tocoo()
        off = coo.row != coo.col
        lo, hi = np.minimum(coo.row, coo.col)[off], np.maximum(coo.row,
                                                               coo.col)[off]
        assert np.all(wave[lo] < wave[hi])
        if max_depth:
            assert len(phase_ptr) - 1 <= max_depth


def test_assembler():
    """Row sums / symmetry / sizes of the host assembler (stands in for
    NGSolve, heateq_mpi.py:63-104)."""
    prob = SquareProblem(3, 2)
    assert prob.M == (2**4 - 1)**2 and prob.N == 5
    for A in (prob.M_x, prob.A_x, prob.A_t, prob.M_t):
        assert abs(A - A.T).max() < 1e-14
    assert abs(prob.M_t.sum() - 1.0) < 1e-14  # int_0^1 1 dt
    assert abs(prob.A_t.sum()) < 1e-12
    assert abs((prob.L_t + prob.L_t.T).toarray() - np.diag(
        [-1] + [0] * (prob.N - 2) + [1])).max() < 1e-14
    ones = np.ones(prob.N)
    assert abs(ones @ (prob.L_t @ ones)) < 1e-14
    assert prob.A_x.getnnz(axis=1).max() == 5  # zeros eliminated
    assert prob.M_x.getnnz(axis=1).max() == 7


def test_level_chain_equals_wavelet_rows():
    """The lifting steps restricted to [local | halo] slices reproduce the
    local rows of W (and, transposed and reversed, of W^T with the partial
    sums for the remote slices) for every rank of every decomposition, down to
    one slice per rank.  The halo set must be the STRUCTURAL closure: W[1, 2]
    is exactly zero although row 1 needs the intermediate value of node 2."""
    W3 = WaveletTransformOp(3, interleaved=True).as_matrix()
    assert W3[1, 2] == 0.0 and wavelet_dependency_pattern(3)[1, 2] == 1.0
    for J in range(1, 8):
        N = 2**J + 1
        W = WaveletTransformOp(J, interleaved=True).as_matrix()
        pattern = wavelet_dependency_pattern(J)
        assert (abs(W) > 0).astype(int).multiply(pattern).sum() == W.nnz
        X = rand((N, 2), seed=J)
        steps = [_level_step(J, j) for j in range(1, J + 1)]
        for P in sorted({1, 2, 3, 5, 8, N // 2, N - 1, N} & set(range(1, N + 1))):
            for r in range(P):
                d = DofDistributionMPI(FakeComm(r, P), N, 1)
                pl = TimeOpPlan(d, pattern)
                assert pl.n_halo <= 2 * J
                a, b = d.t_begin, d.t_end
                E = np.concatenate([np.arange(a, b), pl.halo_cols])
                fw = LevelChain(pl, steps)
                assert np.abs(fw.apply_host(X[E])[:b - a] -
                              (W @ X)[a:b]).max() < 1e-12, (J, P, r)
                ad = LevelChain(pl, [G.T.tocsr() for G in reversed(steps)])
                ext = np.zeros((len(E), 2))
                ext[:b - a] = X[a:b]
                contrib = W[a:b].T @ X[a:b]  # this rank's share of W^T X
                assert np.abs(ad.apply_host(ext) - contrib[E]).max() < 1e-12
                rest = np.ones(N, dtype=bool)
                rest[E] = False
                assert not rest.any() or np.abs(contrib[rest]).max() == 0.0
    # at the BASELINE size the chain does about a third of the multiply-adds
    J, P = 10, 8
    pl = TimeOpPlan(DofDistributionMPI(FakeComm(3, P), 2**J + 1, 1),
                    wavelet_dependency_pattern(J))
    fw = LevelChain(pl, [_level_step(J, j) for j in range(1, J + 1)])
    W = WaveletTransformOp(J, interleaved=True).as_matrix()
    assert 2.5 * len(fw.tcol) < W[pl.dofs_distr.t_begin:pl.dofs_distr.t_end].nnz


def test_wavefront_sweep_equals_sequential_sweep():
    """Emulates what the GPU smoother does -- wavefront by wavefront, all rows
    of a wavefront from the same old values, rows inside a wavefront in the
    locality-sorted order -- and compares with the oracle's sequential
    lexicographic sweeps (multigrid.py:89-97), forward and backward, for the
    three numberings and for the 3-D cube."""
    from oracle import cgs
    from spacetime_fullgrid_parallel_b200.assembly import CubeProblem
    from spacetime_fullgrid_parallel_b200.multigrid import gauss_seidel_schedule
    cases = [SquareProblem(3, 1, order=o, seed=5) for o in ('class', 'lex', 'random')]
    cases.append(CubeProblem(2, 1))
    for prob in cases:
        A = prob.Cinv_j[1].tocsr()
        A.sort_indices()
        n = A.shape[0]
        rows, phase_ptr = gauss_seidel_schedule(A.indptr, A.indices)
        diag = A.diagonal()
        f, u0 = rand((n, ), seed=1), rand((n, ), seed=2)
        for backward in (False, True):
            ref = u0.copy()
            cgs.gauss_seidel(A.indptr.astype(np.int32), A.indices.astype(np.int32),
                             A.data, f, ref, 2, backward=backward)
            u = u0.copy()
            order = range(len(phase_ptr) - 1)
            for _ in range(2):
                for ph in (reversed(order) if backward else order):
                    sel = rows[phase_ptr[ph]:phase_ptr[ph + 1]]
                    u[sel] += (f[sel] - (A[sel] @ u)) / diag[sel]
            assert rel(u, ref) < 1e-13


def test_triangular_solves_as_wavefront_sweeps():
    """`InvLinOp` (linop.py:18-26) on the device = SuperLU's factors applied by
    the smoother's wavefront kernels: one forward sweep on L and one backward
    sweep on U are exact triangular solves.  The factors are NOT structurally
    symmetric, so the wavefronts must come from the symmetrised pattern."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from spacetime_fullgrid_parallel_b200.linop import lu_factors
    from spacetime_fullgrid_parallel_b200.multigrid import gauss_seidel_schedule
    prob = SquareProblem(3, 1)
    A = prob.Cinv_j[1]
    Pr, L, U, Pc = lu_factors(A)
    assert abs(Pr @ A @ Pc - L @ U).max() < 1e-13
    b = rand((A.shape[0], 2), seed=3)
    cur = Pr @ b
    for T, backward in ((L, False), (U, True)):
        T = sp.csr_matrix(T)
        T.sort_indices()
        rows, phase_ptr = gauss_seidel_schedule(T.indptr, T.indices)
        assert len(phase_ptr) - 1 > 1
        diag = T.diagonal()
        u = np.zeros_like(cur)
        order = range(len(phase_ptr) - 1)
        for ph in (reversed(order) if backward else order):
            sel = rows[phase_ptr[ph]:phase_ptr[ph + 1]]
            u[sel] += (cur[sel] - (T[sel] @ u)) / diag[sel][:, None]
        cur = u
    x = Pc @ cur
    ref = spla.splu(sp.csc_matrix(A), options={'SymmetricMode': True},
                    permc_spec='MMD_AT_PLUS_A').solve(b)
    assert rel(x, ref) < 1e-13


def test_chained_wavelet_all_ranks_with_exchange():
    """W and W^T over ALL ranks of a decomposition through the chain, with the
    halo exchange and its adjoint emulated exactly as `fetch` /
    `scatter_add_halo` pair the send and receive lists (what the 8-GPU run
    exercises), down to one slice per rank."""
    for J, P in ((3, 8), (3, 9), (2, 4), (4, 16), (5, 3), (6, 8)):
        N = 2**J + 1
        W = WaveletTransformOp(J, interleaved=True).as_matrix()
        pattern = wavelet_dependency_pattern(J)
        steps = [_level_step(J, j) for j in range(1, J + 1)]
        X = rand((N, 3), seed=10 * J + P)
        plans = [TimeOpPlan(DofDistributionMPI(FakeComm(r, P), N, 3), pattern)
                 for r in range(P)]
        bounds = plans[0].dofs_distr.dof_distribution
        fw = [LevelChain(pl, steps) for pl in plans]
        ad = [LevelChain(pl, [G.T.tocsr() for G in reversed(steps)]) for pl in plans]
        out_w = np.zeros_like(X)
        out_wt = np.zeros_like(X)
        packed = []
        for r, pl in enumerate(plans):
            a, b = bounds[r]
            halo = np.zeros((pl.n_halo, 3))
            for p, (off, cnt) in pl.recv_from.items():  # fetch: p packs send_to[r]
                halo[off:off + cnt] = X[bounds[p][0] + plans[p].send_to[r]]
            out_w[a:b] = fw[r].apply_host(np.concatenate([X[a:b], halo]))[:b - a]
            ext = np.zeros((b - a + pl.n_halo, 3))
            ext[:b - a] = X[a:b]
            y = ad[r].apply_host(ext)
            out_wt[a:b] += y[:b - a]
            packed.append(y[b - a:])
        for r, pl in enumerate(plans):  # scatter_add_halo: r sends packed[off:off+cnt] to p
            for p, (off, cnt) in pl.recv_from.items():
                out_wt[bounds[p][0] + plans[p].send_to[r]] += packed[r][off:off + cnt]
        assert np.abs(out_w - W @ X).max() < 1e-12, (J, P)
        assert np.abs(out_wt - W.T @ X).max() < 1e-12, (J, P)


def test_locality_order_is_a_local_walk():
    """Row schedule of the SpMM kernels (stk_csr_set_row_order): a permutation
    of the rows in which the neighbours of a row are visited close to it, where
    the hierarchical numbering scatters them over the whole index range."""
    from spacetime_fullgrid_parallel_b200.linop import locality_order
    prob = SquareProblem(6, 1)
    A = prob.A_x.tocsr()
    n = A.shape[0]
    order = locality_order(A)
    assert sorted(order.tolist()) == list(range(n))
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    walk = np.abs(pos[rows] - pos[A.indices]).max()
    index = np.abs(rows - A.indices).max()
    assert walk <= 2 * int(np.sqrt(n)) and index > 10 * walk
    assert locality_order(prob.hierarchy.P_mats[-1]) is None  # not square


def test_tridiag_pair_coefficients():
    """Coefficient table of stk_time_tridiag_pair (both brackets of the Schur
    operator, heateq_mpi.py:166-178) for whole time axes and for slabs: applied
    to [previous | local | next] slices it reproduces the local rows of
    A_t X0 + L_t X1 and L_t^T X0 + M_t X1; pads and missing neighbours carry
    zero coefficients."""
    import types
    from spacetime_fullgrid_parallel_b200.heateq_mpi import SchurOperatorMPI
    prob = SquareProblem(1, 4)
    N = prob.N
    mats = [sp.csr_matrix(T) for T in (prob.A_t, prob.L_t, prob.L_t.T, prob.M_t)]
    X0, X1 = rand((N, ), seed=1), rand((N, ), seed=2)
    for a, b in ((0, N), (0, 8), (8, N), (5, 6)):
        n = b - a
        ld = max(4, (n + 3) // 4 * 4)
        stub = types.SimpleNamespace(
            dofs_distr=types.SimpleNamespace(t_begin=a, t_end=b, N=N), _tri=mats)
        c = SchurOperatorMPI._tridiag_coef(stub, ld, 'cpu').numpy()
        assert c.shape == (12, ld) and not c[:, n:].any()
        if a == 0:
            assert not c[0::3, 0].any()  # no previous slice: no sub-diagonal
        if b == N:
            assert not c[2::3, n - 1].any()
        ext = [np.concatenate([[X[a - 1] if a > 0 else 0.0], X[a:b],
                               [X[b] if b < N else 0.0]]) for X in (X0, X1)]
        t = np.arange(n)
        for k, ref in ((0, (mats[0] @ X0 + mats[1] @ X1)[a:b]),
                       (6, (mats[2] @ X0 + mats[3] @ X1)[a:b])):
            y = sum(c[k + 3 * v + o, :n] * ext[v][t + o]
                    for v in (0, 1) for o in (0, 1, 2))
            assert np.abs(y - ref).max() < 1e-14
