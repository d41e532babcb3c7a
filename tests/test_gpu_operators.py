"""GPU parity of the Kronecker operators, vectors and wavelets: the cases of
the reference's own tests (source/mpi_kron_test.py, mpi_vector_test.py,
wavelets_test.py) run through the device classes, plus the golden outputs of
the unmodified reference classes (tests/golden, oracle/gen_golden.py).

Tolerance: BASELINE.json north_star -- every operator apply within 1e-12
relative in the 2-norm.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rand, rel

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _stk():
    import spacetime_fullgrid_parallel_b200 as stk  # noqa: F401
    from spacetime_fullgrid_parallel_b200 import (comm, linalg, linop,
                                                  mpi_kron, mpi_vector,
                                                  multigrid, wavelets)
    return comm, linalg, linop, mpi_kron, mpi_vector, multigrid, wavelets


def _distr(N, M):
    comm, _, _, _, mv, _, _ = _stk()
    return mv.DofDistributionMPI(comm.SerialComm(), N, M)


def _vec(d, X=None):
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
    return KronVectorMPI(d, X)


def linearity(linop):
    """mpi_kron_test.py:12-28."""
    d = linop.dofs_distr
    rs = np.random.RandomState(1)
    x = _vec(d, rs.rand(d.N, d.M))
    y = _vec(d, rs.rand(d.N, d.M))
    z = x + 3.14 * y
    r1 = linop @ x + 3.14 * (linop @ y)
    r2 = linop @ z
    assert rel(r1.X_loc, r2.X_loc) < 1e-13


def linop_check(linop, mat_glob):
    """mpi_kron_test.py:31-36 with the 1e-12 bar instead of np.allclose."""
    linearity(linop)
    mat = linop.as_global_matrix()
    assert rel(mat, mat_glob) < TOL


FIX5 = np.array([[3.5, 13., 28.5, 50., 77.5], [-5., -23., -53., -95., -149.],
                 [2.5, 11., 25.5, 46., 72.5]])


def stiff5():
    return sp.spdiags(FIX5, (1, 0, -1), 5, 5).T.copy().tocsr()


@pytest.mark.parametrize('n_t,M', [(1, 1), (5, 7), (9, 13), (33, 1000),
                                   (4, 3)])
def test_roundtrip_and_blas1(cuda, n_t, M):
    d = _distr(n_t, M)
    rs = np.random.RandomState(0)
    X, Y = rs.rand(n_t, M), rs.rand(n_t, M)
    x, y = _vec(d, X), _vec(d, Y)
    assert np.array_equal(np.asarray(x.X_loc), X)
    # pads stay zero
    assert float(x.data[:, n_t:].abs().sum()) == 0.0
    assert abs(x.dot(y) - np.dot(X.ravel(), Y.ravel())) <= 1e-13 * abs(
        np.dot(X.ravel(), Y.ravel()))  # mpi_vector_test.py:7-28
    z = x + 2.5 * y
    assert rel(z.X_loc, X + 2.5 * Y) < 1e-15
    z -= y
    z *= 3.0
    z /= 4.0
    assert rel(z.X_loc, (X + 1.5 * Y) * 3.0 / 4.0) < 1e-15
    w = x.copy()
    w.X_loc[:] = Y
    assert np.array_equal(np.asarray(w.X_loc), Y)
    assert np.array_equal(np.asarray(x.X_loc), X)
    w.X_loc[0] = 7.0
    Y2 = Y.copy()
    Y2[0] = 7.0
    assert np.array_equal(np.asarray(w.X_loc), Y2)


def test_dot_large(cuda):
    d = _distr(257, 4099)
    rs = np.random.RandomState(5)
    X, Y = rs.randn(257, 4099), rs.randn(257, 4099)
    x, y = _vec(d, X), _vec(d, Y)
    ref = np.dot(X.ravel(), Y.ravel())
    for _ in range(3):  # ticket counter resets itself
        assert abs(x.dot(y) - ref) < 1e-12 * np.sqrt(X.size)
    assert abs(x.dot(x) - np.dot(X.ravel(), X.ravel())) < 1e-12 * X.size


def test_identity_kron_mat(cuda):
    """mpi_kron_test.py:39-45."""
    _, _, _, mk, _, _, _ = _stk()
    N, M = 13, 16
    d = _distr(N, M)
    mat_space = np.arange(0, M * M).reshape(M, M).astype(float)
    linop_check(mk.IdentityKronMatMPI(d, mat_space),
                np.kron(np.eye(N), mat_space))


def test_mat_kron_identity(cuda):
    """mpi_kron_test.py:48-54."""
    _, _, _, mk, _, _, _ = _stk()
    N, M = 9, 16
    d = _distr(N, M)
    mat_time = np.arange(0, N * N).reshape(N, N).astype(float)
    linop_check(mk.MatKronIdentityMPI(d, mat_time),
                np.kron(mat_time, np.eye(M)))


def test_tridiag_and_sparse_kron_identity(cuda):
    """mpi_kron_test.py:57-78."""
    _, _, _, mk, _, _, _ = _stk()
    T = stiff5()
    d = _distr(5, 3)
    linop_check(mk.TridiagKronIdentityMPI(d, T),
                np.kron(T.toarray(), np.eye(3)))
    linop_check(mk.SparseKronIdentityMPI(d, T),
                np.kron(T.toarray(), np.eye(3)))
    linop_check(mk.SparseKronIdentityMPI(d, T, add_identity=True),
                np.kron(T.toarray() + np.eye(5), np.eye(3)))


def test_tridiag_kron_mat(cuda):
    _, _, _, mk, _, _, _ = _stk()
    T = stiff5()
    M = 6
    A = sp.random(M, M, density=0.5, random_state=3, format='csr') + sp.identity(M)
    d = _distr(5, M)
    op = mk.TridiagKronMatMPI(d, T, A.tocsr())
    linop_check(op, np.kron(T.toarray(), A.toarray()))
    assert rel(op.as_matrix(), np.kron(T.toarray(), A.toarray())) < 1e-15


def test_block_diag(cuda):
    """mpi_kron_test.py:82-94 (seed 0): distinct matrix per slice."""
    _, _, _, mk, _, _, _ = _stk()
    np.random.seed(0)
    N, M = 9, 4
    mats = [sp.csr_matrix(np.random.rand(M, M)) for _ in range(N)]
    d = _distr(N, M)
    import scipy.linalg
    linop_check(mk.BlockDiagMPI(d, mats),
                scipy.linalg.block_diag(*[m.toarray() for m in mats]))
    same = mk.BlockDiagMPI(d, [mats[0]] * N)
    linop_check(same, np.kron(np.eye(N), mats[0].toarray()))


def test_composite_and_sum(cuda):
    """mpi_kron_test.py:97-128."""
    _, _, _, mk, _, _, _ = _stk()
    T = stiff5()
    M = 3
    d = _distr(5, M)
    A = np.arange(1, M * M + 1).reshape(M, M).astype(float)
    T_I = mk.TridiagKronIdentityMPI(d, T)
    I_A = mk.IdentityKronMatMPI(d, A)
    kTI, kIA = np.kron(T.toarray(), np.eye(M)), np.kron(np.eye(5), A)
    linop_check(mk.CompositeMPI(d, [T_I, I_A]), kTI @ kIA)
    linop_check(mk.CompositeMPI(d, [I_A, T_I, I_A]), kIA @ kTI @ kIA)
    linop_check(mk.SumMPI(d, [T_I, I_A]), kTI + kIA)
    linop_check(mk.SumMPI(d, [T_I, I_A, mk.IdentityMPI(d)]),
                kTI + kIA + np.eye(5 * M))


@pytest.mark.parametrize('N,M', [(4, 244), (9, 13), (24, 244)])
def test_permute(cuda, N, M):
    """mpi_vector_test.py:31-48: permute == reshape(N, M).T."""
    d = _distr(N, M)
    X = np.random.RandomState(2).rand(N, M)
    p = _vec(d, X).permute()
    assert p.N == M and p.M == N
    assert np.array_equal(np.asarray(p.X_loc), X.T)
    assert np.array_equal(np.asarray(p.permute().X_loc), X)


@pytest.mark.parametrize('J', [1, 2, 3, 4, 5, 6])
def test_wavelets_golden(cuda, golden, J):
    """Reference WaveletTransformOp outputs (wavelets.py:106-134)."""
    _, _, _, _, _, _, wv = _stk()
    g = golden['wavelets']
    X = rand((2**J + 1, 3), seed=J)
    for inter, tag in ((True, 'int'), (False, 'lvl')):
        op = wv.WaveletTransformOp(J, interleaved=inter)
        assert np.array_equal(op.levels, g['levels_J%d_%s' % (J, tag)])
        assert rel(op @ X, g['W_J%d_%s' % (J, tag)]) < TOL
        assert rel(op.T @ X, g['WT_J%d_%s' % (J, tag)]) < TOL
        # the sparse-matrix form used on P > 1 ranks
        assert rel(op.as_matrix() @ X, g['W_J%d_%s' % (J, tag)]) < TOL
        assert rel(op.T.as_matrix() @ X, g['WT_J%d_%s' % (J, tag)]) < TOL
    # device operator on a sharded vector (1 rank): W (x) I and its transpose
    M = 5
    d = _distr(2**J + 1, M)
    Y = rand((2**J + 1, M), seed=40 + J)
    Wd = wv.WaveletTransformKronIdentityMPI(d, J)
    WTd = wv.TransposedWaveletTransformKronIdentityMPI(d, J)
    Wm = wv.WaveletTransformOp(J, interleaved=True).as_matrix()
    assert rel((Wd @ _vec(d, Y)).X_loc, Wm @ Y) < TOL
    assert rel((WTd @ _vec(d, Y)).X_loc, Wm.T @ Y) < TOL


def test_wavelet_known_answers(cuda):
    """wavelets_test.py:22-72: shapes of the coarsest wavelets, the explicit
    J = 2 matrix and W = prod (I + split_j)."""
    _, _, _, _, _, _, wv = _stk()
    for J in range(1, 8):
        N = 2**J + 1
        op = wv.WaveletTransformOp(J)
        E = np.eye(N)
        W = op @ E
        assert np.allclose(W[:, 0], np.linspace(1, 0, N), atol=1e-14)
        assert np.allclose(W[:, 1], np.linspace(0, 1, N), atol=1e-14)
        half = 2**(J - 1)
        expect = np.concatenate([np.linspace(-np.sqrt(2), np.sqrt(2), half + 1),
                                 np.linspace(np.sqrt(2), -np.sqrt(2),
                                             half + 1)[1:]])
        assert np.allclose(W[:, 2], expect, atol=1e-13)
        if J >= 2:
            assert np.isclose(W[0, 3], -2.0) and np.isclose(W[-1, 4], -2.0)
    op = wv.WaveletTransformOp(2, interleaved=True)
    s2 = np.sqrt(2)
    expect = np.array([[1, -2, -s2, 0, 0], [3 / 4, 2, 0, 0, 1 / 4],
                       [1 / 2, -1, s2, -1, 1 / 2], [1 / 4, 0, 0, 2, 3 / 4],
                       [0, 0, -s2, -2, 1]])  # wavelets_test.py:48-51
    assert np.allclose(op @ np.eye(5), expect, atol=1e-14)
    for J in range(1, 7):
        op = wv.WaveletTransformOp(J, interleaved=True)
        prod = np.eye(2**J + 1)
        for j in range(1, J + 1):
            prod = (np.eye(2**J + 1) + op.split(j).toarray()) @ prod
        assert np.allclose(prod, op @ np.eye(2**J + 1), atol=1e-13)


def test_wavelet_split_golden(cuda, golden):
    _, _, _, _, _, _, wv = _stk()
    op = wv.WaveletTransformOp(4, interleaved=True)
    for j in range(1, 5):
        assert rel(op.split(j).toarray(),
                   golden['wavelets']['split_J4_j%d' % j]) < 1e-15


def test_wavelet_transform_large(cuda):
    """Round trip property at BASELINE size J_time = 9 / 10: W^T W against the
    sparse matrix form, on a block wider than one CTA's share."""
    _, _, _, _, _, _, wv = _stk()
    for J in (9, 10):
        N, M = 2**J + 1, 300
        d = _distr(N, M)
        X = rand((N, M), seed=J)
        Wm = wv.WaveletTransformOp(J, interleaved=True).as_matrix()
        y = wv.WaveletTransformKronIdentityMPI(d, J) @ _vec(d, X)
        assert rel(y.X_loc, Wm @ X) < TOL
        z = wv.TransposedWaveletTransformKronIdentityMPI(d, J) @ y
        assert rel(z.X_loc, Wm.T @ (Wm @ X)) < TOL


def test_serial_glue_and_shared_matrices(cuda):
    """linop.py:6-15 (KronLinOp), :68-79 (CompositeLinOp), mpi_shared_mem.py
    (device-resident matrices), wavelets.py:9-42 (WaveletTransformMat)."""
    from spacetime_fullgrid_parallel_b200.linop import (CompositeLinOp,
                                                        KronLinOp)
    from spacetime_fullgrid_parallel_b200.mpi_kron import as_matrix
    from spacetime_fullgrid_parallel_b200.mpi_shared_mem import (
        shared_numpy_array, shared_sparse_matrix)
    from spacetime_fullgrid_parallel_b200.wavelets import (WaveletTransformMat,
                                                           WaveletTransformOp)
    T = stiff5()
    A = sp.random(7, 7, density=0.5, random_state=1, format='csr') + sp.identity(7)
    B = sp.random(7, 7, density=0.5, random_state=2, format='csr')
    x = rand((35, ), seed=3)
    assert rel(KronLinOp(T, A.tocsr()) @ x,
               np.kron(T.toarray(), A.toarray()) @ x) < TOL
    # any sparse time factor, not only tridiagonal ones (linop.py:6-15)
    Wt = WaveletTransformOp(2, interleaved=True)
    assert rel(KronLinOp(Wt, A.tocsr()) @ x,
               np.kron(Wt.as_matrix().toarray(), A.toarray()) @ x) < TOL
    dA = shared_sparse_matrix(A.tocsr())
    assert dA.shape == (7, 7) and rel(dA @ np.eye(7), A.toarray()) < 1e-15
    comp = CompositeLinOp([dA, B.tocsr(), dA])
    assert rel(as_matrix(comp), (A @ B @ A).toarray()) < 1e-14
    v = shared_numpy_array(np.arange(5.0))
    assert v.is_cuda and v.cpu().tolist() == [0, 1, 2, 3, 4]
    for J in (1, 3, 5):
        assert rel(WaveletTransformMat(J).toarray(),
                   WaveletTransformOp(J) @ np.eye(2**J + 1)) < 1e-13


def test_halo_api_single_rank(cuda):
    """communicate_bdr / communicate_dofs exist and are empty on one rank."""
    d = _distr(5, 3)
    v = _vec(d, rand((5, 3)))
    assert v.communicate_bdr() == (None, None)
    assert v.communicate_dofs([(0, 1), (4, 3)]) == {}
