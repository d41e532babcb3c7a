"""GPU: the C ABI entry points called directly (ctypes) against numpy/scipy on
small ragged shapes -- pads, halo columns, accumulate modes, coefficient
families, empty inputs."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rand, rel

pytestmark = pytest.mark.gpu


def _env():
    import torch
    from spacetime_fullgrid_parallel_b200._lib import check, lib, ptr
    from spacetime_fullgrid_parallel_b200.mpi_vector import pitch
    return torch, check, lib(), ptr, pitch


def _block(torch, X, ld):
    """(n_t, M) host -> (M, ld) device block with zero pads."""
    n_t, M = X.shape
    b = torch.zeros((M, ld), dtype=torch.float64, device='cuda')
    b[:, :n_t] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    return b


def _csr(torch, A):
    A = sp.csr_matrix(A, dtype=np.float64)
    A.sort_indices()
    return (torch.from_numpy(A.indptr.astype(np.int32)).cuda(),
            torch.from_numpy(A.indices.astype(np.int32)).cuda(),
            torch.from_numpy(A.data.astype(np.float64)).cuda(), A)


@pytest.mark.parametrize('n_t,M', [(1, 5), (2, 1), (7, 33), (33, 129)])
def test_space_spmm_family_and_modes(cuda, n_t, M):
    torch, check, L, ptr, pitch = _env()
    ld = pitch(n_t)
    rs = np.random.RandomState(n_t * 100 + M)
    A0 = sp.random(M, M, density=0.3, random_state=rs, format='csr') + sp.identity(M)
    A1 = sp.random(M, M, density=0.3, random_state=rs, format='csr')
    pat = (abs(A0) + abs(A1)).tocsr()
    pat.sort_indices()
    ip, ix, _, _ = _csr(torch, pat)

    def on_pattern(A):
        out = pat.copy()
        out.data[:] = 0
        out = (out + A).tocsr()  # scipy drops nothing that pat stores? rebuild:
        D = A.toarray()
        rows = np.repeat(np.arange(M), np.diff(pat.indptr))
        return torch.from_numpy(D[rows, pat.indices].copy()).cuda()

    v0, v1 = on_pattern(A0), on_pattern(A1)
    X, Z = rs.rand(n_t, M), rs.rand(n_t, M)
    c0, c1 = rs.rand(n_t) + 0.5, rs.rand(n_t) + 0.5
    x, z = _block(torch, X, ld), _block(torch, Z, ld)
    y = torch.full((M, ld), np.nan, dtype=torch.float64, device='cuda')
    pad = lambda c: torch.from_numpy(np.concatenate([c, np.full(ld - n_t, c[-1])])).cuda()
    d0, d1 = pad(c0), pad(c1)
    # K = 2, y = 2 (c0 A0 + c1 A1) x - 3 z
    check(L.stk_space_spmm(M, ptr(ip), ptr(ix), 2, ptr(v0), ptr(v1), ptr(d0), ptr(d1), ptr(x),
                           2.0, -3.0, ptr(z), ptr(y), ld, None))
    ref = np.stack([2.0 * ((c0[t] * A0 + c1[t] * A1) @ X[t]) - 3.0 * Z[t] for t in range(n_t)])
    got = y.cpu().numpy()
    assert rel(got[:, :n_t].T, ref) < 1e-14
    assert np.all(got[:, n_t:] == 0.0)  # pads written as zero (z pads are zero)
    # K = 1, beta = 0 into garbage output
    y.fill_(np.nan)
    check(L.stk_space_spmm(M, ptr(ip), ptr(ix), 1, ptr(v0), None, None, None, ptr(x), 1.0, 0.0,
                           None, ptr(y), ld, None))
    got = y.cpu().numpy()
    assert rel(got[:, :n_t].T, (A0 @ X.T).T) < 1e-14 and np.all(got[:, n_t:] == 0.0)
    # split / pair on the shared pattern
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    check(L.stk_space_spmm_split(M, ptr(ip), ptr(ix), ptr(v0), ptr(v1), ptr(x), ptr(y0),
                                 ptr(y1), ld, None))
    assert rel(y0.cpu().numpy()[:, :n_t].T, (A0 @ X.T).T) < 1e-14
    assert rel(y1.cpu().numpy()[:, :n_t].T, (A1 @ X.T).T) < 1e-14
    check(L.stk_space_spmm_pair(M, ptr(ip), ptr(ix), ptr(v0), ptr(v1), ptr(x), ptr(z), ld, 1.5,
                                0.5, ptr(y0), ptr(y), ld, None))
    ref = 1.5 * ((A0 @ X.T).T + (A1 @ Z.T).T) + 0.5 * (A0 @ X.T).T
    assert rel(y.cpu().numpy()[:, :n_t].T, ref) < 1e-14
    # aliasing is refused, not silently wrong
    assert L.stk_space_spmm(M, ptr(ip), ptr(ix), 1, ptr(v0), None, None, None, ptr(x), 1.0, 0.0,
                            None, ptr(x), ld, None) != 0
    assert b'alias' in L.stk_last_error()


def test_row_order_changes_nothing_but_the_walk(cuda):
    """stk_csr_set_row_order: every SpMM entry point gives bit-identical
    results with a registered row schedule (rows are independent), for the
    plain, residual (z), split and pair forms; removal restores index order."""
    torch, check, L, ptr, pitch = _env()
    n_t, M = 9, 257
    ld = pitch(n_t)
    rs = np.random.RandomState(5)
    A0 = sp.random(M, M, density=0.05, random_state=rs, format='csr') + sp.identity(M)
    ip, ix, v0, A = _csr(torch, A0)
    v1 = torch.from_numpy(rs.rand(A.nnz)).cuda()
    x, z = _block(torch, rs.rand(n_t, M), ld), _block(torch, rs.rand(n_t, M), ld)

    def run_all():
        y = [torch.full((M, ld), np.nan, dtype=torch.float64, device='cuda') for _ in range(5)]
        check(L.stk_space_spmm(M, ptr(ip), ptr(ix), 1, ptr(v0), None, None, None, ptr(x), 1.0,
                               0.0, None, ptr(y[0]), ld, None))
        check(L.stk_space_spmm(M, ptr(ip), ptr(ix), 1, ptr(v0), None, None, None, ptr(x), 1.0,
                               -1.0, ptr(z), ptr(y[1]), ld, None))
        check(L.stk_space_spmm_split(M, ptr(ip), ptr(ix), ptr(v0), ptr(v1), ptr(x), ptr(y[2]),
                                     ptr(y[3]), ld, None))
        check(L.stk_space_spmm_pair(M, ptr(ip), ptr(ix), ptr(v0), ptr(v1), ptr(x), ptr(z), ld,
                                    1.0, 0.0, None, ptr(y[4]), ld, None))
        return [t.cpu().numpy() for t in y]

    base = run_all()
    assert rel(base[0][:, :n_t], A @ x.cpu().numpy()[:, :n_t]) < 1e-14
    order = torch.from_numpy(rs.permutation(M).astype(np.int32)).cuda()
    check(L.stk_csr_set_row_order(ptr(ip), M, ptr(order)))
    try:
        for a, b in zip(base, run_all()):
            assert np.array_equal(a, b)
        # an order registered for another row count is ignored, not misapplied
        check(L.stk_csr_set_row_order(ptr(ip), M - 1, ptr(order)))
        for a, b in zip(base, run_all()):
            assert np.array_equal(a, b)
    finally:
        check(L.stk_csr_set_row_order(ptr(ip), 0, None))
    for a, b in zip(base, run_all()):
        assert np.array_equal(a, b)


def test_time_tridiag_pair(cuda):
    """stk_time_tridiag_pair: both brackets of the Schur operator
    (heateq_mpi.py:166-178) for four tridiagonal time matrices in one pass,
    on whole time axes and on slabs with a previous / next neighbour slice,
    odd and even slab lengths, one to 17 warps per row, all three block-size
    variants."""
    import types
    torch, check, L, ptr, pitch = _env()
    from spacetime_fullgrid_parallel_b200.heateq_mpi import SchurOperatorMPI
    rs = np.random.RandomState(11)
    M = 37
    cases = ((9, 0, 9), (9, 0, 4), (9, 4, 9), (5, 2, 3), (257, 0, 257), (257, 224, 257),
             (257, 64, 128), (257, 128, 257), (130, 0, 66), (130, 66, 130), (1025, 0, 1025))
    for N, a, b in cases:
        T = [sp.diags([rs.rand(N - 1), rs.rand(N), rs.rand(N - 1)], [-1, 0, 1], format='csr')
             for _ in range(4)]
        n = b - a
        ld = pitch(n)
        stub = types.SimpleNamespace(
            dofs_distr=types.SimpleNamespace(t_begin=a, t_end=b, N=N), _tri=T)
        coef = SchurOperatorMPI._tridiag_coef(stub, ld, 'cuda')
        X0, X1 = rs.rand(N, M), rs.rand(N, M)
        x0, x1 = _block(torch, X0[a:b], ld), _block(torch, X1[a:b], ld)

        def row(X, t):
            return torch.from_numpy(X[t].copy()).cuda() if 0 <= t < N else None

        p0, n0, p1, n1 = row(X0, a - 1), row(X0, b), row(X1, a - 1), row(X1, b)
        if a == 0:
            p0 = p1 = None
        y = torch.full((M, 2 * ld), np.nan, dtype=torch.float64, device='cuda')
        check(L.stk_time_tridiag_pair(M, n, ld, ptr(coef), ptr(x0), ptr(x1), ld, ptr(p0),
                                      ptr(n0), ptr(p1), ptr(n1), y.data_ptr(),
                                      y.data_ptr() + 8 * ld, 2 * ld, None))
        got = y.cpu().numpy()
        ref1 = (T[0] @ X0 + T[1] @ X1)[a:b]
        ref2 = (T[2] @ X0 + T[3] @ X1)[a:b]
        assert rel(got[:, :n].T, ref1) < 1e-14, (N, a, b)
        assert rel(got[:, ld:ld + n].T, ref2) < 1e-14, (N, a, b)
        assert np.all(got[:, n:ld] == 0.0) and np.all(got[:, ld + n:] == 0.0), (N, a, b)
    # pitches the kernel does not cover come back as an error, not as garbage
    assert L.stk_time_tridiag_pair(M, 2049, 2052, ptr(coef), ptr(x0), ptr(x1), 2052, None, None,
                                   None, None, y.data_ptr(), y.data_ptr() + 8, 4104, None) != 0


@pytest.mark.parametrize('dense', [False, True])
def test_time_apply_with_halo(cuda, dense):
    """Local columns + slice-major halo columns, overwrite and accumulate,
    through the global-memory kernel (tridiagonal) and the shared-memory one
    (dense rows)."""
    torch, check, L, ptr, pitch = _env()
    n, nh, M = 13, 3, 71
    ld = pitch(n)
    rs = np.random.RandomState(7)
    if dense:
        T = sp.csr_matrix(rs.rand(n, n + nh))
    else:
        T = sp.hstack([sp.diags([rs.rand(n - 1), rs.rand(n), rs.rand(n - 1)],
                                [-1, 0, 1]), sp.csr_matrix((n, nh))]).tolil()
        T[0, n] = 0.7  # previous rank's last slice
        T[n - 1, n + 1] = -1.3  # next rank's first slice
        T[5, n + 2] = 2.0
        T = T.tocsr()
    ip, ix, iv, T = _csr(torch, T)
    X, H, Y0 = rs.rand(n, M), rs.rand(nh, M), rs.rand(n, M)
    x = _block(torch, X, ld)
    xh = torch.from_numpy(H).cuda()
    y = torch.full((M, ld), np.nan, dtype=torch.float64, device='cuda')
    check(L.stk_time_apply(M, n, T.nnz, ptr(ip), ptr(ix), ptr(iv), ptr(x), ld, n, ptr(xh), nh,
                           2.0, 0.0, ptr(y), ld, None))
    ref = 2.0 * (T @ np.concatenate([X, H]))
    got = y.cpu().numpy()
    assert rel(got[:, :n].T, ref) < 1e-14 and np.all(got[:, n:] == 0.0)
    y = _block(torch, Y0, ld)
    check(L.stk_time_apply(M, n, T.nnz, ptr(ip), ptr(ix), ptr(iv), ptr(x), ld, n, ptr(xh), nh,
                           1.0, 0.5, ptr(y), ld, None))
    got = y.cpu().numpy()
    assert rel(got[:, :n].T, 0.5 * Y0 + T @ np.concatenate([X, H])) < 1e-14
    assert np.all(got[:, n:] == 0.0)


def test_time_apply2_and_empty_rows(cuda):
    torch, check, L, ptr, pitch = _env()
    n, M = 9, 40
    ld = pitch(n)
    rs = np.random.RandomState(3)
    Ta = sp.diags([rs.rand(n - 1), rs.rand(n), rs.rand(n - 1)], [-1, 0, 1], format='csr')
    Tb = sp.diags([rs.rand(n - 1), rs.rand(n)], [1, 0], format='csr')
    ha, hb = rs.rand(2, M), rs.rand(1, M)
    # stacked columns: [xa | xb | halo a (2) | halo b (1)]
    S = sp.hstack([Ta, Tb, sp.csr_matrix((n, 3))]).tolil()
    S[0, 2 * n] = 1.1
    S[n - 1, 2 * n + 1] = -0.4
    S[3, 2 * n + 2] = 0.9
    ip, ix, iv, S = _csr(torch, S.tocsr())
    Xa, Xb = rs.rand(n, M), rs.rand(n, M)
    xa, xb = _block(torch, Xa, ld), _block(torch, Xb, ld)
    y = torch.full((M, ld), np.nan, dtype=torch.float64, device='cuda')
    dha, dhb = torch.from_numpy(ha).cuda(), torch.from_numpy(hb).cuda()
    check(L.stk_time_apply2(M, n, ptr(ip), ptr(ix), ptr(iv), ptr(xa), ptr(xb), ld, n, ptr(dha),
                            2, ptr(dhb), 1.0, 0.0, ptr(y), ld, ld, None))
    ref = S @ np.concatenate([Xa, Xb, ha, hb])
    got = y.cpu().numpy()
    assert rel(got[:, :n].T, ref) < 1e-14 and np.all(got[:, n:] == 0.0)
    # G_t = e0 e0^T accumulated: only slice 0 changes (heateq_mpi.py:88)
    G = sp.csr_matrix(([1.0], ([0], [0])), shape=(n, n))
    ip, ix, iv, G = _csr(torch, G)
    Y0 = rs.rand(n, M)
    y = _block(torch, Y0, ld)
    check(L.stk_time_apply(M, n, G.nnz, ptr(ip), ptr(ix), ptr(iv), ptr(xa), ld, n, None, 0, 1.0,
                           1.0, ptr(y), ld, None))
    ref = Y0.copy()
    ref[0] += Xa[0]
    assert rel(y.cpu().numpy()[:, :n].T, ref) < 1e-15


def test_pack_unpack_and_empty(cuda):
    torch, check, L, ptr, pitch = _env()
    n, M = 6, 50
    ld = pitch(n)
    X = rand((n, M), seed=4)
    x = _block(torch, X, ld)
    idx = torch.tensor([4, 0, 5], dtype=torch.int32, device='cuda')
    out = torch.empty((3, M), dtype=torch.float64, device='cuda')
    check(L.stk_pack_slices(ptr(x), ld, M, ptr(idx), 3, ptr(out), None))
    assert np.array_equal(out.cpu().numpy(), X[[4, 0, 5]])
    check(L.stk_unpack_slices(ptr(x), ld, M, ptr(idx), 3, ptr(out), 2.0, 1.0, None))
    ref = X.copy()
    ref[[4, 0, 5]] += 2.0 * X[[4, 0, 5]]
    assert rel(x.cpu().numpy()[:, :n].T, ref) < 1e-15
    # empty inputs are no-ops
    check(L.stk_pack_slices(ptr(x), ld, M, ptr(idx), 0, ptr(out), None))
    check(L.stk_axpy(1.0, ptr(x), ptr(x), 0, None))
    check(L.stk_space_spmm(0, None, None, 1, None, None, None, None, ptr(x), 1.0, 0.0, None,
                           ptr(out), ld, None))
    check(L.stk_wavelet_lift(0, 3, 0, ptr(x), ptr(x), 12, None))
    # bad arguments come back as error codes with a message
    assert L.stk_axpy(1.0, ptr(x), ptr(x), 3, None) != 0
    assert L.stk_wavelet_lift(4, 3, 0, ptr(x), ptr(x), 8, None) != 0  # pitch < 2^J + 1
    assert len(L.stk_last_error()) > 0


def test_host_upload_download_entry_points(cuda):
    """stk_block_upload_host / _download_host with HOST pointers (the e2e
    boundary) on ragged shapes."""
    torch, check, L, ptr, pitch = _env()
    for n_t, M in ((1, 1), (3, 1000), (65, 37)):
        ld = pitch(n_t)
        X = rand((n_t, M), seed=n_t)
        blk = torch.full((M, ld), np.nan, dtype=torch.float64, device='cuda')
        tmp = torch.empty(n_t * M, dtype=torch.float64, device='cuda')
        check(L.stk_block_upload_host(X.ctypes.data, n_t, M, ptr(blk), ld, ptr(tmp), None))
        got = blk.cpu().numpy()
        assert np.array_equal(got[:, :n_t], X.T) and np.all(got[:, n_t:] == 0.0)
        back = np.empty_like(X)
        check(L.stk_block_download_host(ptr(blk), ld, n_t, M, back.ctypes.data, ptr(tmp), None))
        assert np.array_equal(back, X)


CHAIN_CASES = ((3, 2), (5, 4), (8, 8))


def test_time_chain_emulated_ranks(cuda):
    """stk_time_chain on the slab of every rank of emulated decompositions
    (halo slices supplied by hand): W rows and the adjoint with its halo
    partial sums, against the sparse matrix."""
    torch, check, L, ptr, pitch = _env()
    from test_host_logic import FakeComm
    from spacetime_fullgrid_parallel_b200.mpi_vector import DofDistributionMPI
    from spacetime_fullgrid_parallel_b200.timeop import LevelChain, TimeOpPlan
    from spacetime_fullgrid_parallel_b200.wavelets import (
        WaveletTransformOp, _level_step, wavelet_dependency_pattern)
    M = 77
    for J, P in CHAIN_CASES:
        N = 2**J + 1
        W = WaveletTransformOp(J, interleaved=True).as_matrix()
        pattern = wavelet_dependency_pattern(J)
        X = rand((N, M), seed=J)
        steps = [_level_step(J, j) for j in range(1, J + 1)]
        WX = W @ X
        for r in range(P):
            d = DofDistributionMPI(FakeComm(r, P), N, M)
            pl = TimeOpPlan(d, pattern)
            a, b = d.t_begin, d.t_end
            n, ld = b - a, pitch(b - a)
            x = _block(torch, X[a:b], ld)
            halo = torch.from_numpy(np.ascontiguousarray(X[pl.halo_cols])).cuda()
            y = torch.full((M, ld), np.nan, dtype=torch.float64, device='cuda')
            LevelChain(pl, steps).apply(x, ld, M, y, ld, halo if pl.n_halo else None, None)
            got = y.cpu().numpy()
            assert rel(got[:, :n].T, WX[a:b]) < 1e-13, (J, P, r)
            assert np.all(got[:, n:] == 0.0)
            ad = LevelChain(pl, [G.T.tocsr() for G in reversed(steps)])
            yh = torch.full((max(pl.n_halo, 1), M), np.nan, dtype=torch.float64, device='cuda')
            ad.apply(x, ld, M, y, ld, None, yh if pl.n_halo else None)
            contrib = W[a:b].T @ X[a:b]
            assert rel(y.cpu().numpy()[:, :n].T, contrib[a:b]) < 1e-13
            if pl.n_halo:
                assert rel(yh.cpu().numpy(), contrib[pl.halo_cols]) < 1e-13
