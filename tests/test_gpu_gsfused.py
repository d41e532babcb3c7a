"""GPU parity of the fused Gauss-Seidel smoother (csrc/stk_gsfused.cu, called
through the C ABI stk_gs_fused) against the sequential lexicographic sweep of
the oracle (oracle/gs.c = /root/reference/source/multigrid.py:89-97): every
(row, sweep) update must see the operands of the sequential sweep, whatever the
tiling.  Tolerance 1e-13 relative (the row sums are rounded in another order
and the division is a multiplication by the reciprocal diagonal)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rel

pytestmark = pytest.mark.gpu


def _level(Js):
    from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
    prob = SquareProblem(Js, 2)
    M = sp.csr_matrix(prob.M_x)
    A = sp.csr_matrix(prob.A_x)
    pat = (abs(M) + abs(A)).tocsr()
    pat.sort_indices()
    n = pat.shape[0]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(pat.indptr))
    keys = rows * n + pat.indices

    def project(m):
        m.sort_indices()
        r = np.repeat(np.arange(n, dtype=np.int64), np.diff(m.indptr))
        v = np.zeros(len(keys))
        v[np.searchsorted(keys, r * n + m.indices)] = m.data
        return v

    return pat, [project(M), project(A)], [M.diagonal(), A.diagonal()]


def _oracle(pat, vals, f, u0, nsw, backward):
    from oracle import cgs
    u = np.ascontiguousarray(u0.T.copy())
    cgs.gauss_seidel(pat.indptr.astype(np.int32), pat.indices.astype(np.int32),
                     np.ascontiguousarray(vals), np.ascontiguousarray(f.T), u,
                     nsw, backward=backward)
    return u.T


@pytest.mark.parametrize('generic', [False, True])
@pytest.mark.parametrize('Js,capacity', [(4, None), (6, 1500), (6, None)])
def test_fused_sweeps_match_sequential(cuda, Js, capacity, generic):
    import torch
    from spacetime_fullgrid_parallel_b200 import gs_program
    from spacetime_fullgrid_parallel_b200.multigrid import FusedLevel
    from spacetime_fullgrid_parallel_b200.mpi_vector import pitch
    pat, vals, diags = _level(Js)
    n = pat.shape[0]
    wave, depth = gs_program.wavefronts(pat.indptr, pat.indices)
    assert depth == 4
    nt = 21  # not a multiple of T: a partial last chunk, pads up to ld = 24
    ld = pitch(nt)
    rng = np.random.RandomState(5)
    F = rng.rand(n, nt)
    U0 = rng.rand(n, nt)
    # every slice has its own matrix c0(t) M + c1 A: 6 groups, interleaved
    c0 = 2.0**rng.randint(0, 6, size=nt)
    groups = sorted(set(c0))
    gvals = [c * vals[0] + 0.3 * vals[1] for c in groups]
    grp = np.array([groups.index(c) for c in c0] + [groups.index(c0[-1])] *
                   (ld - nt), dtype=np.int32)

    def dev(a):
        t = torch.zeros((n, ld), dtype=torch.float64, device=cuda)
        t[:, :nt] = torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
        return t

    for nsw in (3, 1):
        fl = FusedLevel(pat.indptr, pat.indices, wave, nsw, vals, cuda,
                        chunks=1, sms=3, generic=generic, capacity=capacity)
        assert fl.ok, 'program did not compile'
        if not generic:  # entries sorted by value: one kind covers the interior
            assert np.bincount(fl.kind_of_row)[fl.bulk_kind] > 0.8 * n
        if capacity is not None:
            assert fl.programs[0].nitems > 1
        f_d, u_d = dev(F), dev(U0)
        dvals = fl.values_for(gvals)
        assert dvals is not None
        dgrp = torch.from_numpy(grp).to(cuda)
        for backward in (False, True):
            for zero in (True, False):
                out = torch.full((n, ld), 7.0, dtype=torch.float64, device=cuda)
                fl.sweeps(backward, dvals, dgrp, f_d, None if zero else u_d, out)
                got = out.cpu().numpy()
                assert np.all(got[:, nt:] == 0.0), 'pads must stay zero'
                for t in (0, 7, nt - 1):
                    a = gvals[grp[t]]
                    u0 = np.zeros((n, 1)) if zero else U0[:, t:t + 1]
                    ref = _oracle(pat, a, F[:, t:t + 1], u0, nsw, backward)
                    assert rel(got[:, t:t + 1], ref) < 1e-13, (
                        Js, capacity, generic, nsw, backward, zero, t)
        # one group: the same matrix for every slice
        a = 4.0 * vals[0] + 0.3 * vals[1]
        out = torch.empty((n, ld), dtype=torch.float64, device=cuda)
        fl.sweeps(True, fl.values_for([a]), None, f_d, u_d, out)
        ref = _oracle(pat, a, F, U0, nsw, True)
        assert rel(out.cpu().numpy()[:, :nt], ref) < 1e-13
