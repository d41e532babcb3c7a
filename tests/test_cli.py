"""CPU: the drivers' command line, banner and `data:` blob (heateq_mpi.py:205-312)
with the device objects replaced by numpy stand-ins."""
import base64
import pickle
import zlib

import numpy as np
import scipy.sparse as sp

from spacetime_fullgrid_parallel_b200 import _cli, heateq_mpi


class _Op:
    """Host operator with the counters the driver serialises."""
    def __init__(self, mat):
        self.mat = mat
        self.num_applies, self.time_applies, self.time_communication = 0, 0.0, 0.0

    def __matmul__(self, x):
        self.num_applies += 1
        self.time_applies += 1e-3
        return self.mat @ x

    def time_per_apply(self):
        return (self.time_applies / max(self.num_applies, 1), 0.0)


class _FakeHeatEq:
    N, M = 5, 7
    setup_time = 0.0
    mem_after_ngsolve = mem_after_shared_matrices = mem_after_precond = 1.0

    def __init__(self):
        n = self.N * self.M
        A = sp.diags([np.full(n - 1, -1.0), np.full(n, 2.5), np.full(n - 1, -1.0)],
                     [-1, 0, 1], format='csr')
        self.W = self.WT = _Op(sp.identity(n, format='csr'))
        self.S = self.WT_S_W = _Op(A)
        self.P = _Op(sp.identity(n, format='csr'))
        self.rhs = np.ones(n)

    print_time_per_apply = heateq_mpi.HeatEquationMPI.print_time_per_apply


def test_flags_match_the_reference_defaults():
    a = _cli.parse('x', 'composite', argv=[])
    assert vars(a) == dict(problem='square', J_time=7, J_space=7, smoothsteps=3, vcycles=2,
                           alpha=0.3, wavelettransform='composite')  # heateq_mpi.py:208-234
    t = _cli.parse('x', 'original', extra=(('--iters', int, 10, ''), ), argv=['--iters', '3'])
    assert t.wavelettransform == 'original' and t.iters == 3  # heateq_mpi_timing.py:35-42


def _host_pcg(T, P, b, callback=None, eps=1e-6):
    """NumPy stand-in for the device PCG (the product has no host form): only
    here so that the driver's report can be checked without a GPU."""
    w = np.zeros_like(b)
    r = b.copy()
    p = P @ r
    rz = r @ p
    iters = 0
    while rz >= eps * eps:
        iters += 1
        t = T @ p
        a = rz / (p @ t)
        w += a * p
        r -= a * t
        if callback is not None:
            callback(w, r, iters)
        z = P @ r
        rz, rz_old = r @ z, rz
        p = z + (rz / rz_old) * p
    return w, iters


def test_driver_main_report_and_blob(monkeypatch, capsys):
    import torch
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(_cli, 'build', lambda args, comm: _FakeHeatEq())
    monkeypatch.setattr(heateq_mpi, 'PCG', _host_pcg)
    u, iters = heateq_mpi.main(['--J_time', '2', '--J_space', '1'])
    out = capsys.readouterr().out
    assert 'Completed in {} PCG steps.'.format(iters) in out and 'N = 5. M = 7.' in out
    blob = [l for l in out.splitlines() if l.startswith('data: ')][0][6:]
    records = pickle.loads(zlib.decompress(base64.b64decode(blob)))
    assert len(records) == 1
    rec = records[0]
    assert rec['rank'] == 0 and rec['size'] == 1 and rec['iters'] == iters
    assert rec['args']['J_time'] == 2 and rec['N'] == 5 and rec['M'] == 7
    for name in ('W', 'S', 'WT', 'P', 'WT_S_W'):  # heateq_mpi.py:293-300
        assert set(rec[name]) == {'time_applies', 'time_communication', 'num_applies'}
    assert rec['WT_S_W']['num_applies'] == iters  # zero start: no initial apply
    assert {'solve_time', 'mem_after_solve', 'mem_after_construction'} <= set(rec)


def test_too_many_ranks_guard(monkeypatch):
    class Big:
        def Get_size(self):
            return 6

        def Get_rank(self):
            return 0

    from spacetime_fullgrid_parallel_b200 import comm as stk_comm
    monkeypatch.setattr(stk_comm, 'init_from_env', lambda: Big())
    args = _cli.parse('x', 'composite', argv=['--J_time', '2'])  # N = 5 < 6 ranks
    try:
        _cli.start(args)
    except SystemExit as e:
        assert str(e.code) == '1'  # heateq_mpi.py:241-243
    else:
        raise AssertionError('expected SystemExit')


def test_timing_driver_report_and_blob(monkeypatch, capsys):
    """heateq_mpi_timing.py:81-128: per-operator lists in the blob."""
    import torch
    from spacetime_fullgrid_parallel_b200 import heateq_mpi_timing as timing

    class FakeVec:
        def __init__(self, dofs_distr):
            self.n_loc, self.M = 5, 7
            self._x = np.zeros((5, 7))

        @property
        def X_loc(self):
            return self._x

        def _invalidate(self):
            pass

    class VecOp(_Op):
        def __matmul__(self, v):
            self.num_applies += 1
            self.time_applies += 2e-3
            return v

    fake = _FakeHeatEq()
    fake.dofs_distr = None
    fake.W, fake.WT, fake.S, fake.P = (VecOp(None) for _ in range(4))
    monkeypatch.setattr(timing, 'KronVectorMPI', FakeVec)
    monkeypatch.setattr(_cli, 'build', lambda args, comm: fake)
    monkeypatch.setattr(torch.cuda, 'max_memory_reserved', lambda *a: 0)
    records = timing.main(['--J_time', '2', '--J_space', '1', '--iters', '3'])
    out = capsys.readouterr().out
    assert 'Completed 3 iters steps.' in out and out.count('median') == 4
    rec = records[0]
    assert rec['args']['wavelettransform'] == 'original' and rec['args']['iters'] == 3
    for name in ('W', 'S', 'WT', 'P'):  # heateq_mpi_timing.py:104-111
        assert {'time_applies', 'time_communication', 'time_applies_iter',
                'time_communication_iter', 'num_applies', 'time_total'} <= set(rec[name])
        assert rec[name]['num_applies'] == 3 and len(rec[name]['time_applies_iter']) == 3
    from spacetime_fullgrid_parallel_b200 import comm as stk_comm
    assert stk_comm.SYNC_TIMING is False  # restored
