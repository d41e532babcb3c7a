"""CPU, world_size 2 and 3 over gloo: the N > 1 host path -- communicator,
halo / boundary-slice exchange plans, the time<->space all-to-all and the
scalar allreduce -- with numpy standing in for the device kernels."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, rand


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, size, port, errors):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(size))
    try:
        dist.init_process_group('gloo', rank=rank, world_size=size)
        _body(rank, size)
    except Exception:
        import traceback
        errors.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _body(rank, size):
    from spacetime_fullgrid_parallel_b200 import comm as stk_comm
    from spacetime_fullgrid_parallel_b200.mpi_vector import DofDistributionMPI
    from spacetime_fullgrid_parallel_b200.permute import PermutePlan
    from spacetime_fullgrid_parallel_b200.timeop import TimeOpPlan
    from spacetime_fullgrid_parallel_b200.wavelets import WaveletTransformOp

    comm = stk_comm.world()
    assert isinstance(comm, stk_comm.TorchComm)
    assert (comm.Get_rank(), comm.Get_size()) == (rank, size)
    assert comm.bcast('x' if rank == 0 else None) == 'x'
    got = comm.gather(rank)
    assert got == list(range(size)) if rank == 0 else got is None
    assert comm.allreduce(rank + 1.5) == sum(r + 1.5 for r in range(size))

    J, M = 4, 6
    N = 2**J + 1
    d = DofDistributionMPI(comm, N, M)
    a, b = d.t_begin, d.t_end
    X = rand((N, M), seed=9)  # every rank builds the same global vector

    # Krylov scalar: local dots + one allreduce (mpi_vector.py:205-210)
    loc = torch.tensor([float(np.dot(X[a:b].ravel(), X[a:b].ravel()))],
                       dtype=torch.float64)
    assert abs(comm.allreduce_sum(loc).item() - np.dot(X.ravel(), X.ravel())
               ) < 1e-12

    tri = sp.diags([np.arange(1., N), -np.arange(2., N + 2), np.arange(3., N + 2)],
                   [-1, 0, 1], format='csr')
    W = WaveletTransformOp(J, interleaved=True).as_matrix()
    for T in (tri, W):
        plan = TimeOpPlan(d, T)
        # forward: pack, exchange, local product on [local | halo]
        sends = {p: torch.from_numpy(np.ascontiguousarray(X[a:b][idx]))
                 for p, idx in plan.send_to.items()}
        halo = torch.zeros((plan.n_halo, M), dtype=torch.float64)
        recvs = {p: halo[off:off + cnt]
                 for p, (off, cnt) in plan.recv_from.items()}
        comm.exchange(sends, recvs)
        y = plan.local @ np.concatenate([X[a:b], halo.numpy()])
        assert np.allclose(y, (T @ X)[a:b], rtol=0, atol=1e-12)
        # adjoint: local partial sums, halo partials sent home and added
        out = plan.adj_local @ X[a:b]
        part = torch.from_numpy(np.ascontiguousarray(plan.adj_halo @ X[a:b]))
        sends = {p: part[off:off + cnt]
                 for p, (off, cnt) in plan.recv_from.items()}
        recvs = {p: torch.zeros((len(idx), M), dtype=torch.float64)
                 for p, idx in plan.send_to.items()}
        comm.exchange(sends, recvs)
        for p, idx in plan.send_to.items():
            out[idx] += recvs[p].numpy()
        assert np.allclose(out, (T.T @ X)[a:b], rtol=0, atol=1e-12)

    # time <-> space all-to-all on the block layout (mpi_vector.py:212-240)
    from spacetime_fullgrid_parallel_b200.mpi_vector import pitch
    ld = pitch(b - a)
    block = torch.zeros((M, ld), dtype=torch.float64)
    block[:, :b - a] = torch.from_numpy(X[a:b].T.copy())
    def host_copy_cols(src, c_in, n, dst, c_out):  # stand-in for stk_copy_cols
        dst[:, c_out:c_out + n] = src[:, c_in:c_in + n]

    pp = PermutePlan(d, copy_cols=host_copy_cols)
    sblock = pp.forward(block, b - a, ld)
    xa, xb = pp.space_distr.t_begin, pp.space_distr.t_end
    assert np.array_equal(sblock[:, :N].numpy(), X.T[xa:xb])
    assert float(sblock[:, N:].abs().sum()) == 0.0
    back = pp.backward(sblock, b - a, ld)
    assert torch.equal(back, block)


@pytest.mark.parametrize('size', [2, 3])
def test_multi_rank_host_path(size):
    ctx = mp.get_context('spawn')
    errors = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, errors))
             for r in range(size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    failed = []
    while not errors.empty():
        failed.append(errors.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            failed.append(('timeout', ''))
        elif p.exitcode != 0:
            failed.append(('exit', p.exitcode))
    assert not failed, failed
