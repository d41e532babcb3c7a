"""Multi-GPU parity script (one process per GPU, NCCL):

    torchrun --nproc-per-node P --master-addr 127.0.0.1 tests/run_multi_gpu.py

Every rank builds the graph of heateq_mpi.py:126-191 on its time slab and
compares its slab of W, S, WT, P, WT_S_W applied to the seed-128 vector, the
PCG iteration count and the residual history with the golden outputs of the
unmodified reference classes (tests/golden/graph.npz).  Exercises the halo
exchange, the one-shot wavelet boundary exchange and its adjoint, the
all-to-all transpose and the scalar allreduce over NVLink.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from conftest import rand, rel  # noqa: E402


def main():
    import torch
    from spacetime_fullgrid_parallel_b200 import comm as stk_comm
    from spacetime_fullgrid_parallel_b200.heateq_mpi import HeatEquationMPI
    from spacetime_fullgrid_parallel_b200.linalg import PCG
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
    comm = stk_comm.init_from_env()
    rank, size = comm.Get_rank(), comm.Get_size()
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'graph.npz'))
    worst = 0.0
    cases = [(3, 3, 'composite', 'Jt3_Js3_composite_P1'),
             (3, 3, 'original', 'Jt3_Js3_original_P3'),
             (3, 3, 'interleaved', 'Jt3_Js3_composite_P1'),
             (4, 2, 'composite', 'Jt4_Js2_composite_P4'),
             (3, 6, 'composite', 'Jt3_Js6_composite_P1')]
    for Jt, Js, mode, tag in cases:
        if size > 2**Jt + 1:
            continue
        heq = HeatEquationMPI(J_space=Js, J_time=Jt, wavelettransform=mode,
                              comm=comm)
        a, b = heq.dofs_distr.t_begin, heq.dofs_distr.t_end
        X = rand((heq.N, heq.M))
        x = KronVectorMPI(heq.dofs_distr, X[a:b])
        big = Js >= 6
        for name in ('W', 'S', 'WT', 'P', 'WT_S_W'):
            y = getattr(heq, name) @ x
            if big:
                nrm = np.sqrt(y.dot(y))
                err = abs(nrm - float(g['%s__norm_%s' % (tag, name)])) / nrm
            else:
                err = rel(y.X_loc, g['%s__%s' % (tag, name)][a:b])
            worst = max(worst, err)
            assert err < 1e-12, (tag, mode, name, err, rank)
        rr = []
        w, iters = PCG(heq.WT_S_W, heq.P, heq.rhs,
                       callback=lambda w, r, k: rr.append(r.dot(r)))
        ref_iters = int(g[tag + '__iters'])
        assert abs(iters - ref_iters) <= 1, (tag, iters, ref_iters)
        n = min(iters, ref_iters)
        ref_rr = g[tag + '__hist_rr']
        assert np.max(np.abs(np.sqrt(rr[:n]) - np.sqrt(ref_rr[:n]))) < (
            1e-10 * np.sqrt(ref_rr[0])), tag
        u = heq.W @ w
        ref = float(g[tag + '__norm_u'])
        assert abs(np.sqrt(u.dot(u)) - ref) < 1e-10 * ref, tag
        # halo API (mpi_vector.py:140-203): neighbour slices and arbitrary
        # remote slices arrive as device rows
        prev, nxt = x.communicate_bdr()
        if a > 0:
            assert np.array_equal(prev.cpu().numpy(), X[a - 1]), (tag, 'bdr')
        if b < heq.N:
            assert np.array_equal(nxt.cpu().numpy(), X[b]), (tag, 'bdr')
        want = [t for t in (0, heq.N // 2, heq.N - 1) if not a <= t < b]
        got = x.communicate_dofs([(a, t) for t in want])
        assert sorted(got) == want, (tag, 'dofs', sorted(got), want)
        for t in want:
            assert np.array_equal(got[t].cpu().numpy(), X[t]), (tag, 'dofs')
        # public permute round trip across ranks
        p = x.permute()
        assert np.array_equal(
            np.asarray(p.X_loc), X.T[p.t_begin:p.t_end]), (tag, 'permute')
        if rank == 0:
            print('case %s (%s) on %d GPUs: iters %d (ref %d), worst apply '
                  'err %.2e' % (tag, mode, size, iters, ref_iters, worst),
                  flush=True)
    # precond='direct' (heateq_mpi.py:154-157) on the live communicator
    gd = np.load(os.path.join(ROOT, 'tests', 'golden', 'direct.npz'))
    tag = 'direct_Jt3_Js3_composite_P2'
    if size <= 9:
        heq = HeatEquationMPI(J_space=3, J_time=3, precond='direct', comm=comm)
        a, b = heq.dofs_distr.t_begin, heq.dofs_distr.t_end
        x = KronVectorMPI(heq.dofs_distr, rand((heq.N, heq.M))[a:b])
        for name in ('S', 'P', 'WT_S_W'):
            err = rel((getattr(heq, name) @ x).X_loc, gd['%s__%s' % (tag, name)][a:b])
            assert err < 1e-12, (tag, name, err, rank)
        w, iters = PCG(heq.WT_S_W, heq.P, heq.rhs)
        assert abs(iters - int(gd[tag + '__iters'])) <= 1, (tag, iters)
        if rank == 0:
            print('case %s on %d GPUs: iters %d' % (tag, size, iters), flush=True)
    torch.cuda.synchronize()
    comm.Barrier()
    if rank == 0:
        print('MGPU OK size=%d' % size, flush=True)


if __name__ == '__main__':
    main()
