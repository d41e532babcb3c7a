"""Times the multi-GPU pieces with device synchronisation (diagnosis):
    torchrun --nproc-per-node P tools/diag_comm.py --J_time 10 --J_space 10"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--J_time', type=int, default=10)
    ap.add_argument('--J_space', type=int, default=10)
    args = ap.parse_args()
    from spacetime_fullgrid_parallel_b200 import comm as stk_comm
    from spacetime_fullgrid_parallel_b200.heateq_mpi import HeatEquationMPI
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
    comm = stk_comm.init_from_env()
    rank = comm.Get_rank()
    heq = HeatEquationMPI(J_space=args.J_space, J_time=args.J_time, comm=comm)
    x = KronVectorMPI(heq.dofs_distr)
    x.data[:, :x.n_loc] = torch.rand((heq.M, x.n_loc), dtype=torch.float64, device='cuda')

    def sync():
        torch.cuda.synchronize()
        comm.Barrier()
        torch.cuda.synchronize()

    def timeit(name, fn, reps=3):
        fn()
        sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        sync()
        dt = (time.perf_counter() - t0) / reps
        if rank == 0:
            print('%-28s %9.2f ms' % (name, dt * 1e3), flush=True)

    pl = heq.W.plan
    if rank == 0 and pl is not None:
        print('W plan: n_halo', pl.n_halo, 'recv_from', {p: c for p, (o, c) in pl.recv_from.items()},
              'send_to', {p: len(v) for p, v in pl.send_to.items()}, flush=True)
    y = x.empty_like()

    def w_fetch():
        x._invalidate()
        pl.fetch(x)

    if pl is not None:
        timeit('W halo fetch (pack+exchange)', w_fetch)
    timeit('W apply', lambda: (x._invalidate(), heq.W._matvec(x, y)))
    timeit('WT apply', lambda: (x._invalidate(), heq.WT._matvec(x, y)))
    timeit('S apply', lambda: (x._invalidate(), heq.S._matvec(x, y)))
    timeit('P apply', lambda: heq.P._matvec(x, y))
    timeit('dot', lambda: x.dot(y))
    sync()
    if rank == 0:
        print('max reserved GB', torch.cuda.max_memory_reserved() / 1e9, 'allocated',
              torch.cuda.max_memory_allocated() / 1e9, flush=True)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
