"""Per-operator microbenchmark: W, S, WT, P applied `--iters` times each to a
seeded random vector.

Drop-in for /root/reference/heateq_mpi_timing.py (:14-128): same flags (default
wavelet transform 'original', i.e. the all-to-all transpose path), same seed
(128), same report and `data:` blob keys (`time_applies`, `time_communication`,
`time_applies_iter`, `time_communication_iter`, `num_applies`, `time_total`,
`mem_*`).  This is BASELINE.json configs[2] (J_time=9, J_space=9 on one B200).
Each apply is bracketed by a device synchronisation so the times mean what the
reference's mean.  Added: achieved algorithmic GB/s per operator.

    python -m spacetime_fullgrid_parallel_b200.heateq_mpi_timing --J_time 9 --J_space 9
"""
import numpy as np

from . import comm as stk_comm
from .comm import Wtime
from .heateq_mpi import mem
from .mpi_kron import LinearOperatorMPI
from .mpi_vector import KronVectorMPI

# algorithmic bytes per space-time dof of one apply:
# W1 = 16; regrouped S = 2 MG (523) + split/brackets/pair (24+2*24+24);
# P = 2 MG + one SpMM
ALG_BYTES = {'W': 16.0, 'WT': 16.0, 'S': 2 * 523.0 + 96.0, 'P': 2 * 523.0 + 16.0}


def time_operator(op, vec, iters, comm):
    """`iters` applies of `op` to `vec` (heateq_mpi_timing.py:86-111), one
    untimed warm-up first (graph capture, NCCL connections, allocator)."""
    op @ vec
    op.num_applies, op.time_applies, op.time_communication = 0, 0, 0
    per_apply, per_comm = [], []
    started = Wtime()
    for _ in range(iters):
        before = (op.time_applies, op.time_communication)
        vec._invalidate()
        op @ vec
        per_apply.append(op.time_applies - before[0])
        per_comm.append(op.time_communication - before[1])
        comm.Barrier()  # wait for all other ranks as well
    return {
        'time_applies': op.time_applies,
        'time_communication': op.time_communication,
        'time_applies_iter': per_apply,
        'time_communication_iter': per_comm,
        'num_applies': op.num_applies,
        'time_total': Wtime() - started,
    }


def main(argv=None):
    import torch
    from . import _cli
    args = _cli.parse('Time the operators of the heat equation on B200s.',
                      'original', extra=(('--iters', int, 10,
                                          'number of iterations per operator'), ),
                      argv=argv)
    comm, data = _cli.start(args)
    rank = data['rank']
    heq = _cli.build(args, comm)
    if rank == 0:
        data.update(args=vars(args), N=heq.N, M=heq.M)
        print('\n\nCreating mesh with {} time refines and {} space refines.'.
              format(args.J_time, args.J_space))
        print('MPI tasks: ', data['size'])
        print('Arguments:', args)
        print('N = {}. M = {}.'.format(heq.N, heq.M))
        print('Constructed bilinear forms in {} s.'.format(heq.setup_time))
        print('Memory after construction: {}mb.'.format(mem()))
    data['mem_after_construction'] = mem()

    stk_comm.SYNC_TIMING = True
    comm.Barrier()
    t0 = Wtime()
    vec = KronVectorMPI(heq.dofs_distr)
    np.random.seed(128)  # heateq_mpi_timing.py:82
    vec.X_loc[:] = np.random.rand(*vec.X_loc.shape)
    dofs_loc = vec.n_loc * vec.M
    for name in ('W', 'S', 'WT', 'P'):
        data[name] = time_operator(getattr(heq, name), vec, args.iters, comm)
        data[name]['alg_GBs'] = (ALG_BYTES[name] * dofs_loc / np.median(
            data[name]['time_applies_iter']) / 1e9)
    comm.Barrier()
    data.update(time_total=Wtime() - t0, mem_after_timing=mem(),
                gpu_mem_reserved_gb=torch.cuda.max_memory_reserved() / 1e9)
    stk_comm.SYNC_TIMING = False
    if rank == 0:
        print('')
        print('Completed {} iters steps.'.format(args.iters))
        print('Total time: {}s.'.format(data['time_total']))
        heq.print_time_per_apply()
        for name in ('W', 'S', 'WT', 'P'):
            print('{}: median {:.3f} ms, {:.0f} GB/s algorithmic per GPU'.format(
                name, 1e3 * np.median(data[name]['time_applies_iter']),
                data[name]['alg_GBs']))
        print('Memory after solve: {}mb.'.format(mem()))
    return _cli.finish(comm, data)


if __name__ == '__main__':
    main()
