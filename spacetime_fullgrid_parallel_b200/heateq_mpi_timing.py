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
import argparse
import base64
import pickle
import sys
import zlib

import numpy as np

from .comm import Wtime, init_from_env
from .heateq_mpi import HeatEquationMPI, mem
from .mpi_kron import LinearOperatorMPI
from .mpi_vector import KronVectorMPI

# algorithmic bytes per space-time dof of one apply:
# W1 = 16; regrouped S = 2 MG (523) + split/brackets/pair (24+2*24+24);
# P = 2 MG + one SpMM
ALG_BYTES = {'W': 16.0, 'WT': 16.0, 'S': 2 * 523.0 + 96.0, 'P': 2 * 523.0 + 16.0}


def main(argv=None):
    parser = argparse.ArgumentParser(
        description='Time the operators of the heat equation on B200s.')
    parser.add_argument('--problem', default='square')
    parser.add_argument('--J_time', type=int, default=7)
    parser.add_argument('--J_space', type=int, default=7)
    parser.add_argument('--smoothsteps', type=int, default=3)
    parser.add_argument('--vcycles', type=int, default=2)
    parser.add_argument('--wavelettransform', default='original')
    parser.add_argument('--alpha', type=float, default=0.3)
    parser.add_argument('--iters', type=int, default=10)
    args = parser.parse_args(argv)

    import torch
    comm = init_from_env()
    rank, size = comm.Get_rank(), comm.Get_size()
    data = {'rank': rank, 'size': size}
    if size > 2**args.J_time + 1:
        print('Too many MPI processors!')
        sys.exit('1')

    heq = HeatEquationMPI(J_space=args.J_space, J_time=args.J_time,
                          problem=args.problem, smoothsteps=args.smoothsteps,
                          vcycles=args.vcycles, alpha=args.alpha,
                          wavelettransform=args.wavelettransform, comm=comm)
    if rank == 0:
        data['args'] = vars(args)
        data['N'], data['M'] = heq.N, heq.M
        print('\n\nCreating mesh with {} time refines and {} space refines.'.
              format(args.J_time, args.J_space))
        print('MPI tasks: ', size)
        print('Arguments:', args)
        print('N = {}. M = {}.'.format(heq.N, heq.M))
        print('Constructed bilinear forms in {} s.'.format(heq.setup_time))
        print('Memory after construction: {}mb.'.format(mem()))
    data['mem_after_construction'] = mem()

    LinearOperatorMPI.sync_timing = True
    comm.Barrier()
    time_total = Wtime()
    vec = KronVectorMPI(heq.dofs_distr)
    np.random.seed(128)  # heateq_mpi_timing.py:82
    vec.X_loc[:] = np.random.rand(*vec.X_loc.shape)
    dofs_loc = vec.n_loc * vec.M
    for name in ('W', 'S', 'WT', 'P'):
        op = getattr(heq, name)
        op @ vec  # untimed warm-up: graph capture, NCCL connections, allocator
        op.num_applies -= 1
        op.time_applies = op.time_communication = 0
        times, comms = [], []
        time_total_op = Wtime()
        for _ in range(args.iters):
            t_a, t_c = op.time_applies, op.time_communication
            vec._invalidate()
            op @ vec
            times.append(op.time_applies - t_a)
            comms.append(op.time_communication - t_c)
            comm.Barrier()
        data[name] = {
            'time_applies': op.time_applies,
            'time_communication': op.time_communication,
            'time_applies_iter': times,
            'time_communication_iter': comms,
            'num_applies': op.num_applies,
            'time_total': Wtime() - time_total_op,
            'alg_GBs': ALG_BYTES[name] * dofs_loc / np.median(times) / 1e9,
        }
    comm.Barrier()
    data['time_total'] = Wtime() - time_total
    data['mem_after_timing'] = mem()
    data['gpu_mem_reserved_gb'] = torch.cuda.max_memory_reserved() / 1e9
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
    LinearOperatorMPI.sync_timing = False
    data = comm.gather(data, root=0)
    if rank == 0:
        print('\ndata: {}'.format(
            str(base64.b64encode(zlib.compress(pickle.dumps(data))), 'ascii')))
    return data


if __name__ == '__main__':
    main()
