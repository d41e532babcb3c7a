"""What the two drivers (heateq_mpi, heateq_mpi_timing) share: the reference's
command-line flags (heateq_mpi.py:208-234, heateq_mpi_timing.py:15-42), the
start-up banner and the final `data:` blob (base64(zlib(pickle(list of
per-rank dicts))), heateq_mpi.py:309-312)."""
import argparse
import base64
import pickle
import sys
import zlib

FLAGS = (
    ('--problem', str, 'square', 'problem type (square, cube)'),
    ('--J_time', int, 7, 'number of time refines'),
    ('--J_space', int, 7, 'number of space refines'),
    ('--smoothsteps', int, 3, 'number of smoothing steps'),
    ('--vcycles', int, 2, 'number of vcycles'),
    ('--alpha', float, 0.3, 'alpha'),
)


def parse(description, wavelettransform, extra=(), argv=None):
    parser = argparse.ArgumentParser(description=description)
    for flag, typ, default, text in FLAGS + tuple(extra):
        parser.add_argument(flag, type=typ, default=default, help=text)
    parser.add_argument('--wavelettransform', default=wavelettransform,
                        help='composite | original | interleaved')
    return parser.parse_args(argv)


def start(args):
    """Communicator from the launcher's environment, the rank guard of
    heateq_mpi.py:241-243, and the per-rank record."""
    from .comm import init_from_env
    comm = init_from_env()
    if comm.Get_size() > 2**args.J_time + 1:
        print('Too many MPI processors!')
        sys.exit('1')
    return comm, {'rank': comm.Get_rank(), 'size': comm.Get_size()}


def build(args, comm):
    from .heateq_mpi import HeatEquationMPI
    keys = ('J_space', 'J_time', 'problem', 'smoothsteps', 'vcycles', 'alpha',
            'wavelettransform')
    return HeatEquationMPI(comm=comm, **{k: getattr(args, k) for k in keys})


def finish(comm, data):
    """Gather the per-rank records on rank 0 and print the blob."""
    records = comm.gather(data, root=0)
    if comm.Get_rank() == 0:
        blob = base64.b64encode(zlib.compress(pickle.dumps(records)))
        print('\ndata: {}'.format(str(blob, 'ascii')))
    return records
