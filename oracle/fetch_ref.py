"""ORACLE / TEST INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED
reference classes available where /root/reference does not exist (the GPU box).

    python -m oracle.fetch_ref

copies the pure-Python hot-path modules of /root/reference/source verbatim
into oracle/_ref/source/ (git-ignored, so no reference source enters the
history; not gpurun-ignored, so it travels to the GPU box like a built .so).
`bench.py --impl reference` then times those classes -- through the
mpi4py / petsc4py stand-ins of oracle/standins -- on the box's host cores.
Nothing in the product package reads oracle/_ref.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get('STK_REFERENCE_ROOT', '/root/reference'),
                   'source')
DST = os.path.join(HERE, '_ref', 'source')
MODULES = ['lanczos.py', 'linalg.py', 'linop.py', 'mpi_kron.py',
           'mpi_shared_mem.py', 'mpi_vector.py', 'multigrid.py',
           'wavelets.py']


def fetch():
    """Returns True if oracle/_ref/source is in place afterwards."""
    if os.path.isdir(SRC):
        os.makedirs(DST, exist_ok=True)
        for name in MODULES:
            shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        open(os.path.join(DST, '__init__.py'), 'a').close()
    return all(os.path.exists(os.path.join(DST, m)) for m in MODULES)


if __name__ == '__main__':
    print('oracle/_ref/source:', 'ready' if fetch() else 'reference not present')
