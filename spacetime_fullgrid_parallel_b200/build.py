"""Builds libstk.so (the sm_100a device library) in-tree with nvcc.

`python -m spacetime_fullgrid_parallel_b200.build` or `build_lib()`.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libstk.so')
SOURCES = ['stk_blas1.cu', 'stk_kron.cu', 'stk_wavelet.cu', 'stk_mg.cu',
           'stk_gsfused.cu']
NVCC_FLAGS = [
    '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-Xcompiler', '-fPIC', '-shared', '--fmad=true',
    '-I' + os.path.join(ROOT, 'include'), '-I' + CSRC
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand)
                     or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def sources():
    deps = [os.path.join(CSRC, s) for s in SOURCES]
    deps += [os.path.join(CSRC, 'stk_common.cuh'),
             os.path.join(ROOT, 'include', 'stk.h')]
    return deps


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in sources())


def build_lib(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + [
        '-o', LIB
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    build_lib(force=True, verbose='-v' in sys.argv)
    print(LIB)
