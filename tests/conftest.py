import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


def rand(shape, seed=128):
    """The seeded inputs the golden fixtures were generated from
    (oracle/gen_golden.py)."""
    return np.random.RandomState(seed).rand(*shape)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


@pytest.fixture(scope='session')
def golden():
    return {
        name: np.load(os.path.join(GOLDEN, name + '.npz'))
        for name in ('wavelets', 'multigrid', 'graph', 'lanczos', 'cube', 'direct')
    }


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from spacetime_fullgrid_parallel_b200 import _lib, build
    try:  # rebuild if a source is newer than the library (no-op otherwise)
        build.build_lib()
    except Exception:
        pass
    _lib.lib()  # fail loudly if the extension is missing on a GPU box
    return torch.device('cuda', 0)
