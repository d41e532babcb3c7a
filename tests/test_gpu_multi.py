"""GPU, >= 2 devices: runs tests/run_multi_gpu.py under torchrun (NCCL)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('nproc', [2, 4])
def test_multi_gpu_parity(cuda, nproc):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip('needs %d GPUs' % nproc)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', str(nproc), '--master-addr', '127.0.0.1',
           '--master-port', str(29500 + nproc),
           os.path.join(ROOT, 'tests', 'run_multi_gpu.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and 'MGPU OK' in out.stdout, (
        out.stdout[-3000:] + out.stderr[-3000:])
