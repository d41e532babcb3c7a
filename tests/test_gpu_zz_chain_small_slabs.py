"""GPU: the chained wavelet transform on decompositions down to ONE slice per
rank (N = 9 over 8 / 9 ranks, N = 17 over 16), where the exchange lists must be
the structural closure of the lifting steps.  Kept in
its own, last-sorting file: these shapes were fixed after the round's GPU
budget ended and are verified on the CPU by tests/test_host_logic.py
(test_chained_wavelet_all_ranks_with_exchange)."""
import pytest

import test_gpu_abi

pytestmark = pytest.mark.gpu


def test_time_chain_one_slice_slabs(cuda, monkeypatch):
    monkeypatch.setattr(test_gpu_abi, 'CHAIN_CASES', ((3, 8), (3, 9), (4, 16), (2, 4)))
    test_gpu_abi.test_time_chain_emulated_ranks(cuda)
