"""GPU: the chained wavelet transform on decompositions down to ONE slice per
rank (N = 9 over 8 / 9 ranks, N = 17 over 16), where the exchange lists must be
the structural closure of the lifting steps.  The ranks
are emulated on one GPU here; the same shapes run over NCCL in
tests/run_multi_gpu.py (case Jt3_Js3 on 8 ranks: one-slice slabs), green on
8 B200s (profiles/r2i_mgpu8_parity.log), and on the CPU in
tests/test_host_logic.py (test_chained_wavelet_all_ranks_with_exchange)."""
import pytest

import test_gpu_abi

pytestmark = pytest.mark.gpu


def test_time_chain_one_slice_slabs(cuda, monkeypatch):
    monkeypatch.setattr(test_gpu_abi, 'CHAIN_CASES', ((3, 8), (3, 9), (4, 16), (2, 4)))
    test_gpu_abi.test_time_chain_emulated_ranks(cuda)
