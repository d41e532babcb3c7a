"""TEST INFRASTRUCTURE: NumPy interpreter of a fused Gauss-Seidel program
(spacetime_fullgrid_parallel_b200/gs_program.py) with the exact semantics of
`k_gs_fused` (csrc/stk_gsfused.cu): per item a private window of slots; per
macro-step the loads issued LOOKAHEAD steps earlier are visible, all ops of the
step (all its passes) read the window as it was at the start of the step (they
run concurrently on the device) and then write their own slot; ops with the
store flag write the global result.  Used by the CPU tests to check the
compiler against the sequential sweep of oracle/gs.c."""
import numpy as np

from spacetime_fullgrid_parallel_b200.gs_program import LOOKAHEAD, PREFETCH


def emulate(prog, indptr, data, diag, f, u_in, check_hazards=True,
            indices=None):
    data = np.asarray(data)[prog.canon]  # the program's entry order
    """u_out (n, k) after the program's sweeps; f, u_in: (n, k) (u_in None =
    zero guess); data: CSR values on the program's pattern."""
    n, k = f.shape
    u_out = np.full((n, k), np.nan)
    nnz_row = np.diff(indptr)
    op = prog.records_aos()
    G = prog.ngrp
    for it in range(prog.nitems):
        win = np.full((prog.nslots, k), np.nan)
        landing = {}  # step -> (slots, values) becoming visible after that step
        s0, s1 = prog.item_step[it], prog.item_step[it + 1]
        pa, la = prog.item_pass[it], (prog.step_info[s0 - 1, 1] if s0 else 0)
        p_end_item = prog.item_pass[it + 1]
        for m in range(s0, s1):
            pb, lb = prog.step_info[m]
            if lb > la:
                rows = prog.ld[la:lb, 0]
                slots = prog.ld[la:lb, 1].astype(np.int64)
                vals = np.zeros((lb - la, k)) if u_in is None else u_in[rows]
                if check_hazards:
                    # a slot being (re)loaded must not be read until it lands
                    win[slots] = np.nan
                landing.setdefault(m + LOOKAHEAD, []).append((slots, vals))
            if pb > pa:
                o = op[pa * G:pb * G]
                # the f prefetch hint names the record PREFETCH passes later
                idx = np.arange(pa * G, pb * G) + PREFETCH * G
                ok = idx < p_end_item * G
                assert np.array_equal(
                    o[ok, 1], op[idx[ok], 0] & np.uint32(0x7fffffff))
                live = (o[:, 3] >> 16) > 0
                o = o[live]
                row = (o[:, 0] & 0x7fffffff).astype(np.int64)
                store = (o[:, 0] >> 31).astype(bool)
                self_slot = (o[:, 3] & 0xffff).astype(np.int64)
                nnz = (o[:, 3] >> 16).astype(np.int64)
                assert np.array_equal(nnz, nnz_row[row])
                assert len(np.unique(self_slot)) == len(self_slot)
                new = np.empty((len(o), k))
                slots16 = np.ascontiguousarray(o[:, 4:]).view(np.uint16)
                for q in range(len(o)):
                    i = row[q]
                    sl = slots16[q, :nnz[q]].astype(np.int64)
                    vals = data[indptr[i]:indptr[i + 1]]
                    if prog.generic:
                        assert int(o[q, 2]) == indptr[i]
                    # the diagonal entry first; unused entries = own slot
                    assert sl[0] == self_slot[q]
                    assert (slots16[q, nnz[q]:] == self_slot[q]).all()
                    ax = vals @ win[sl]
                    new[q] = win[self_slot[q]] + (f[i] - ax) / diag[i]
                assert not np.isnan(new).any(), 'op read an invalid window slot'
                win[self_slot] = new
                u_out[row[store]] = new[store]
            # end of step m: loads issued at m - LOOKAHEAD have landed
            for slots, vals in landing.pop(m, []):
                win[slots] = vals
            pa, la = pb, lb
    assert not np.isnan(u_out).any(), 'some row was never stored'
    return u_out
