"""Per-kernel timings at a BASELINE size (CUDA events, inputs >> L2).
    python tools/microbench.py [--J_time 8 --J_space 9]
Prints algorithmic GB/s per kernel against MEASURED_PEAKS.json."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--J_time', type=int, default=8)
    ap.add_argument('--J_space', type=int, default=9)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    from spacetime_fullgrid_parallel_b200._lib import check, lib, ptr, stream
    from spacetime_fullgrid_parallel_b200.heateq_mpi import HeatEquationMPI
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        peak = 6650.0
    heq = HeatEquationMPI(J_space=args.J_space, J_time=args.J_time)
    D = heq.N * heq.M
    x = KronVectorMPI(heq.dofs_distr)
    x.data[:, :heq.N] = torch.rand((heq.M, heq.N), dtype=torch.float64, device='cuda')
    y = x.copy()
    fam = heq.family
    top = len(fam.num_phases) - 1
    ld = x.ld
    ctxK = fam.context([(0.0, 1.0)], ld)
    out = {}

    def timeit(name, fn, bytes_per_dof, reps=args.reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = bytes_per_dof * D / (ms * 1e-3) / 1e9
        out[name] = {'ms': ms, 'alg_GBs': gbs, 'frac': gbs / peak}
        print('%-34s %9.3f ms  %8.1f GB/s (alg, %4.0f B/dof)  %.3f of peak' %
              (name, ms, gbs, bytes_per_dof, gbs / peak), flush=True)

    if args.only == 'wav':
        timeit('wavelet W', lambda: heq.W._matvec(x, y), 16, reps=2)
        timeit('wavelet WT', lambda: heq.WT._matvec(x, y), 16, reps=2)
        return
    if args.only == 'mg':
        timeit('MG K_x apply (2 V(3,3))', lambda: heq.Kinv_x.apply_block(x.data, y.data), 523,
               reps=1)
        return
    u = torch.zeros_like(x.data)
    ctxP = heq.P._chain[0][1]  # the groups of P: one per wavelet level
    hP, hK = ctxP.handle, ctxK.handle
    timeit('gs sweep fwd finest (P groups)', lambda: check(lib().stk_mg_smooth(
        hP, top, 1, 0, ptr(ctxP.group), ptr(x.data), ptr(u), ld, stream())), 24)
    timeit('gs 3 sweeps fwd finest (P groups)', lambda: check(lib().stk_mg_smooth(
        hP, top, 3, 0, ptr(ctxP.group), ptr(x.data), ptr(u), ld, stream())), 72)
    timeit('gs 3 sweeps bwd finest (1 group)', lambda: check(lib().stk_mg_smooth(
        hK, top, 3, 1, None, ptr(x.data), ptr(u), ld, stream())), 72)
    fl = fam._levels[top].get('fused')
    if fl is not None:
        lvh = fam._levels[top]
        print('fused program (fwd):', fl.programs[0].stats, flush=True)

        def group_values(groups):
            return [_project(lvh['keys'], fam._galerkin(fam.combined(g))[top],
                             lvh['n'], lvh['pattern']) for g in groups]

        from spacetime_fullgrid_parallel_b200.multigrid import _project
        vP = fl.values_for(group_values(ctxP.groups))
        vK = fl.values_for(group_values(ctxK.groups))
        u2 = torch.zeros_like(x.data)
        timeit('FUSED gs 3 sweeps fwd zero-guess (P groups)', lambda: fl.sweeps(
            False, vP, ctxP.group, x.data, None, u2), 48)
        timeit('FUSED gs 3 sweeps fwd (P groups)', lambda: fl.sweeps(
            False, vP, ctxP.group, x.data, u, u2), 72)
        timeit('FUSED gs 3 sweeps bwd (P groups)', lambda: fl.sweeps(
            True, vP, ctxP.group, x.data, u, u2), 72)
        timeit('FUSED gs 3 sweeps bwd (1 group)', lambda: fl.sweeps(
            True, vK, None, x.data, u, u2), 72)
    if args.only == 'fused':
        timeit('MG K_x apply (2 V(3,3))', lambda: heq.Kinv_x.apply_block(x.data, y.data), 523)
        timeit('S apply', lambda: heq.S._matvec(x, y), 1300)
        timeit('P apply', lambda: heq.P._matvec(x, y), 1100)
        return
    if args.only == 'ncu':
        # one launch of every hot kernel at the bench size, for `ncu --set full`
        # (tools/ncu_r2.sh): fused smoother fwd/bwd, residual, restriction,
        # prolongation + correction, A_x / split / pair, time stencil, wavelets
        import scipy.sparse as sp
        from spacetime_fullgrid_parallel_b200.linop import DeviceCSR
        lvl = fam._levels[top]
        Ptop = DeviceCSR(sp.csr_matrix(fam.hierarchy.P_mats[top - 1]))
        Rtop = DeviceCSR(sp.csr_matrix(fam.hierarchy.R_mats[top - 1]))
        nc = Ptop.shape[1]
        uc = torch.rand((nc, ld), dtype=torch.float64, device='cuda')
        fc = torch.empty_like(uc)
        z = x.copy()
        S = heq.S
        mx, ax = x.empty_like(), x.empty_like()
        z1 = torch.empty_like(x.data)
        steps = [
            ('residual A u - f', lambda: heq.A_x.spmm(x.data, y.data, 1.0, -1.0, z.data), 24),
            ('restriction', lambda: Rtop.spmm(x.data, fc), 10),
            ('prolongation + correction', lambda: Ptop.spmm(uc, y.data, -1.0, 1.0, y.data), 18),
            ('spmm A_x', lambda: heq.A_x.spmm(x.data, y.data), 16),
            ('S: split', lambda: S.MA.split(x.data, mx.data, ax.data), 24),
            ('S: bracket', lambda: S.bracket1.apply(mx, ax, y.data), 24),
            ('S: pair', lambda: S.MA.pair(mx.data, ax.data, z1), 24),
            ('wavelet W', lambda: heq.W._matvec(x, y), 16),
            ('wavelet WT', lambda: heq.WT._matvec(x, y), 16),
        ]
        for name, fn, b in steps:
            timeit(name, fn, b, reps=1)
        return
    if args.only in ('gs', 'synth'):
        import scipy.sparse as sp
        from spacetime_fullgrid_parallel_b200.linop import DeviceCSR
        M = heq.M
        eye = DeviceCSR(sp.identity(M, format='csr'))
        band = DeviceCSR(sp.diags([np.ones(M - abs(k)) for k in range(-3, 4)],
                                  list(range(-3, 4)), format='csr'))
        far = DeviceCSR(sp.diags([np.ones(M - abs(k)) for k in (-2048, -1024, -1, 0, 1, 1024, 2048)],
                                 [-2048, -1024, -1, 0, 1, 1024, 2048], format='csr'))
        z = x.copy()
        timeit('synth identity spmm + z (3 units)', lambda: eye.spmm(x.data, y.data, 1.0, 1.0, z.data), 24)
        timeit('synth band-7 spmm + z (3 units)', lambda: band.spmm(x.data, y.data, 1.0, 1.0, z.data), 24)
        timeit('synth far-7 spmm + z (3 units)', lambda: far.spmm(x.data, y.data, 1.0, 1.0, z.data), 24)
        timeit('synth identity spmm (2 units)', lambda: eye.spmm(x.data, y.data), 16)
        return
    if top >= 2:
        n1 = fam.level_mats[0][top - 1].shape[0]
        u1 = torch.zeros((n1, ld), dtype=torch.float64, device='cuda')
        f1 = torch.rand((n1, ld), dtype=torch.float64, device='cuda')
        timeit('gs 3 sweeps fwd level-1', lambda: check(lib().stk_mg_smooth(
            hP, top - 1, 3, 0, ptr(ctxP.group), ptr(f1), ptr(u1), ld, stream())),
            72 * n1 / heq.M)
    timeit('spmm A_x (K=1)', lambda: heq.A_x.spmm(x.data, y.data), 16)
    timeit('spmm M_x (K=1)', lambda: heq.M_x.spmm(x.data, y.data), 16)
    timeit('MG K_x apply (2 V(3,3))', lambda: heq.Kinv_x.apply_block(x.data, y.data), 523)
    timeit('time tridiag A_t', lambda: heq.A_MKM.T_I._matvec(x, y), 16)
    timeit('wavelet W', lambda: heq.W._matvec(x, y), 16)
    timeit('wavelet WT', lambda: heq.WT._matvec(x, y), 16)
    if hasattr(heq.S, 'MA'):
        S = heq.S
        mx, ax = x.empty_like(), x.empty_like()
        z1 = torch.empty_like(x.data)
        timeit('S: split (Mx, Ax)', lambda: S.MA.split(x.data, mx.data, ax.data), 24)
        timeit('S: bracket (2-input time op)', lambda: S.bracket1.apply(mx, ax, y.data), 24)
        timeit('S: pair (M z1 + A z2)', lambda: S.MA.pair(mx.data, ax.data, z1), 24)
        timeit('S: G term', lambda: S.plans['G'].apply(mx, z1, 1.0, 1.0), 24)
    timeit('S apply', lambda: heq.S._matvec(x, y), 1300)
    timeit('P apply', lambda: heq.P._matvec(x, y), 1100)
    timeit('dot', lambda: x.dot_device(y), 16)
    timeit('axpy', lambda: y.axpy(0.5, x), 24)
    timeit('copy (torch clone)', lambda: x.data.clone(), 16)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'microbench.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
