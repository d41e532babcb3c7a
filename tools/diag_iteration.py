"""Where one PCG iteration spends its time, per rank, from CUDA events (no host
synchronisation inside the measured sequence):

    python tools/diag_iteration.py [--J_time 8 --J_space 9]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/diag_iteration.py

One iteration of linalg.PCG is  t = WT S W p;  alpha = r.z / p.t;  w, r update;
z = P r;  r.z;  p = z + beta p.  The phases below are enqueued back to back
exactly as PCG enqueues them; events bracket every phase, the table is the
mean over `--reps` iterations and, for P > 1, the min / max over the ranks.
The exchanges are additionally timed in isolation with their byte counts:
GB/s per direction against the measured NVLink peer-copy rate (770 GB/s,
B200_PROFILING.md).  Writes gpurun_out/diag_iteration_P<P>.json.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

NVLINK_GBS = 770.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--J_time', type=int, default=8)
    ap.add_argument('--J_space', type=int, default=9)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--wavelettransform', default='composite')
    args = ap.parse_args()
    from spacetime_fullgrid_parallel_b200 import comm as stk_comm
    from spacetime_fullgrid_parallel_b200._lib import check, lib, ptr, stream
    from spacetime_fullgrid_parallel_b200.heateq_mpi import HeatEquationMPI
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
    comm = stk_comm.init_from_env()
    rank, size = comm.Get_rank(), comm.Get_size()
    heq = HeatEquationMPI(J_space=args.J_space, J_time=args.J_time, comm=comm,
                          wavelettransform=args.wavelettransform)
    d = heq.dofs_distr

    def vec():
        v = KronVectorMPI(d)
        v.data[:, :v.n_loc] = torch.rand((heq.M, v.n_loc), dtype=torch.float64,
                                         device='cuda')
        return v

    p, r, w, z = vec(), vec(), vec(), vec()
    a, b, t = p.empty_like(), p.empty_like(), p.empty_like()
    S = heq.S
    phases = []

    def phase(name, fn):
        phases.append((name, fn))

    phase('W p', lambda: (p._invalidate(), heq.W._matvec(p, a)))
    if hasattr(S, 'MA'):
        ld = p.ld
        st = {}

        def s_split():
            b._invalidate()
            st['mx'], st['ax'] = a.empty_like(), a.empty_like()
            pl = S.plans['A']
            if pl.n_halo and S.overlap:
                halo = pl.fetch(a, callback=lambda: S.MA.split(
                    a.data, st['mx'].data, st['ax'].data))
                pl.halo_of_image(halo, S.MA, st['mx'], st['ax'])
            else:
                S.MA.split(a.data, st['mx'].data, st['ax'].data)

        def s_brackets():
            st['y'] = torch.empty((heq.M, 2 * ld), dtype=torch.float64, device='cuda')
            st['z'] = torch.empty_like(st['y'])
            if getattr(S, '_tri', None) is not None and ld <= 2048:
                S._brackets_tridiag(st['mx'], st['ax'], st['y'], ld)
            else:
                S.bracket1.apply(st['mx'], st['ax'], st['y'].data_ptr(), ldy=2 * ld)
                S.bracket2.apply(st['mx'], st['ax'], st['y'].data_ptr() + 8 * ld,
                                 ldy=2 * ld)

        phase('S: M x, A x (+ halo of x in flight)', s_split)
        phase('S: time stencils (2 brackets)', s_brackets)
        phase('S: K on both brackets (MG, 2 V-cycles)',
              lambda: S.K.apply_block(st['y'], st['z']))
        phase('S: M z1 + A z2, + G term', lambda: (
            S.MA.pair(st['z'].data_ptr(), st['z'].data_ptr() + 8 * ld, b.data, ldx=2 * ld),
            S.plans['G'].apply(st['mx'], b.data, 1.0, 1.0)))
    else:
        phase('S', lambda: (a._invalidate(), S._matvec(a, b)))
    phase('WT', lambda: (b._invalidate(), heq.WT._matvec(b, t)))
    phase('p.t (dot + allreduce + readback)', lambda: p.dot(t))
    phase('w += a p; r -= a t', lambda: check(lib().stk_pcg_update(
        0.5, ptr(p.data), ptr(t.data), ptr(w.data), ptr(r.data), w.numel, stream())))
    phase('z = P r (2 MG + A_x)', lambda: (r._invalidate(), heq.P._matvec(r, z)))
    phase('r.z (dot + allreduce + readback)', lambda: r.dot(z))
    phase('p = z + b p', lambda: check(lib().stk_xpay(
        ptr(z.data), 0.5, ptr(p.data), p.numel, stream())))

    def run_once(events):
        for k, (_name, fn) in enumerate(phases):
            events[k].record()
            fn()
        events[len(phases)].record()

    def barrier():
        torch.cuda.synchronize()
        comm.Barrier()
        torch.cuda.synchronize()

    mk = lambda: [torch.cuda.Event(enable_timing=True) for _ in range(len(phases) + 1)]
    for _ in range(4):  # warm-up: graphs captured, NCCL connections up
        run_once(mk())
    barrier()
    acc = np.zeros(len(phases))
    total = 0.0
    for _ in range(args.reps):
        ev = mk()
        barrier()
        run_once(ev)
        torch.cuda.synchronize()
        acc += [ev[k].elapsed_time(ev[k + 1]) for k in range(len(phases))]
        total += ev[0].elapsed_time(ev[-1])
    acc /= args.reps
    total /= args.reps

    # ---- exchanges in isolation ----
    exch = {}

    def time_exchange(name, fn, nbytes):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        exch[name] = {'ms': ms, 'bytes_sent_per_rank': nbytes,
                      'GBs_per_direction': nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0}

    if size > 1:
        plS = S.plans['A'] if hasattr(S, 'plans') else None
        if plS is not None and plS.n_halo:
            nb = sum(len(v) for v in plS.send_to.values()) * heq.M * 8
            time_exchange('S halo (+-1 slice of x): pack + send/recv',
                          lambda: (a._invalidate(), plS.fetch(a)), nb)
        plW = getattr(heq.W, 'plan', None)
        if plW is not None and plW.n_halo:
            nb = sum(len(v) for v in plW.send_to.values()) * heq.M * 8
            time_exchange('W boundary slices: pack + send/recv',
                          lambda: (p._invalidate(), plW.fetch(p)), nb)
        if hasattr(heq.W, 'pplan'):
            nb = (size - 1) * (heq.M // size) * p.n_loc * 8
            time_exchange('permute (all-to-all transpose, one way)',
                          lambda: heq.W.pplan.forward(p.data, p.n_loc, p.ld), nb)
    rec = {'rank': rank, 'n_loc': p.n_loc, 'ld': p.ld, 'total_ms': total,
           'phases': {name: float(v) for (name, _), v in zip(phases, acc)},
           'exchanges': exch}
    allrec = comm.gather(rec)
    if rank == 0:
        allrec = allrec if size > 1 else [rec]
        out = {'J_time': args.J_time, 'J_space': args.J_space, 'ranks': size,
               'N': heq.N, 'M': heq.M, 'wavelettransform': args.wavelettransform,
               'nvlink_peer_copy_GBs': NVLINK_GBS, 'per_rank': allrec}
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        path = os.path.join(ROOT, 'gpurun_out', 'diag_iteration_P%d.json' % size)
        json.dump(out, open(path, 'w'), indent=1)
        print('one PCG iteration, J_time=%d J_space=%d on %d GPU(s); slices per rank %s'
              % (args.J_time, args.J_space, size, [r_['n_loc'] for r_ in allrec]))
        print('%-46s %9s %9s %9s' % ('phase (ms)', 'min', 'max', 'share'))
        tmax = max(r_['total_ms'] for r_ in allrec)
        for name, _ in phases:
            vals = [r_['phases'][name] for r_ in allrec]
            print('%-46s %9.3f %9.3f %8.1f%%' % (name, min(vals), max(vals),
                                                 100 * max(vals) / tmax))
        print('%-46s %9.3f %9.3f' % ('iteration', min(r_['total_ms'] for r_ in allrec), tmax))
        for name in allrec[0]['exchanges']:
            vals = [r_['exchanges'][name] for r_ in allrec if name in r_['exchanges']]
            worst = max(vals, key=lambda v: v['ms'])
            print('%-46s %9.3f ms  %7.1f MB sent  %6.1f GB/s per direction = %.2f of the '
                  'NVLink peer-copy rate' % (name, worst['ms'],
                                             worst['bytes_sent_per_rank'] / 1e6,
                                             worst['GBs_per_direction'],
                                             worst['GBs_per_direction'] / NVLINK_GBS))
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
