#!/usr/bin/env python
"""Benchmark of the space-time solve path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (all N): BASELINE.json configs[3], the full preconditioned
Schur-complement PCG solve at J_time = 8, J_space = 9 on the unit square
(N = 257 time slices x M = 1,046,529 space dofs = 268,957,953 space-time dofs,
one fp64 vector = 2.15 GB, i.e. 17x the 126 MB L2, so no L2 flush is needed
between steps).  It is the largest BASELINE configuration that fits one B200
and the one BASELINE quotes at 1/2/4/8 GPUs; total work is fixed as N grows
(strong scaling), time slabs are sharded over the ranks.

A step = one PCG iteration (linalg.py:26-40): t = WT S W p, two dots, the
fused updates, z = P r.  Metric: space-time DoF-applies per second per PCG
iteration = D / (seconds per iteration), whole job.

`value`   : device-resident iterations, CUDA events, max over ranks.
`e2e`     : the public API call a user makes -- KronVectorMPI built from a HOST
            (pinned) right-hand side, PCG(WT_S_W, P, rhs), solution read back to
            the host -- D * iterations / wall time, copies inside the timed region.
`roofline`: the dominant kernel (k_gs_fused: the nu = 3 backward Gauss-Seidel
            sweeps of the finest level in one launch), timed with CUDA events
            inside this process.
`parity`  : golden cases of the UNMODIFIED reference classes run on the live
            communicator (apply errors, iteration counts, residual histories)
            and the iteration count / solution norm of the timed solve.
`cpu_baseline`: the unmodified reference classes on one host core, on a bounded
            sample; reported, not the target.  `--impl reference` runs them on
            all host cores, one emulated MPI rank per core.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = ('BASELINE.json configs[3]: full preconditioned Schur-complement '
            'PCG solve, J_time=%d J_space=%d square')
SAMPLE_JT, SAMPLE_JS = 5, 7  # CPU sample: 33 x 65,025 = 2,145,825 dofs


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,'
         'clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.proc, self.gpu = None, gpu_index
        self.path = os.path.join('/tmp', 'stk_clocks_%d.csv' % os.getpid())

    def start(self):
        try:
            self.out = open(self.path, 'w')
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        return {
            'sm_mhz': float(np.median(sm)) if sm else None,
            'sm_max_mhz': max(smax) if smax else None,
            'samples': len(sm),
            'reasons': sorted(reasons)
        }


# --------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# --------------------------------------------------------------------------
def cpu_iteration_factory(threads):
    """One PCG iteration's operator work (WT_S_W p and P r) of the oracle at
    the sample size; returns (fn, dofs)."""
    from oracle import restate
    from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
    prob = SquareProblem(SAMPLE_JS, SAMPLE_JT)
    orc = restate.HeatEqOracle(prob, threads=threads)
    X = np.random.RandomState(128).rand(prob.N, prob.M)

    def step():
        t = orc.WT_S_W(X)
        z = orc.P(X)
        return float(X.reshape(-1) @ t.reshape(-1)) + float(
            X.reshape(-1) @ z.reshape(-1))

    return step, prob.N * prob.M


# ---- the CPU arm with one PROCESS per time slab (the reference's mpirun -np P)
_worker_oracle = None


def _worker_init():
    global _worker_oracle
    from oracle import restate
    from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
    _worker_oracle = restate.HeatEqOracle(SquareProblem(SAMPLE_JS, SAMPLE_JT))


def _worker_S(TX_slab):
    return _worker_oracle.S_slab(TX_slab)


def _worker_P(X_slab, levels_slab):
    return _worker_oracle.P_slab(X_slab, levels_slab)


def cpu_iteration_factory_mp(procs):
    """Same work as cpu_iteration_factory, slabs of time slices farmed out to
    `procs` worker processes; the time stencils and the wavelet transforms
    (1 % of the work) stay in the parent.  Returns (fn, dofs, pool)."""
    import multiprocessing as mp
    from oracle import restate
    from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
    prob = SquareProblem(SAMPLE_JS, SAMPLE_JT)
    orc = restate.HeatEqOracle(prob)
    procs = max(1, min(procs, prob.N))
    pool = mp.get_context('spawn').Pool(procs, initializer=_worker_init)
    slabs = restate.slab_bounds(prob.N, procs)
    X = np.random.RandomState(128).rand(prob.N, prob.M)

    def step():
        Y = orc.W(X)
        TX = [restate.apply_time(T, Y) for T, _ in orc.terms]
        parts = pool.starmap(_worker_S, [([tx[a:b] for tx in TX], )
                                         for a, b in slabs])
        t = orc.WT(np.concatenate(parts, axis=0))
        parts = pool.starmap(_worker_P, [(X[a:b], orc.levels[a:b])
                                         for a, b in slabs])
        z = np.concatenate(parts, axis=0)
        return float(X.reshape(-1) @ t.reshape(-1)) + float(
            X.reshape(-1) @ z.reshape(-1))

    return step, prob.N * prob.M, pool, procs


def cpu_baseline(threads=1, min_seconds=10.0):
    step, dofs = cpu_iteration_factory(threads)
    step()  # warm-up (builds nothing lazily, but pages the matrices in)
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or n >= 50:
            break
    return {
        'value': dofs * n / el,
        'unit': 'DoF-applies/s',
        'cores': threads,
        'kind': 'port',
        'sample': ('%d PCG iterations\' operator applies (WT_S_W p + P r) of '
                   'the oracle at J_time=%d J_space=%d (%d dofs), %.1f s' %
                   (n, SAMPLE_JT, SAMPLE_JS, dofs, el))
    }


def reference_available():
    from oracle import ref_harness
    return ref_harness.available()


def reference_sample(cores, J_time, J_space):
    """The bounded sample of the workload the CPU arm runs: the SAME space
    mesh (J_space: same M, same per-slice cost and cache behaviour) with the
    smallest J_time that gives every core a time slice."""
    jt = 1
    while 2**jt + 1 < cores and jt < J_time:
        jt += 1
    return jt, J_space, min(cores, 2**jt + 1)


def cpu_baseline_reference(J_space):
    """One core: the reference classes on a small time slab, a few PCG
    iterations (about 20 s)."""
    from oracle import ref_bench
    jt, js, steps = 2, min(J_space, 8), 3
    seconds, dofs = ref_bench.run(jt, js, 1, steps, 1)
    return {
        'value': dofs * steps / seconds, 'unit': 'DoF-applies/s', 'cores': 1,
        'kind': 'reference',
        'sample': ('%d full PCG iterations of the unmodified reference classes '
                   '(oracle/_ref through the mpi4py/petsc4py stand-ins) at '
                   'J_time=%d J_space=%d (%d dofs), %.1f s' %
                   (steps, jt, js, dofs, seconds))
    }


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the host
    cores.  The UNMODIFIED classes of /root/reference/source (verbatim copy in
    oracle/_ref, see oracle/fetch_ref.py) run through the mpi4py / petsc4py
    stand-ins with one emulated MPI rank per core -- the decomposition of
    `mpirun -np P heateq_mpi.py`; a step is one full PCG iteration
    (linalg.py:26-40) between barriers (heateq_mpi.py:281-288).  The sample
    keeps the workload's space mesh and shortens the time axis (stated in
    `config.sample`).  Falls back to the oracle port if the copy is missing."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    config = {'workload': WORKLOAD % (args.J_time, args.J_space),
              'J_time': args.J_time, 'J_space': args.J_space}
    if reference_available():
        from oracle import ref_bench
        jt, js, ranks = reference_sample(cores, args.J_time, args.J_space)
        seconds, dofs = ref_bench.run(jt, js, ranks, args.steps,
                                      max(args.warmup, 1))
        ms = seconds / args.steps * 1e3
        kind = 'reference'
        sample = ('each step = one full PCG iteration (T p, two dots, the '
                  'updates, P r) of the unmodified reference classes at '
                  'J_time=%d J_space=%d (%d time slices x %d space dofs = %d '
                  'dofs: the workload\'s space mesh, shorter time axis), %d '
                  'emulated MPI ranks = one per host core' %
                  (jt, js, 2**jt + 1, dofs // (2**jt + 1), dofs, ranks))
        config['sample'] = {'J_time': jt, 'J_space': js, 'dofs': dofs,
                            'ranks': ranks}
        threads = ranks
    else:
        step, dofs, pool, threads = cpu_iteration_factory_mp(cores)
        for _ in range(max(args.warmup, 1)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        ms = (time.perf_counter() - t0) / args.steps * 1e3
        pool.close()
        kind = 'port'
        sample = ('each step = one PCG iteration\'s operator applies of the '
                  'oracle port at J_time=%d J_space=%d (%d dofs), time slabs on '
                  '%d worker processes' % (SAMPLE_JT, SAMPLE_JS, dofs, threads))
        config['sample'] = {'J_time': SAMPLE_JT, 'J_space': SAMPLE_JS,
                            'dofs': dofs, 'ranks': threads}
    val = dofs / (ms * 1e-3)
    line = {
        'impl': 'reference',
        'metric': 'space-time DoF-applies/sec per PCG iteration',
        'value': val, 'unit': 'DoF-applies/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': config,
        'cpu_baseline': {'value': val, 'unit': 'DoF-applies/s',
                         'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': 'DoF-applies/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
class PCGState:
    """The loop body of PCG (linalg.py:26-40) as a resumable step."""
    def __init__(self, heq):
        from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
        self.heq = heq
        self.T, self.P, self.b = heq.WT_S_W, heq.P, heq.rhs
        self.KV = KronVectorMPI
        self.restart()

    def restart(self):
        self.w = self.KV(self.b.dofs_distr)
        self.r = self.b.copy()
        self.p = self.P @ self.r
        self.abs_r = self.r.dot(self.p)
        self.abs_r0 = self.abs_r
        self.iters = 0

    def step(self):
        from spacetime_fullgrid_parallel_b200._lib import (check, lib, ptr,
                                                           stream)
        t = self.T @ self.p
        alpha = self.abs_r / self.p.dot(t)
        check(lib().stk_pcg_update(alpha, ptr(self.p.data), ptr(t.data),
                                   ptr(self.w.data), ptr(self.r.data),
                                   self.w.numel, stream()))
        del t
        z = self.P @ self.r
        old, self.abs_r = self.abs_r, self.r.dot(z)
        self.iters += 1
        if not (self.abs_r > 1e-24 * self.abs_r0):
            # converged far below eps^2: start the next solve (counted in the
            # timed region; does not happen with the default K + W)
            self.restart()
            return
        check(lib().stk_xpay(ptr(z.data), self.abs_r / old, ptr(self.p.data),
                             self.p.numel, stream()))


def live_parity(comm):
    """Golden cases of the UNMODIFIED reference classes (tests/golden/graph.npz,
    generated by oracle/gen_golden.py) on the LIVE communicator: every rank
    checks its slab.  Tolerances of BASELINE.json's north_star: applies 1e-12
    relative in the 2-norm, PCG iterations +-1, residual history and solution
    norm 1e-10 relative.  Mirrors /root/reference/heateq_mpi_test.py:138-189
    (test_solve) and README.md:27-31 (mpirun -np 2 pytest)."""
    import torch
    from spacetime_fullgrid_parallel_b200.heateq_mpi import HeatEquationMPI
    from spacetime_fullgrid_parallel_b200.linalg import PCG
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'graph.npz'))
    size = comm.Get_size()
    out = {'cases': [], 'ok': True}
    for Jt, Js, tag in ((3, 3, 'Jt3_Js3_composite_P1'),
                        (4, 2, 'Jt4_Js2_composite_P4')):
        if size > 2**Jt + 1:
            continue
        heq = HeatEquationMPI(J_space=Js, J_time=Jt, comm=comm)
        a, b = heq.dofs_distr.t_begin, heq.dofs_distr.t_end
        X = np.random.RandomState(128).rand(heq.N, heq.M)
        x = KronVectorMPI(heq.dofs_distr, X[a:b])
        worst = 0.0
        for name in ('W', 'S', 'WT', 'P', 'WT_S_W'):
            y = getattr(heq, name) @ x
            ref = g['%s__%s' % (tag, name)]
            num = KronVectorMPI(heq.dofs_distr, np.asarray(y.X_loc) - ref[a:b])
            err = np.sqrt(num.dot(num)) / np.linalg.norm(ref)
            worst = max(worst, float(err))
        rr = []
        w, iters = PCG(heq.WT_S_W, heq.P, heq.rhs,
                       callback=lambda w, r, k: rr.append(r.dot(r)))
        ref_iters = int(g[tag + '__iters'])
        ref_rr = g[tag + '__hist_rr']
        n = min(iters, ref_iters)
        hist = float(np.max(np.abs(np.sqrt(rr[:n]) - np.sqrt(ref_rr[:n]))) /
                     np.sqrt(ref_rr[0]))
        u = heq.W @ w
        nu = float(np.sqrt(u.dot(u)))
        ref_nu = float(g[tag + '__norm_u'])
        case = {'case': tag, 'ranks': size, 'apply_err': worst,
                'iters': int(iters), 'ref_iters': ref_iters,
                'residual_history_err': hist,
                'norm_u_err': abs(nu - ref_nu) / ref_nu}
        case['ok'] = bool(worst < 1e-12 and abs(iters - ref_iters) <= 1 and
                          hist < 1e-10 and case['norm_u_err'] < 1e-10)
        out['cases'].append(case)
        out['ok'] = out['ok'] and case['ok']
        del heq, x, w, u
        torch.cuda.empty_cache()
    return out


def timed_solve(heq, comm, barrier):
    """The user's call with host buffers in and out: returns (seconds, iters,
    ||W w||_2)."""
    import torch
    from spacetime_fullgrid_parallel_b200.linalg import PCG
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
    a, b = heq.rhs.t_begin, heq.rhs.t_end
    rhs_host = torch.from_numpy(
        np.kron(heq.u0_t[a:b], heq.u0_x).reshape(-1, heq.M)).pin_memory()
    sol_host = torch.empty((b - a, heq.M), dtype=torch.float64).pin_memory()
    barrier()
    t0 = time.perf_counter()
    rhs = KronVectorMPI(heq.dofs_distr, rhs_host.numpy())
    w, iters = PCG(heq.WT_S_W, heq.P, rhs)
    w.to_host(out=sol_host.numpy())
    barrier()
    seconds = time.perf_counter() - t0
    t = torch.tensor([seconds], dtype=torch.float64, device='cuda')
    if comm.Get_size() > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u = heq.W @ w
    return float(t.item()), iters, float(np.sqrt(u.dot(u)))


def run_stk(args):
    import torch
    from spacetime_fullgrid_parallel_b200 import comm as stk_comm
    from spacetime_fullgrid_parallel_b200._lib import check, lib, ptr, stream
    from spacetime_fullgrid_parallel_b200.heateq_mpi import HeatEquationMPI
    from spacetime_fullgrid_parallel_b200.linalg import PCG
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (there is no CPU fallback; '
                         'use --impl reference for the CPU arm)')
    comm = stk_comm.init_from_env()
    rank, size = comm.Get_rank(), comm.Get_size()
    assert size == args.gpus, ('launch with torchrun --nproc-per-node %d' %
                               args.gpus)
    dev = torch.cuda.current_device()

    heq = HeatEquationMPI(J_space=args.J_space, J_time=args.J_time, comm=comm,
                          order=args.order)
    D = heq.N * heq.M
    state = PCGState(heq)

    def barrier():
        torch.cuda.synchronize()
        comm.Barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        state.step()
    sampler = ClockSampler(dev)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = lib().stk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
        enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        state.step()
    e1.record()
    barrier()
    launches = lib().stk_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64,
                      device='cuda')
    if size > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = D / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel: the fused smoother, finest level ----
    fam = heq.family
    top = len(fam.num_phases) - 1
    nph = fam.num_phases[top]
    ld = heq.rhs.ld
    ctx = heq.P._chain[0][1]  # the groups of P: one matrix per wavelet level
    u = torch.zeros_like(heq.rhs.data)
    u2 = torch.empty_like(u)
    f = heq.rhs.data
    reps = 5
    lvh = fam._levels[top]
    fl = lvh.get('fused')
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if fl is not None:
        from spacetime_fullgrid_parallel_b200.multigrid import _project
        nu = fam.smoothsteps
        vals = fl.values_for([
            _project(lvh['keys'], fam._galerkin(fam.combined(g))[top],
                     lvh['n'], lvh['pattern']) for g in ctx.groups])

        def sweeps():
            fl.sweeps(True, vals, ctx.group, f, u, u2)

        kernel = ('k_gs_fused<4,true,false,512,7> (the %d backward Gauss-Seidel '
                  'sweeps of the finest level in one launch, one matrix per '
                  'wavelet level)' % nu)
        # G1 of SURVEY.md 8(d): 24 B per level-dof per sweep (u read, f read, u
        # written) x nu sweeps, n_loc live time slices per row
        bytes_per_launch = 24.0 * nu * heq.M * heq.rhs.n_loc
        launches_per_rep = 1
        key = 'k_gs_fused_bytes_per_launch'
    else:
        def sweeps():
            check(lib().stk_mg_smooth(ctx.handle, top, 1, 0, ptr(ctx.group),
                                      ptr(f), ptr(u), ld, stream()))

        kernel = ('k_gs_phase<true,false> (one Gauss-Seidel wavefront, finest '
                  'level, one matrix per wavelet level)')
        bytes_per_launch = 24.0 * heq.M * heq.rhs.n_loc / nph
        launches_per_rep = nph
        key = 'k_gs_phase_bytes_per_launch'
    sweeps()
    sweeps()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
        enable_timing=True)
    g0.record()
    for _ in range(reps):
        sweeps()
    g1.record()
    torch.cuda.synchronize()
    launch_ms = g0.elapsed_time(g1) / (reps * launches_per_rep)
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    peak, which = peaks()
    if os.path.exists(tpath) and size == 1 and (args.J_time, args.J_space) == (8, 9):
        try:
            traffic = json.load(open(tpath)).get(key)
        except Exception:
            traffic = None
    roofline = {
        'kernel': kernel,
        'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
        'frac': achieved / peak, 'traffic': traffic, 'peak_source': which,
        'launch_ms': launch_ms, 'algorithmic_bytes_per_launch': bytes_per_launch
    }
    del u, u2

    if args.no_e2e:  # profiling runs: skip the full solve and the CPU leg
        if rank == 0:
            print(json.dumps({'ms_per_step': ms_per_step, 'value': value,
                              'order': args.order,
                              'gpu_launches': int(launches),
                              'roofline': roofline, 'note': 'profiling run'}),
                  flush=True)
        return

    # ---- e2e: the user's call, host buffers in and out ----
    e2e_s, iters, norm_u = timed_solve(heq, comm, barrier)
    e2e = {
        'value': D * iters / e2e_s, 'unit': 'DoF-applies/s',
        'h2d_bytes_per_step': 8.0 * D / iters,
        'd2h_bytes_per_step': 8.0 * D / iters,
        'solve_seconds': e2e_s, 'pcg_iterations': iters,
        'call': 'KronVectorMPI(host rhs) -> PCG(WT_S_W, P, rhs) -> host w'
    }

    # ---- parity on the live communicator ----
    parity = live_parity(comm)
    parity['bench_solve'] = {'pcg_iterations': iters, 'norm_W_w': norm_u}
    try:  # the same solve on ONE GPU (profiles/bench_solve_norms.json)
        exp = json.load(open(os.path.join(
            ROOT, 'profiles', 'bench_solve_norms.json')))['%d_%d' % (
                args.J_time, args.J_space)]
        rel = abs(norm_u - exp['norm_W_w']) / exp['norm_W_w']
        parity['bench_solve'].update(
            expected_iterations=exp['pcg_iterations'],
            norm_rel_diff_to_1gpu=rel,
            ok=bool(rel < 1e-10 and abs(iters - exp['pcg_iterations']) <= 1))
        parity['ok'] = parity['ok'] and parity['bench_solve']['ok']
    except Exception:
        pass

    # ---- the target configuration (BASELINE configs[4]) when it fits ----
    config5 = None
    N_main, M_main = heq.N, heq.M
    if size == 8 and not args.no_config5 and (args.J_time, args.J_space) == (8, 9):
        del state, heq, fam, ctx
        torch.cuda.empty_cache()
        try:
            big = HeatEquationMPI(J_space=10, J_time=10, comm=comm)
            timed_solve(big, comm, barrier)  # first solve: NCCL/graph warm-up
            s5, it5, n5 = timed_solve(big, comm, barrier)
            D5 = big.N * big.M
            config5 = {
                'workload': 'BASELINE.json configs[4]: J_time=10 J_space=10 square',
                'dofs': D5, 'solve_seconds': s5, 'pcg_iterations': it5,
                'ms_per_iteration': s5 / it5 * 1e3,
                'value': D5 * it5 / s5, 'unit': 'DoF-applies/s (e2e, host rhs -> '
                'host solution)', 'norm_W_w': n5}
            try:
                exp = json.load(open(os.path.join(
                    ROOT, 'profiles', 'bench_solve_norms.json')))['10_10']
                config5['norm_rel_diff_to_recorded'] = abs(
                    n5 - exp['norm_W_w']) / exp['norm_W_w']
            except Exception:
                pass
        except Exception as exc:  # the headline line must still be printed
            config5 = {'workload': 'BASELINE.json configs[4]: J_time=10 '
                       'J_space=10 square', 'error': repr(exc)[:300]}

    if rank != 0:
        return
    cpu = None
    if size == 1 and not args.no_cpu:
        cpu = (cpu_baseline_reference(args.J_space) if reference_available()
               else cpu_baseline(threads=1))
    line = {
        'metric': 'space-time DoF-applies/sec per PCG iteration',
        'value': value, 'unit': 'DoF-applies/s', 'n_gpus': size,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {
            'workload': WORKLOAD % (args.J_time, args.J_space),
            'J_time': args.J_time, 'J_space': args.J_space, 'N': N_main,
            'M': M_main, 'dofs': D, 'smoothsteps': 3, 'vcycles': 2,
            'alpha': 0.3, 'wavelettransform': 'composite',
            'order': args.order,
            'parallelism': 'time-slab x%d' % size,
            'l2': 'inputs larger than L2 (one vector = %.2f GB), no flush' %
            (8e-9 * D / size)
        },
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
        'roofline': roofline, 'cpu_baseline': cpu, 'parity': parity,
    }
    if config5 is not None:
        line['config5'] = config5
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='stk', choices=['stk', 'reference'])
    ap.add_argument('--J_time', type=int, default=8)
    ap.add_argument('--J_space', type=int, default=9)
    ap.add_argument('--no-e2e', dest='no_e2e', action='store_true',
                    help='profiling runs: timed iterations and roofline only')
    ap.add_argument('--no-cpu', dest='no_cpu', action='store_true',
                    help='skip the cpu_baseline leg (profiling runs)')
    ap.add_argument('--order', default='class',
                    help="numbering of the new vertices of a level (assembly.py): "
                    "'class' (4 Gauss-Seidel wavefronts), 'lex', 'random'")
    ap.add_argument('--no-config5', dest='no_config5', action='store_true',
                    help='at 8 GPUs: skip the J_time=10 J_space=10 solve')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_stk(args)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
