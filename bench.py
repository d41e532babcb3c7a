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
`roofline`: the dominant kernel (k_gs_phase, one Gauss-Seidel wavefront of the
            finest level), timed with CUDA events inside this process.
`cpu_baseline`: the oracle (CPU port of the reference algorithm) on the host
            cores, on a bounded sample; reported, not the target.
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


def run_reference(args):
    """--impl reference: the reference algorithm on the host cores (oracle
    port -- the reference itself is Python + NGSolve/PETSc/MPI and cannot run
    on the box), one worker process per time slab on all host cores (its
    mpirun -np P decomposition), same metric/unit/config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    step, dofs, pool, threads = cpu_iteration_factory_mp(os.cpu_count() or 1)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    pool.close()
    ms = el / args.steps * 1e3
    val = dofs / (ms * 1e-3)
    line = {
        'impl': 'reference',
        'metric': 'space-time DoF-applies/sec per PCG iteration',
        'value': val, 'unit': 'DoF-applies/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD % (args.J_time, args.J_space),
                   'J_time': args.J_time, 'J_space': args.J_space},
        'cpu_baseline': {
            'value': val, 'unit': 'DoF-applies/s', 'cores': threads,
            'kind': 'port',
            'sample': ('each step = one PCG iteration\'s operator applies of '
                       'the oracle at J_time=%d J_space=%d (%d dofs), time '
                       'slabs on %d worker processes' %
                       (SAMPLE_JT, SAMPLE_JS, dofs, threads))
        },
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

    heq = HeatEquationMPI(J_space=args.J_space, J_time=args.J_time, comm=comm)
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

    # ---- roofline of the dominant kernel: one GS wavefront, finest level ----
    fam = heq.family
    top = len(fam.num_phases) - 1
    nph = fam.num_phases[top]
    ld = heq.rhs.ld
    ctx = fam.context([(0.0, 1.0)], ld)
    u = torch.zeros_like(heq.rhs.data)
    f = heq.rhs.data
    reps = 5

    def sweeps(n):
        check(lib().stk_mg_smooth(fam.handle, top, n, 0, ptr(ctx.coef[0]),
                                  ptr(ctx.coef[1]), ptr(f), ptr(u), ld,
                                  stream()))

    sweeps(2)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(
        enable_timing=True)
    g0.record()
    sweeps(reps)
    g1.record()
    torch.cuda.synchronize()
    launch_ms = g0.elapsed_time(g1) / (reps * nph)
    # G1 of SURVEY.md 8(d): 24 B per level-dof per sweep (u read, f read, u
    # written), n_loc live time slices per row.
    bytes_per_launch = 24.0 * heq.M * heq.rhs.n_loc / nph
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    peak, which = peaks()
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get('k_gs_phase_bytes_per_launch')
        except Exception:
            traffic = None
    roofline = {
        'kernel': ('k_gs_phase4<2,false> (one Gauss-Seidel wavefront, finest '
                   'level, per-slice coefficients)'),
        'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
        'frac': achieved / peak, 'traffic': traffic, 'peak_source': which,
        'launch_ms': launch_ms, 'algorithmic_bytes_per_launch': bytes_per_launch
    }
    del u

    if args.no_e2e:  # profiling runs: skip the full solve and the CPU leg
        if rank == 0:
            print(json.dumps({'ms_per_step': ms_per_step, 'value': value,
                              'gpu_launches': int(launches),
                              'roofline': roofline, 'note': 'profiling run'}),
                  flush=True)
        return

    # ---- e2e: the user's call, host buffers in and out ----
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
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
    if size > 1:
        import torch.distributed as dist
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    e2e = {
        'value': D * iters / e2e_s, 'unit': 'DoF-applies/s',
        'h2d_bytes_per_step': 8.0 * D / iters,
        'd2h_bytes_per_step': 8.0 * D / iters,
        'solve_seconds': e2e_s, 'pcg_iterations': iters,
        'call': 'KronVectorMPI(host rhs) -> PCG(WT_S_W, P, rhs) -> host w'
    }

    if rank != 0:
        return
    cpu = cpu_baseline(threads=1) if size == 1 and not args.no_cpu else None
    line = {
        'metric': 'space-time DoF-applies/sec per PCG iteration',
        'value': value, 'unit': 'DoF-applies/s', 'n_gpus': size,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {
            'workload': WORKLOAD % (args.J_time, args.J_space),
            'J_time': args.J_time, 'J_space': args.J_space, 'N': heq.N,
            'M': heq.M, 'dofs': D, 'smoothsteps': 3, 'vcycles': 2,
            'alpha': 0.3, 'wavelettransform': 'composite',
            'parallelism': 'time-slab x%d' % size,
            'l2': 'inputs larger than L2 (one vector = %.2f GB), no flush' %
            (8e-9 * D / size)
        },
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
        'roofline': roofline, 'cpu_baseline': cpu,
    }
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
