"""ORACLE / TEST INFRASTRUCTURE ONLY -- the CPU arm of bench.py.

Times the UNMODIFIED reference classes (oracle/_ref/source or
/root/reference/source: mpi_kron, mpi_vector, wavelets, multigrid, linalg's
recurrences) on the host cores, one emulated MPI rank per core through the
mpi4py / petsc4py stand-ins (oracle/standins; MatSOR is the plain-C sweep of
oracle/gs.c), exactly the decomposition of `mpirun -np P heateq_mpi.py`:
every rank owns a time slab, applies W, S, WT and P to it and exchanges halos
and wavelet rows with its peers.  A step is one full PCG iteration
(/root/reference/source/linalg.py:26-40), timed between barriers as the
reference times its solve (heateq_mpi.py:281-288).
"""
import time


def rank_main(J_time, J_space, steps, warmup):
    """Runs on every emulated rank; returns (seconds, dofs, ranks)."""
    from oracle import ref_harness
    ref_harness.activate()
    from mpi4py import MPI
    from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
    comm = MPI.COMM_WORLD
    prob = SquareProblem(J_space, J_time)
    g = ref_harness.RefGraph(prob)
    T, P, b = g.WT_S_W, g.P, g.rhs

    state = {}

    def restart():
        from source.mpi_vector import KronVectorMPI
        state['w'] = KronVectorMPI(b.dofs_distr)
        state['r'] = b - T @ state['w']
        state['p'] = P @ state['r']
        state['abs_r'] = state['r'].dot(state['p'])
        state['abs_r0'] = state['abs_r']

    def step():  # linalg.py:26-40, statement by statement
        w, r, p = state['w'], state['r'], state['p']
        t = T @ p
        alpha = state['abs_r'] / p.dot(t)
        w += alpha * p
        r -= alpha * t
        z = P @ r
        abs_r_old = state['abs_r']
        state['abs_r'] = r.dot(z)
        if not state['abs_r'] > 1e-24 * state['abs_r0']:
            restart()  # converged far below eps^2: start the next solve
            return
        beta = state['abs_r'] / abs_r_old
        p *= beta
        p += z

    restart()
    for _ in range(warmup):
        step()
    comm.Barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    comm.Barrier()
    return time.perf_counter() - t0, prob.N * prob.M, comm.Get_size()


def run(J_time, J_space, ranks, steps, warmup):
    from oracle import ref_harness
    ref_harness.activate()
    from mpi4py import MPI
    out = MPI.launch(ranks, rank_main, J_time, J_space, steps, warmup)
    seconds = max(o[0] for o in out)
    return seconds, out[0][1]
