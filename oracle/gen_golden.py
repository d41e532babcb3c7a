"""ORACLE / TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.npz.

Run in the build container (needs /root/reference):

    python -m oracle.gen_golden

Every array stored here is an OUTPUT OF THE UNMODIFIED REFERENCE CLASSES
(source/mpi_kron.py, mpi_vector.py, wavelets.py, multigrid.py, linalg.py,
lanczos.py) imported through the mpi4py/petsc4py stand-ins, on matrices from
the product's host assembler and seeded random inputs.  Tests regenerate the
inputs from the same seeds and compare the oracle (CPU) and the CUDA path (GPU)
against these files; nothing at test time reads /root/reference.
"""
import io
import contextlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness  # noqa: E402

ref_harness.activate()

from spacetime_fullgrid_parallel_b200.assembly import (CubeProblem,  # noqa: E402
                                                       SquareProblem)

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
SEED = 128  # heateq_mpi_timing.py:82


def rand(shape, seed=SEED):
    return np.random.RandomState(seed).rand(*shape)


def gen_wavelets():
    from source.wavelets import WaveletTransformOp
    out = {}
    for J in range(1, 7):
        X = rand((2**J + 1, 3), seed=J)
        for inter in (False, True):
            op = WaveletTransformOp(J, interleaved=inter)
            tag = 'J%d_%s' % (J, 'int' if inter else 'lvl')
            out['W_' + tag] = op @ X
            out['WT_' + tag] = op.T @ X
            out['levels_' + tag] = np.asarray(op.levels)
    J = 4
    op = WaveletTransformOp(J, interleaved=True)
    for j in range(1, J + 1):
        out['split_J4_j%d' % j] = op.split(j).toarray()
    np.savez_compressed(os.path.join(GOLDEN, 'wavelets.npz'), **out)


def gen_multigrid():
    from source.lanczos import Lanczos
    from source.multigrid import MultiGrid
    out = {}
    for order in ('class', 'lex', 'random'):
        for Js in range(0, 5):
            prob = SquareProblem(Js, 2, order=order, seed=7)
            B = rand((prob.M, 4), seed=10 + Js)
            for nu, vc in ((3, 2), (1, 1)):
                tag = '%s_J%d_nu%d_vc%d' % (order, Js, nu, vc)
                mgA = MultiGrid(prob.A_x, prob.hierarchy, smoothsteps=nu,
                                vcycles=vc)
                out['KinvB_' + tag] = np.stack(
                    [mgA @ B[:, k] for k in range(B.shape[1])], axis=1)
                mgC = MultiGrid(prob.Cinv_j[2], prob.hierarchy,
                                smoothsteps=nu, vcycles=vc)
                out['C2B_' + tag] = np.stack(
                    [mgC @ B[:, k] for k in range(B.shape[1])], axis=1)
            if order == 'class' and Js <= 3:
                mg = MultiGrid(prob.A_x, prob.hierarchy)  # defaults (2, 1)
                np.random.seed(3)
                w = 2.0 * np.random.rand(prob.M) - 1.0
                lz = Lanczos(prob.A_x, mg, w=w.copy())
                out['lanczos_mgA_J%d' % Js] = np.array(
                    [lz.lmax, lz.lmin, lz.iterations])
    np.savez_compressed(os.path.join(GOLDEN, 'multigrid.npz'), **out)


def graph_outputs(prob, mode, P=1, store_solution=True, precond='multigrid'):
    """Runs on every emulated rank; returns rank-local pieces."""
    g = ref_harness.RefGraph(prob, wavelettransform=mode, precond=precond)
    a, b = g.dofs_distr.t_begin, g.dofs_distr.t_end
    X = rand((prob.N, prob.M))[a:b]
    res = {}
    for name in ('W', 'S', 'WT', 'P', 'WT_S_W'):
        res[name] = (getattr(g, name) @ g.vector(X)).X_loc.copy()
    for k, term in enumerate(g.S_terms):
        res['S_term%d' % k] = (term @ g.vector(X)).X_loc.copy()
    res['rhs'] = g.rhs.X_loc.copy()
    hist_rr, hist_ww = [], []
    w, iters = g.solve(callback=lambda w, r, k: (hist_rr.append(r.dot(r)),
                                                hist_ww.append(w.dot(w))))
    res['iters'] = iters
    res['hist_rr'] = np.array(hist_rr)
    res['hist_ww'] = np.array(hist_ww)
    u = g.W @ w
    res['norm_u'] = sqrt_(u.dot(u))
    if store_solution:
        res['w'] = w.X_loc.copy()
    return res


def sqrt_(x):
    return float(np.sqrt(x))


def _graph_rank(args):
    Jt, Js, mode, store = args[:4]
    prob = (CubeProblem if len(args) > 4 and args[4] == 'cube' else
            SquareProblem)(Js, Jt)
    precond = args[5] if len(args) > 5 else 'multigrid'
    with contextlib.redirect_stdout(io.StringIO()):
        return graph_outputs(prob, mode, store_solution=store, precond=precond)


def merge(parts):
    out = {}
    for k in parts[0]:
        if isinstance(parts[0][k], np.ndarray) and parts[0][k].ndim == 2:
            out[k] = np.concatenate([p[k] for p in parts], axis=0)
        else:
            out[k] = parts[0][k]
    return out


def gen_graph():
    from mpi4py import MPI
    out = {}
    cases = [(2, 2, 'composite', 1, True), (2, 2, 'original', 1, True),
             (2, 2, 'interleaved', 1, True), (3, 3, 'composite', 1, True),
             (3, 3, 'composite', 2, True), (3, 3, 'original', 3, True),
             (4, 2, 'composite', 4, True), (3, 6, 'composite', 1, False),
             (3, 6, 'composite', 2, False)]
    for Jt, Js, mode, P, store in cases:
        args = (Jt, Js, mode, store)
        if P == 1:
            res = _graph_rank(args)
        else:
            res = merge(MPI.launch(P, _graph_rank, args))
        tag = 'Jt%d_Js%d_%s_P%d' % (Jt, Js, mode, P)
        print(tag, 'iters', res['iters'], 'norm_u', res['norm_u'])
        for k, v in res.items():
            if Js >= 6 and isinstance(v, np.ndarray) and v.ndim == 2:
                # config 1/2: keep norms of the big fields only
                out['%s__norm_%s' % (tag, k)] = np.linalg.norm(v)
            else:
                out['%s__%s' % (tag, k)] = v
    np.savez_compressed(os.path.join(GOLDEN, 'graph.npz'), **out)


def gen_cube():
    """problem='cube' (problem.py:21-32, heateq_mpi_test.py:209-243): the
    reference classes on the Kuhn-triangulation matrices of assembly.py."""
    from mpi4py import MPI
    from source.multigrid import MultiGrid
    out = {}
    for Jt, Js, mode, P in ((2, 1, 'original', 1), (2, 2, 'composite', 1),
                            (3, 2, 'composite', 2)):
        args = (Jt, Js, mode, True, 'cube')
        res = _graph_rank(args) if P == 1 else merge(
            MPI.launch(P, _graph_rank, args))
        tag = 'cube_Jt%d_Js%d_%s_P%d' % (Jt, Js, mode, P)
        print(tag, 'iters', res['iters'], 'norm_u', res['norm_u'])
        for k, v in res.items():
            out['%s__%s' % (tag, k)] = v
    prob = CubeProblem(2, 1)
    B = rand((prob.M, 3), seed=31)
    mg = MultiGrid(prob.Cinv_j[1], prob.hierarchy, smoothsteps=3, vcycles=2)
    out['cube_C1B_J2'] = np.stack([mg @ B[:, k] for k in range(3)], axis=1)
    np.savez_compressed(os.path.join(GOLDEN, 'cube.npz'), **out)


def gen_direct():
    """precond='direct' (linop.py:18-26, heateq_mpi.py:154-157): the cases of
    heateq_mpi_test.py:66-135 (J_time=4, J_space=2, 'original' and the default
    wavelet transform) plus a two-rank one, from the reference classes."""
    from mpi4py import MPI
    from source.linop import InvLinOp
    out = {}
    for Jt, Js, mode, P in ((4, 2, 'original', 1), (4, 2, 'composite', 1),
                            (3, 3, 'composite', 2)):
        args = (Jt, Js, mode, True, 'square', 'direct')
        res = _graph_rank(args) if P == 1 else merge(
            MPI.launch(P, _graph_rank, args))
        tag = 'direct_Jt%d_Js%d_%s_P%d' % (Jt, Js, mode, P)
        print(tag, 'iters', res['iters'], 'norm_u', res['norm_u'])
        for k, v in res.items():
            out['%s__%s' % (tag, k)] = v
    prob = SquareProblem(3, 2)
    B = rand((prob.M, 3), seed=41)
    for name, mat in (('A', prob.A_x), ('C1', prob.Cinv_j[1])):
        inv = InvLinOp(mat)
        out['inv_%s_J3' % name] = np.stack([inv @ B[:, k] for k in range(3)],
                                           axis=1)
    np.savez_compressed(os.path.join(GOLDEN, 'direct.npz'), **out)


def gen_lanczos():
    from source.lanczos import Lanczos
    prob = SquareProblem(2, 3)
    with contextlib.redirect_stdout(io.StringIO()):
        g = ref_harness.RefGraph(prob)
    w = g.vector(rand((prob.N, prob.M), seed=5))
    lz = Lanczos(g.WT_S_W, g.P, w=w)
    np.savez_compressed(os.path.join(GOLDEN, 'lanczos.npz'),
                        Jt3_Js2=np.array([lz.lmax, lz.lmin, lz.iterations]))
    print('lanczos kappa', lz.cond(), 'its', lz.iterations)


if __name__ == '__main__':
    os.makedirs(GOLDEN, exist_ok=True)
    if 'cube' in sys.argv[1:]:  # python -m oracle.gen_golden cube
        gen_cube()
        sys.exit(0)
    if 'direct' in sys.argv[1:]:  # python -m oracle.gen_golden direct
        gen_direct()
        sys.exit(0)
    gen_wavelets()
    gen_multigrid()
    gen_lanczos()
    gen_graph()
    gen_cube()
    gen_direct()
