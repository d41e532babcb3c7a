"""CPU: the oracle (oracle/restate.py) against the known answers of the
reference's own tests and against the golden outputs of the unmodified
reference classes (tests/golden, made by oracle/gen_golden.py)."""
import numpy as np
import scipy.sparse as sp

from conftest import rand, rel
from oracle import restate as orc
from spacetime_fullgrid_parallel_b200.assembly import CubeProblem, SquareProblem


def test_slab_bounds():
    """mpi_vector.py:18-31: leftovers go to the last ranks."""
    assert orc.slab_bounds(9, 2) == [(0, 4), (4, 9)]
    assert orc.slab_bounds(257, 8) == [(32 * p, 32 * p + 32) for p in range(7)
                                      ] + [(224, 257)]
    assert orc.slab_bounds(5, 5) == [(k, k + 1) for k in range(5)]


def test_wavelet_known_answers():
    """wavelets_test.py:22-72."""
    s2 = np.sqrt(2)
    W2 = orc.wavelet_synthesis(np.eye(5), 2, interleaved=True)
    assert np.allclose(W2, [[1, -2, -s2, 0, 0], [3 / 4, 2, 0, 0, 1 / 4],
                            [1 / 2, -1, s2, -1, 1 / 2],
                            [1 / 4, 0, 0, 2, 3 / 4], [0, 0, -s2, -2, 1]])
    for J in range(1, 8):
        N = 2**J + 1
        W = orc.wavelet_synthesis(np.eye(N), J, interleaved=False)
        assert np.allclose(W[:, 0], np.linspace(1, 0, N))
        assert np.allclose(W[:, 1], np.linspace(0, 1, N))
        h = 2**(J - 1)
        assert np.allclose(W[:h + 1, 2], np.linspace(-s2, s2, h + 1))
        assert np.allclose(W[h:, 2], np.linspace(s2, -s2, h + 1))
        if J >= 2:
            assert np.isclose(W[0, 3], -2) and np.isclose(W[-1, 4], -2)
        Wi = orc.wavelet_synthesis(np.eye(N), J, interleaved=True)
        prod = np.eye(N)
        for j in range(1, J + 1):
            prod = prod + orc.wavelet_split(J, j) @ prod
        assert np.allclose(prod, Wi)
        WT = orc.wavelet_analysis(np.eye(N), J, interleaved=True)
        assert np.allclose(WT, Wi.T)


def test_wavelets_golden(golden):
    g = golden['wavelets']
    for J in range(1, 7):
        X = rand((2**J + 1, 3), seed=J)
        for inter, tag in ((True, 'int'), (False, 'lvl')):
            assert np.array_equal(orc.wavelet_levels(J, inter),
                                  g['levels_J%d_%s' % (J, tag)])
            assert rel(orc.wavelet_synthesis(X, J, inter),
                       g['W_J%d_%s' % (J, tag)]) < 1e-14
            assert rel(orc.wavelet_analysis(X, J, inter),
                       g['WT_J%d_%s' % (J, tag)]) < 1e-14
    for j in range(1, 5):
        assert rel(orc.wavelet_split(4, j).toarray(),
                   g['split_J4_j%d' % j]) < 1e-15


def test_kron_fixture():
    """mpi_kron_test.py:57-78: the 5x5 tridiagonal fixture vs np.kron."""
    mat = np.array([[3.5, 13., 28.5, 50., 77.5],
                    [-5., -23., -53., -95., -149.],
                    [2.5, 11., 25.5, 46., 72.5]])
    T = sp.spdiags(mat, (1, 0, -1), 5, 5).T.copy().tocsr()
    A = sp.csr_matrix(np.arange(1, 10).reshape(3, 3).astype(float))
    X = rand((5, 3), seed=0)
    ref = (np.kron(T.toarray(), A.toarray()) @ X.reshape(-1)).reshape(5, 3)
    assert rel(orc.kron_apply(T, A, X), ref) < 1e-14
    assert np.array_equal(orc.permute(X), X.reshape(5, 3).T)  # mpi_vector_test.py:31-48


def test_multigrid_golden(golden):
    g = golden['multigrid']
    for order in ('class', 'lex', 'random'):
        for Js in range(0, 5):
            prob = SquareProblem(Js, 2, order=order, seed=7)
            B = rand((prob.M, 4), seed=10 + Js)
            for nu, vc in ((3, 2), (1, 1)):
                tag = '%s_J%d_nu%d_vc%d' % (order, Js, nu, vc)
                mgA = orc.MultiGridOracle(prob.A_x, prob.hierarchy.P_mats, nu,
                                          vc)
                assert rel(mgA @ B, g['KinvB_' + tag]) < 1e-13
                mgC = orc.MultiGridOracle(prob.Cinv_j[2],
                                          prob.hierarchy.P_mats, nu, vc,
                                          threads=2)
                assert rel(mgC @ B, g['C2B_' + tag]) < 1e-13


def test_galerkin_and_symmetry():
    """multigrid_test.py:14-37,87-98 restated on the synthetic hierarchy: the
    Galerkin product equals the coarse-mesh assembly; MG is symmetric."""
    prob = SquareProblem(3, 1)
    h = prob.hierarchy
    A = prob.A_x
    for j in reversed(range(h.J)):
        A = (h.R_mats[j] @ A @ h.P_mats[j]).tocsr()
        assert rel(A.toarray(), h.assemble('stiff', j).toarray()) < 1e-13
    mg = orc.MultiGridOracle(prob.A_x, h.P_mats, 2, 1)
    Pm = mg @ np.eye(prob.M)
    assert rel(Pm, Pm.T) < 1e-13


def test_graph_golden(golden):
    g = golden['graph']
    for Jt, Js, inter, tag in ((2, 2, True, 'Jt2_Js2_composite_P1'),
                               (2, 2, False, 'Jt2_Js2_original_P1'),
                               (3, 3, True, 'Jt3_Js3_composite_P2'),
                               (4, 2, True, 'Jt4_Js2_composite_P4')):
        prob = SquareProblem(Js, Jt)
        o = orc.HeatEqOracle(prob, interleaved=inter)
        X = rand((prob.N, prob.M))
        for name in ('W', 'S', 'WT', 'P', 'WT_S_W'):
            assert rel(getattr(o, name)(X), g['%s__%s' % (tag, name)]) < 1e-13
        for k, (T, op) in enumerate(o.terms):
            assert rel(orc.kron_apply(T, op, X),
                       g['%s__S_term%d' % (tag, k)]) < 1e-13
        assert rel(o.rhs, g[tag + '__rhs']) < 1e-15
        rr, ww = [], []
        w, iters = o.solve(callback=lambda w, r, k: (rr.append(orc.dot(r, r)),
                                                     ww.append(orc.dot(w, w))))
        assert iters == int(g[tag + '__iters'])
        assert np.allclose(np.sqrt(rr), np.sqrt(g[tag + '__hist_rr']),
                           rtol=0, atol=1e-10 * np.sqrt(rr[0]))
        assert rel(w, g[tag + '__w']) < 1e-10
        u = o.W(w)
        assert abs(np.linalg.norm(u) - float(g[tag + '__norm_u'])) < 1e-10


def test_direct_golden(golden):
    """precond='direct' (linop.py:18-26, heateq_mpi.py:154-157): the oracle
    against outputs of the reference classes -- the cases of
    heateq_mpi_test.py:66-135 (J_time=4, J_space=2) and a two-rank one."""
    g = golden['direct']
    prob = SquareProblem(3, 2)
    B = rand((prob.M, 3), seed=41)
    for name, mat in (('A', prob.A_x), ('C1', prob.Cinv_j[1])):
        assert rel(orc.InvOracle(mat) @ B, g['inv_%s_J3' % name]) < 1e-13
    for Jt, Js, inter, tag in ((4, 2, False, 'direct_Jt4_Js2_original_P1'),
                               (4, 2, True, 'direct_Jt4_Js2_composite_P1'),
                               (3, 3, True, 'direct_Jt3_Js3_composite_P2')):
        prob = SquareProblem(Js, Jt)
        o = orc.HeatEqOracle(prob, interleaved=inter, precond='direct')
        X = rand((prob.N, prob.M))
        for name in ('W', 'S', 'WT', 'P', 'WT_S_W'):
            assert rel(getattr(o, name)(X), g['%s__%s' % (tag, name)]) < 1e-12
        w, iters = o.solve()
        assert iters == int(g[tag + '__iters'])
        assert rel(w, g[tag + '__w']) < 1e-10


def test_config1_norms(golden):
    """BASELINE config 1/2 (J_time=3, J_space=6): norms of the big fields."""
    g = golden['graph']
    tag = 'Jt3_Js6_composite_P1'
    prob = SquareProblem(6, 3)
    o = orc.HeatEqOracle(prob, threads=4)
    X = rand((prob.N, prob.M))
    for name in ('W', 'S', 'P'):
        ref = float(g['%s__norm_%s' % (tag, name)])
        assert abs(np.linalg.norm(getattr(o, name)(X)) - ref) < 1e-12 * ref
    assert int(g[tag + '__iters']) == int(g['Jt3_Js6_composite_P2__iters'])


def test_lanczos_golden(golden):
    g = golden['multigrid']
    for Js in range(0, 4):
        prob = SquareProblem(Js, 2, order='class', seed=7)
        mg = orc.MultiGridOracle(prob.A_x, prob.hierarchy.P_mats, 2, 1)
        np.random.seed(3)
        w = 2.0 * np.random.rand(prob.M) - 1.0
        with np.errstate(all='ignore'):
            lmax, lmin, its = orc.lanczos(
                lambda x: prob.A_x @ x, lambda x: mg @ x, w)
        rmax, rmin, rits = g['lanczos_mgA_J%d' % Js]
        assert abs(lmax - rmax) < 1e-10 * rmax and abs(lmin - rmin) < 1e-10 * rmin
        assert its == int(rits)


def test_cube_golden(golden):
    """problem='cube' (problem.py:21-32): oracle vs the reference classes on
    the Kuhn-triangulation matrices (tests/golden/cube.npz)."""
    g = golden['cube']
    for Jt, Js, inter, tag in ((2, 1, False, 'cube_Jt2_Js1_original_P1'),
                               (2, 2, True, 'cube_Jt2_Js2_composite_P1'),
                               (3, 2, True, 'cube_Jt3_Js2_composite_P2')):
        prob = CubeProblem(Js, Jt)
        o = orc.HeatEqOracle(prob, interleaved=inter)
        X = rand((prob.N, prob.M))
        for name in ('W', 'S', 'WT', 'P', 'WT_S_W'):
            assert rel(getattr(o, name)(X), g['%s__%s' % (tag, name)]) < 1e-13
        w, iters = o.solve()
        assert iters == int(g[tag + '__iters'])
        assert rel(w, g[tag + '__w']) < 1e-10
    prob = CubeProblem(2, 1)
    B = rand((prob.M, 3), seed=31)
    mg = orc.MultiGridOracle(prob.Cinv_j[1], prob.hierarchy.P_mats, 3, 2)
    assert rel(mg @ B, g['cube_C1B_J2']) < 1e-13


def test_cube_assembler():
    """Galerkin identity on the Kuhn hierarchy (multigrid_test.py:14-37 covers
    the cube), stencil sizes, and a Poisson solve against the exact solution."""
    import scipy.sparse.linalg as sla
    prob = CubeProblem(2, 1)
    h = prob.hierarchy
    assert prob.M == 7**3
    assert prob.M_x.getnnz(axis=1).max() == 15
    assert prob.A_x.getnnz(axis=1).max() == 7
    for kind in ('stiff', 'mass'):
        A = h.assemble(kind)
        for j in reversed(range(h.J)):
            A = (h.R_mats[j] @ A @ h.P_mats[j]).tocsr()
            ref = h.assemble(kind, j)
            assert abs(A - ref).max() < 1e-13 * abs(ref).max()
    prob = CubeProblem(3, 1)
    u_ex = prob._u0()
    u = sla.spsolve(prob.A_x.tocsc(), prob.M_x @ (3 * np.pi**2 * u_ex))
    assert rel(u, u_ex) < 0.03
