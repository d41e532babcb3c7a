"""GPU parity AT THE BASELINE SIZES against the CPU oracle (oracle/restate.py,
pinned to the unmodified reference classes by tests/test_oracle.py): index and
overflow bugs live at scale, where the golden fixtures cannot go.

 * J_time = 8, J_space = 9 (BASELINE configs[3], 2.7e8 dofs): a handful of time
   slices of P x and S x.  Both are slice-local after the time stencil
   (/root/reference/source/mpi_kron.py:122-132 and :143-150, :186-201), so the
   oracle needs only those slices: one multigrid solve per slice and operator.
 * J_time = 8, 9, 10 (the time axes of configs[2..4]): the full W x and W^T x
   on a few hundred space dofs (wavelets.py:172-198).

Tolerance: 1e-12 relative in the 2-norm (north_star).
"""
import numpy as np
import pytest

from conftest import rel

pytestmark = pytest.mark.gpu
TOL = 1e-12


def test_P_and_S_slices_at_config4(cuda):
    import torch
    from oracle import restate
    from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
    from spacetime_fullgrid_parallel_b200.heateq_mpi import HeatEquationMPI
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
    Jt, Js = 8, 9
    prob = SquareProblem(Js, Jt)
    heq = HeatEquationMPI(J_space=Js, J_time=Jt, problem=prob)
    N, M = heq.N, heq.M
    x = KronVectorMPI(heq.dofs_distr)
    gen = torch.Generator(device='cuda').manual_seed(20)
    x.data[:, :N] = torch.rand((M, N), dtype=torch.float64, device='cuda',
                               generator=gen)
    P_mats = prob.hierarchy.P_mats

    def slices(vec, ts):
        """Time slices ts of a device vector as host rows (len(ts), M)."""
        return vec.data[:, ts].t().contiguous().cpu().numpy()

    # ---- P: slice t gets C_j A_x C_j, j = wavelet level of t ----
    levels = np.asarray(heq.W.levels)
    ts = [0, 1, 2, 130, N - 1]
    Px = slices(heq.P @ x, ts)
    xs = slices(x, ts)
    for k, t in enumerate(ts):
        j = int(levels[t])
        C = restate.MultiGridOracle(prob.Cinv_j[j], P_mats, 3, 2)
        ref = C(restate.apply_space(prob.A_x, C(xs[k:k + 1])))
        assert rel(Px[k:k + 1], ref) < TOL, ('P', t, j)
    del Px

    # ---- S: five Kronecker terms; the time stencils need slices t-1, t, t+1
    ts = [0, 77, N - 1]  # t = 0 carries the G_t term (heateq_mpi.py:178)
    Sx = slices(heq.S @ x, ts)
    need = sorted({s for t in ts for s in (t - 1, t, t + 1) if 0 <= s < N})
    xn = dict(zip(need, slices(x, need)))
    K = restate.MultiGridOracle(prob.A_x, P_mats, 3, 2)
    Mx, Ax = prob.M_x, prob.A_x
    terms = [(prob.A_t, restate.chain(Mx, K, Mx)),
             (prob.L_t, restate.chain(Mx, K, Ax)),
             (prob.L_t.T.tocsr(), restate.chain(Ax, K, Mx)),
             (prob.M_t, restate.chain(Ax, K, Ax)),
             (prob.G_t, restate.chain(Mx))]
    for k, t in enumerate(ts):
        ref = np.zeros((1, M))
        for T, op in terms:
            row = T.getrow(t)
            if row.nnz == 0:
                continue
            tx = sum(v * xn[int(c)] for c, v in zip(row.indices, row.data))
            ref += restate.apply_space(op, np.ascontiguousarray(tx[None, :]))
        assert rel(Sx[k:k + 1], ref) < TOL, ('S', t)


@pytest.mark.parametrize('Jt', [8, 9, 10])
def test_wavelet_transforms_long_time_axes(cuda, Jt):
    import torch
    from oracle import restate
    from spacetime_fullgrid_parallel_b200.comm import world
    from spacetime_fullgrid_parallel_b200.mpi_vector import (
        DofDistributionMPI, KronVectorMPI)
    from spacetime_fullgrid_parallel_b200.wavelets import (
        TransposedWaveletTransformKronIdentityMPI,
        WaveletTransformKronIdentityMPI)
    N, M = 2**Jt + 1, 300
    d = DofDistributionMPI(world(), N, M)
    X = np.random.RandomState(Jt).rand(N, M)
    x = KronVectorMPI(d, X)
    W = WaveletTransformKronIdentityMPI(d, Jt)
    WT = TransposedWaveletTransformKronIdentityMPI(d, Jt)
    assert rel((W @ x).X_loc, restate.wavelet_synthesis(X, Jt)) < TOL
    assert rel((WT @ x).X_loc, restate.wavelet_analysis(X, Jt)) < TOL
