"""Elementwise parity of P x and K x against the CPU oracle as the space mesh
grows (diagnosis of rounding growth):  python tools/diag_parity_scale.py 4 5 6 7 8"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    import torch
    from oracle import restate
    from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
    from spacetime_fullgrid_parallel_b200.heateq_mpi import HeatEquationMPI
    from spacetime_fullgrid_parallel_b200.mpi_vector import KronVectorMPI
    Jt = 2
    for Js in [int(a) for a in sys.argv[1:]]:
        prob = SquareProblem(Js, Jt)
        heq = HeatEquationMPI(J_space=Js, J_time=Jt, problem=prob)
        N, M = heq.N, heq.M
        X = np.random.RandomState(3).rand(N, M)
        x = KronVectorMPI(heq.dofs_distr, X)
        Px = np.asarray((heq.P @ x).X_loc)
        levels = np.asarray(heq.W.levels)
        P_mats = prob.hierarchy.P_mats
        out = []
        for t in range(N):
            j = int(levels[t])
            C = restate.MultiGridOracle(prob.Cinv_j[j], P_mats, 3, 2)
            cx = C(X[t:t + 1])
            ref = C(restate.apply_space(prob.A_x, cx))
            out.append('t%d(j%d) %.1e' % (t, j, np.linalg.norm(Px[t] - ref[0]) / np.linalg.norm(ref)))
        # K_x = MG(A_x) on one block (the uniform K = 1 path)
        y = x.empty_like()
        heq.Kinv_x.apply_block(x.data, y.data)
        K = restate.MultiGridOracle(prob.A_x, P_mats, 3, 2)
        refK = K(X[:1])
        eK = np.linalg.norm(np.asarray(y.X_loc)[0] - refK[0]) / np.linalg.norm(refK)
        fused = [lv['fused'] is not None for lv in heq.family._levels]
        print('Js %d M %d fused levels %s | P: %s | K: %.1e' % (
            Js, M, ''.join('F' if f else '-' for f in fused), ' '.join(out), eK), flush=True)
        del heq, x, y
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
