"""ORACLE / TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference classes from /root/reference/source (imported
through the mpi4py / petsc4py stand-ins in oracle/standins) on matrices from
the product's host assembler.  Only usable in the build container (the GPU box
has no /root/reference); it exists to pin oracle/restate.py and to generate the
golden fixtures in tests/golden/ (see oracle/gen_golden.py).

The reference's NGSolve block (heateq_mpi.py:63-104) cannot run here, so the
operator graph of heateq_mpi.py:126-191 is wired by hand from the same
reference classes, fed with `SquareProblem` matrices.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _root():
    """The reference tree: $STK_REFERENCE_ROOT, /root/reference (the build
    container), else the verbatim copy oracle/fetch_ref.py left in oracle/_ref
    (what the GPU box has)."""
    for cand in (os.environ.get('STK_REFERENCE_ROOT'), '/root/reference',
                 os.path.join(_HERE, '_ref')):
        if cand and os.path.isdir(os.path.join(cand, 'source')):
            return cand
    return '/root/reference'


REFERENCE_ROOT = _root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'source'))


def activate():
    """Make `import source.*` resolve to the reference and mpi4py/petsc4py to
    the stand-ins.  Idempotent."""
    assert available(), 'reference tree not present'
    for p in (os.path.join(_HERE, 'standins'), REFERENCE_ROOT,
              os.path.dirname(_HERE)):
        if p not in sys.path:
            sys.path.insert(0, p)


class RefGraph:
    """The objects HeatEquationMPI.__init__ builds after assembly, made of
    reference classes only."""
    def __init__(self, prob, wavelettransform='composite', smoothsteps=3,
                 vcycles=2, comm=None, precond='multigrid'):
        activate()
        import numpy as np
        from mpi4py import MPI
        from source.linop import CompositeLinOp, InvLinOp
        from source.mpi_kron import (BlockDiagMPI, CompositeMPI,
                                     MatKronIdentityMPI, SumMPI,
                                     TridiagKronMatMPI)
        from source.mpi_vector import DofDistributionMPI, KronVectorMPI
        from source.multigrid import MultiGrid
        from source.wavelets import (
            TransposedWaveletTransformKronIdentityMPI,
            WaveletTransformKronIdentityMPI, WaveletTransformOp)

        comm = MPI.COMM_WORLD if comm is None else comm
        p = self.prob = prob
        d = self.dofs_distr = DofDistributionMPI(comm, p.N, p.M)
        if wavelettransform == 'composite':  # heateq_mpi.py:127-131
            self.W = WaveletTransformKronIdentityMPI(d, p.J_time)
            self.WT = TransposedWaveletTransformKronIdentityMPI(d, p.J_time)
        else:  # heateq_mpi.py:132-139
            self.W_t = WaveletTransformOp(
                p.J_time, interleaved=(wavelettransform == 'interleaved'))
            self.W = MatKronIdentityMPI(d, self.W_t)
            self.WT = MatKronIdentityMPI(d, self.W_t.T)
        h = p.hierarchy
        mg = dict(smoothsteps=smoothsteps, vcycles=vcycles)
        if precond == 'multigrid':
            self.Kinv_x = MultiGrid(p.A_x, h, **mg)  # heateq_mpi.py:144-147
            self.C_j = [MultiGrid(mat, h, **mg) for mat in p.Cinv_j]
        else:  # heateq_mpi.py:154-157
            assert precond == 'direct'
            self.Kinv_x = InvLinOp(p.A_x)
            self.C_j = [InvLinOp(mat) for mat in p.Cinv_j]
        self.CAC_j = [
            CompositeLinOp([C, p.A_x, C]) for C in self.C_j
        ]  # heateq_mpi.py:159-162
        K, Mx, Ax = self.Kinv_x, p.M_x, p.A_x
        self.S_terms = [  # heateq_mpi.py:166-178
            TridiagKronMatMPI(d, p.A_t, CompositeLinOp([Mx, K, Mx])),
            TridiagKronMatMPI(d, p.L_t, CompositeLinOp([Mx, K, Ax])),
            TridiagKronMatMPI(d, p.L_t.T.tocsr(), CompositeLinOp([Ax, K,
                                                                  Mx])),
            TridiagKronMatMPI(d, p.M_t, CompositeLinOp([Ax, K, Ax])),
            TridiagKronMatMPI(d, p.G_t, Mx),
        ]
        self.S = SumMPI(d, self.S_terms)
        self.P = BlockDiagMPI(d, [self.CAC_j[j] for j in self.W.levels])
        self.WT_S_W = CompositeMPI(d, [self.WT, self.S, self.W])
        self.rhs = KronVectorMPI(d)  # heateq_mpi.py:189-191
        self.rhs.X_loc[:] = np.kron(p.u0_t[self.rhs.t_begin:self.rhs.t_end],
                                    p.u0_x).reshape(-1, p.M)
        self.KronVectorMPI = KronVectorMPI

    def vector(self, X_loc=None):
        v = self.KronVectorMPI(self.dofs_distr)
        if X_loc is not None:
            v.X_loc[:] = X_loc
        return v

    def solve(self, **kw):
        from source.linalg import PCG
        return PCG(self.WT_S_W, self.P, self.rhs, **kw)
