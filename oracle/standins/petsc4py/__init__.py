"""TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the absent `petsc4py`.

Provides the handful of PETSc calls the reference smoother makes
(/root/reference/source/multigrid.py:100-127): Mat().createAIJWithArrays,
Vec().createWithArray / setArray and Mat.SOR with FORWARD/BACKWARD sweeps.
The sweep itself is the oracle's C Gauss-Seidel (oracle/gs.c), i.e. the
lexicographic update  x_i <- x_i + (b_i - A_i.x)/a_ii  that the reference's
own pure-Python `Smoother` (multigrid.py:83-97) defines.
Never imported by the product package.
"""
from . import PETSc  # noqa: F401
