"""TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the absent `mpi4py` package.

It exists so that the *unmodified* reference modules under /root/reference can
be imported in the build container (which has no MPI) to validate the oracle
and to generate the golden fixtures under tests/golden/.  It is never imported
by the product package.  Surface = SURVEY.md Appendix A.

Two communicators are provided:
  * a one-rank in-process world (default), and
  * a multi-process world built on multiprocessing pipes (see `launch`),
    which gives "mpirun -np P" semantics for the reference's Isend/Irecv/Recv,
    allreduce, bcast, gather, Scatterv and Gatherv call sites.
"""
from . import MPI  # noqa: F401
