"""Device-resident read-only matrices.

The reference publishes every assembled matrix to the ranks of a node through
MPI-3 shared-memory windows (/root/reference/source/mpi_shared_mem.py:6-50).
On a B200 box each rank owns one GPU, so the equivalent of "shared, read-only,
zero-copy" is a single upload into that GPU's HBM, done once:
`shared_sparse_matrix(mat, comm)` keeps its call shape and returns a handle
accepted wherever the operators take a `mat_space`.
"""
import numpy as np
import torch

from .linop import DeviceCSR


def shared_sparse_matrix(mat, shared_comm=None):
    """CSR (fp64 data, int32 indices/indptr, mpi_shared_mem.py:46-48) in HBM.
    Rank 0 of `shared_comm` supplies `mat`; other ranks may pass None and
    receive it by broadcast (mpi_shared_mem.py:31-50)."""
    if shared_comm is not None and shared_comm.Get_size() > 1:
        mat = shared_comm.bcast(mat if shared_comm.Get_rank() == 0 else None)
    return DeviceCSR(mat)


def shared_numpy_array(arr, shared_comm=None):
    """Dense read-only array in HBM (mpi_shared_mem.py:6-28)."""
    from .mpi_vector import _device
    if shared_comm is not None and shared_comm.Get_size() > 1:
        arr = shared_comm.bcast(arr if shared_comm.Get_rank() == 0 else None)
    return torch.from_numpy(np.ascontiguousarray(arr)).to(_device())
