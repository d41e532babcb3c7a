"""Time <-> space re-sharding (the reference's `KronVectorMPI.permute`,
mpi_vector.py:212-240) as one all-to-all over NVLink.

In the time-fastest block layout a rank's block is (M, ld) with the space dof
as the slow index, so the part of it that rank q needs after re-sharding by
space -- rows [x_q, x_q') -- is one CONTIGUOUS range: the send side needs no
packing and no transpose.  The receive side places the (M_loc x n_p) piece
from rank p at columns [t_p, t_p') of its (M_loc, ld(N)) "space-sharded block",
in which the whole time axis of each owned space dof is contiguous: exactly
what stk_time_apply wants.  `PermutePlan` is host logic (CPU-testable).
"""
import torch

from ._lib import check, lib, ptr, stream
from .mpi_vector import DofDistributionMPI, KronVectorMPI, pitch


def _device_copy_cols(src, c_in, n, dst, c_out):
    """dst[:, c_out:c_out+n] = src[:, c_in:c_in+n] on device blocks (libstk)."""
    check(lib().stk_copy_cols(src.shape[0], n, ptr(src), src.shape[1], c_in,
                              None, ptr(dst), dst.shape[1], c_out, stream()))


class PermutePlan:
    """`copy_cols(src, c_in, n, dst, c_out)` places pieces; the default is the
    device kernel (the CPU tests of the plan's host logic pass a stand-in)."""
    def __init__(self, dofs_distr, copy_cols=_device_copy_cols):
        self.copy_cols = copy_cols
        d = self.dofs_distr = dofs_distr
        self.space_distr = DofDistributionMPI(d.comm, d.M, d.N)
        self.t_bounds = d.dof_distribution
        self.x_bounds = self.space_distr.dof_distribution
        self.m_loc = self.space_distr.t_end - self.space_distr.t_begin
        self.ld_full = pitch(d.N)

    def forward(self, block, n_loc, ld):
        """(M, ld) time-sharded block -> (m_loc, ld_full) space-sharded."""
        d = self.dofs_distr
        P = d.size
        out = torch.zeros((self.m_loc, self.ld_full), dtype=block.dtype,
                          device=block.device)
        sends = [block[xa:xb].reshape(-1) for xa, xb in self.x_bounds]
        recvs = [
            torch.empty(self.m_loc * pitch(tb - ta), dtype=block.dtype,
                        device=block.device) for ta, tb in self.t_bounds
        ]
        d.comm.all_to_all(sends, recvs)
        for p, (ta, tb) in enumerate(self.t_bounds):  # place the pieces
            self.copy_cols(recvs[p].view(self.m_loc, pitch(tb - ta)), 0,
                           tb - ta, out, ta)
        return out

    def backward(self, sblock, n_loc, ld):
        """(m_loc, ld_full) space-sharded block -> (M, ld) time-sharded."""
        d = self.dofs_distr
        ta, tb = d.t_begin, d.t_end
        out = torch.zeros((d.M, ld), dtype=sblock.dtype, device=sblock.device)
        sends = []
        for pa, pb in self.t_bounds:
            piece = torch.zeros((self.m_loc, pitch(pb - pa)),
                                dtype=sblock.dtype, device=sblock.device)
            self.copy_cols(sblock, pa, pb - pa, piece, 0)
            sends.append(piece.reshape(-1))
        recvs = [out[xa:xb].reshape(-1) for xa, xb in self.x_bounds]
        # out rows are contiguous, so the pieces land in place
        d.comm.all_to_all(sends, recvs)
        return out


_plans = {}


def plan_for(dofs_distr):
    key = id(dofs_distr)
    if key not in _plans:
        _plans[key] = (dofs_distr, PermutePlan(dofs_distr))
    return _plans[key][1]


def permute_vector(vec):
    """Public `permute`: a KronVectorMPI over the swapped distribution
    (space index sharded, time index local), in the standard block layout of
    that distribution; involves one on-device transposition of the
    space-sharded block (API parity path, used by tests)."""
    plan = plan_for(vec.dofs_distr)
    sblock = plan.forward(vec.data, vec.n_loc, vec.ld)  # (m_loc, ld(N))
    sd = plan.space_distr
    out = KronVectorMPI(sd)
    # standard layout of the swapped vector: [time index N][local space dofs]
    out.data[:, :plan.m_loc].copy_(sblock[:, :vec.N].t())
    return out
