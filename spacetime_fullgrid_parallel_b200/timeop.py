"""Sparse operators along the (sharded) time axis: (T (x) I) x and (T^T (x) I) x.

One mechanism serves the reference's three flavours of time coupling
(SURVEY.md 2b): the +-1-slice halo of `TridiagKronIdentityMPI`
(mpi_kron.py:153-201, mpi_vector.py:140-187), the per-level sparse row
exchange of `SparseKronIdentityMPI` (mpi_kron.py:259-317,
mpi_vector.py:189-203) and, with all wavelet levels multiplied together, the
whole wavelet transform in ONE exchange of <= 2J-1 boundary slices.

`TimeOpPlan` is host logic (numpy only, testable under gloo on CPU): which
global time slices this rank needs from / owes to which peer, and the local
CSR with its columns renumbered [local | halo].  `fetch`, `apply` and
`apply_adjoint` run it on the device through libstk.
"""
import numpy as np
import scipy.sparse as sp


class TimeOpPlan:
    def __init__(self, dofs_distr, T):
        T = sp.csr_matrix(T, dtype=np.float64)
        d = self.dofs_distr = dofs_distr
        N = d.N
        assert T.shape == (N, N)
        a, b = d.t_begin, d.t_end
        self.n_loc = n = b - a
        bounds = d.dof_distribution
        loc = T[a:b].tocsr()
        loc.sort_indices()
        cols = np.unique(loc.indices)
        remote = cols[(cols < a) | (cols >= b)]
        self.halo_cols = remote  # sorted, hence grouped by owner rank
        self.n_halo = len(remote)
        # recv_from[p] = (offset, count) into the halo list
        self.recv_from = {}
        for p, (pa, pb) in enumerate(bounds):
            if p == d.rank:
                continue
            sel = np.nonzero((remote >= pa) & (remote < pb))[0]
            if len(sel):
                self.recv_from[p] = (int(sel[0]), len(sel))
        # send_to[p] = local indices of my slices that rank p's rows touch
        self.send_to = {}
        for p, (pa, pb) in enumerate(bounds):
            if p == d.rank:
                continue
            theirs = np.unique(T[pa:pb].indices)
            mine = theirs[(theirs >= a) & (theirs < b)]
            if len(mine):
                self.send_to[p] = (mine - a).astype(np.int32)
        # local CSR, columns renumbered: local t -> t-a, halo k -> n + k
        remap = np.full(N, -1, dtype=np.int64)
        remap[a:b] = np.arange(n)
        remap[remote] = n + np.arange(len(remote))
        self.local = sp.csr_matrix(
            (loc.data, remap[loc.indices].astype(np.int32), loc.indptr),
            shape=(n, n + len(remote)))
        # adjoint: (T_loc)^T split into the rows that stay here and the rows
        # that belong to the halo owners
        lt = self.local.T.tocsr()
        lt.sort_indices()
        self.adj_local = lt[:n].tocsr()
        self.adj_halo = lt[n:].tocsr()
        self._dev = None
        # plans that move the same slices share one exchange per vector state
        self._key = ('halo', self.halo_cols.tobytes(),
                     tuple(sorted((p, v.tobytes())
                                  for p, v in self.send_to.items())))

    # -- device execution --------------------------------------------------
    def _device_arrays(self, device):
        import torch
        if self._dev is None or self._dev['device'] != device:

            def csr(m):
                return (torch.from_numpy(m.indptr.astype(np.int32)).to(device),
                        torch.from_numpy(m.indices.astype(np.int32)).to(device),
                        torch.from_numpy(m.data.astype(np.float64)).to(device))

            self._dev = {
                'device': device,
                'local': csr(self.local),
                'adj_local': csr(self.adj_local),
                'adj_halo': csr(self.adj_halo),
                'send_idx': {
                    p: torch.from_numpy(idx).to(device)
                    for p, idx in self.send_to.items()
                },
                'halo_iota': torch.arange(self.n_halo, dtype=torch.int32,
                                          device=device),
            }
        return self._dev

    def fetch(self, vec, callback=None):
        """Halo slices of `vec` as an (n_halo, M) slice-major device buffer;
        cached on the vector until it is written (mpi_vector.py:143-145)."""
        import torch
        from ._lib import check, lib, ptr, stream
        key = self._key
        if key in vec._halo:
            if callback is not None:
                callback()
            return vec._halo[key]
        dev = self._device_arrays(vec.data.device)
        M = vec.M
        halo = torch.empty((self.n_halo, M), dtype=torch.float64,
                           device=vec.data.device)
        sends, recvs = {}, {}
        for p, idx in dev['send_idx'].items():
            buf = torch.empty((len(idx), M), dtype=torch.float64,
                              device=vec.data.device)
            check(lib().stk_pack_slices(ptr(vec.data), vec.ld, M, ptr(idx),
                                        len(idx), ptr(buf), stream()))
            sends[p] = buf
        for p, (off, cnt) in self.recv_from.items():
            recvs[p] = halo[off:off + cnt]
        # post the transfers, run the caller's independent work while they are
        # in flight, then make the stream wait (mpi_vector.py:155-183)
        t0 = _now()
        comm = vec.dofs_distr.comm
        reqs = comm.exchange_begin(sends, recvs)
        if callback is not None:
            self.time_communication = getattr(self, 'time_communication',
                                              0.0) + _now() - t0
            callback()
            t0 = _now()
        comm.exchange_end(reqs)
        self.time_communication = getattr(self, 'time_communication',
                                          0.0) + _now() - t0
        vec._halo[key] = halo
        return halo

    def halo_of_image(self, halo, pair, out0, out1):
        """Halo slices of M x and A x from the halo slices of x: space
        operators act slice by slice, so the neighbours' boundary slices of
        (I (x) M) x are M applied to their boundary slices of x.  `halo`:
        (n_halo, M) slice-major; `pair`: a DeviceCSRPair; returns two
        slice-major buffers that `out0._halo` / `out1._halo` then hold for
        this plan's exchange pattern (one exchange of x instead of two of its
        images, issued before the products it overlaps)."""
        import torch
        from ._lib import check, lib, ptr, stream
        from .mpi_vector import pitch
        nh, M = halo.shape
        ldh = pitch(nh)
        dev = halo.device
        blk = torch.empty((M, ldh), dtype=torch.float64, device=dev)
        check(lib().stk_block_from_rowmajor(ptr(halo), nh, M, ptr(blk), ldh,
                                            stream()))
        b0, b1 = torch.empty_like(blk), torch.empty_like(blk)
        pair.split(blk, b0, b1)
        for b, vec in ((b0, out0), (b1, out1)):
            h = torch.empty((nh, M), dtype=torch.float64, device=dev)
            check(lib().stk_block_to_rowmajor(ptr(b), ldh, nh, M, ptr(h),
                                              stream()))
            vec._halo[self._key] = h

    def apply(self, vec_in, out_block, alpha=1.0, beta=0.0):
        """out_block = alpha * (T (x) I) vec_in + beta * out_block."""
        from ._lib import check, lib, ptr, stream
        dev = self._device_arrays(vec_in.data.device)
        halo = self.fetch(vec_in) if self.n_halo else None
        indptr, indices, vals = dev['local']
        check(lib().stk_time_apply(vec_in.M, self.n_loc, self.local.nnz,
                                   ptr(indptr), ptr(indices), ptr(vals),
                                   ptr(vec_in.data), vec_in.ld, self.n_loc,
                                   ptr(halo), self.n_halo, float(alpha),
                                   float(beta), ptr(out_block), vec_in.ld,
                                   stream()))

    def apply_adjoint(self, vec_in, vec_out):
        """vec_out = (T^T (x) I) vec_in: local partial sums, then the partial
        sums that belong to other ranks' slices are sent there and added (the
        adjoint of `fetch`)."""
        import torch
        from ._lib import check, lib, ptr, stream
        from .mpi_vector import pitch
        dev = self._device_arrays(vec_in.data.device)
        M, n = vec_in.M, self.n_loc
        vec_out._invalidate()
        indptr, indices, vals = dev['adj_local']
        check(lib().stk_time_apply(M, n, self.adj_local.nnz, ptr(indptr),
                                   ptr(indices), ptr(vals), ptr(vec_in.data),
                                   vec_in.ld, n, None, 0, 1.0, 0.0,
                                   ptr(vec_out.data), vec_out.ld, stream()))
        if not self.n_halo and not self.send_to:
            return
        if self.n_halo:
            ldh = pitch(self.n_halo)
            part = torch.empty((M, ldh), dtype=torch.float64,
                               device=vec_in.data.device)
            indptr, indices, vals = dev['adj_halo']
            check(lib().stk_time_apply(M, self.n_halo, self.adj_halo.nnz,
                                       ptr(indptr), ptr(indices), ptr(vals),
                                       ptr(vec_in.data), vec_in.ld, n, None, 0,
                                       1.0, 0.0, ptr(part), ldh, stream()))
            packed = torch.empty((self.n_halo, M), dtype=torch.float64,
                                 device=vec_in.data.device)
            check(lib().stk_pack_slices(ptr(part), ldh, M,
                                        ptr(dev['halo_iota']), self.n_halo,
                                        ptr(packed), stream()))
            self.scatter_add_halo(vec_out, packed)
        else:
            self.scatter_add_halo(vec_out, None)

    def scatter_add_halo(self, vec_out, packed):
        """The adjoint of `fetch`: `packed` (n_halo, M) holds this rank's
        partial sums for slices owned by other ranks; they are sent to their
        owners, and what the peers computed for this rank's slices is added
        into vec_out."""
        import torch
        from ._lib import check, lib, ptr, stream
        dev = self._device_arrays(vec_out.data.device)
        M = vec_out.M
        sends, recvs = {}, {}
        if packed is not None:
            for p, (off, cnt) in self.recv_from.items():
                sends[p] = packed[off:off + cnt]
        for p, idx in dev['send_idx'].items():
            recvs[p] = torch.empty((len(idx), M), dtype=torch.float64,
                                   device=vec_out.data.device)
        t0 = _now()
        vec_out.dofs_distr.comm.exchange(sends, recvs)
        self.time_communication = getattr(self, 'time_communication',
                                          0.0) + _now() - t0
        for p, idx in dev['send_idx'].items():
            check(lib().stk_unpack_slices(ptr(vec_out.data), vec_out.ld, M,
                                          ptr(idx), len(idx), ptr(recvs[p]),
                                          1.0, 1.0, stream()))


class LevelChain:
    """A product of sparse level steps  G_L ... G_2 G_1  (x) I  restricted to
    the extended index set [local slices | halo slices] of a TimeOpPlan.

    The plan's halo set is the dependency closure of the local rows of the
    whole product, so running the restricted steps one after the other on the
    extended column reproduces the local rows exactly; entries of a step that
    point outside the set belong to values no local row depends on and are
    dropped.  Only the rows a step changes are stored.  Host logic (numpy);
    `apply` runs stk_time_chain."""
    def __init__(self, plan, steps):
        self.plan = plan
        n, nh = plan.n_loc, plan.n_halo
        a = plan.dofs_distr.t_begin
        E = np.concatenate([np.arange(a, a + n), plan.halo_cols]).astype(np.int64)
        lev_ptr, trow, tptr, tcol, tval = [0], [], [0], [], []
        eye = sp.identity(len(E), format='csr')
        for G in steps:
            G = sp.csr_matrix(G)[E][:, E].tocsr()
            G.sort_indices()
            changed = np.unique((G - eye).tocoo().row)
            for r in changed:
                lo, hi = G.indptr[r], G.indptr[r + 1]
                trow.append(int(r))
                tcol.extend(G.indices[lo:hi].tolist())
                tval.extend(G.data[lo:hi].tolist())
                tptr.append(len(tcol))
            lev_ptr.append(len(trow))
        self.lev_ptr = np.asarray(lev_ptr, dtype=np.int32)
        self.trow = np.asarray(trow, dtype=np.int32)
        self.tptr = np.asarray(tptr, dtype=np.int32)
        self.tcol = np.asarray(tcol, dtype=np.int32)
        self.tval = np.asarray(tval, dtype=np.float64)
        self.nlev = len(steps)
        self._dev = None

    def apply_host(self, ext):
        """The chain on an extended (n_loc + n_halo, k) host array (tests)."""
        v = np.array(ext, dtype=np.float64, copy=True)
        for lev in range(self.nlev):
            q0, q1 = self.lev_ptr[lev], self.lev_ptr[lev + 1]
            new = [sum(self.tval[p] * v[self.tcol[p]]
                       for p in range(self.tptr[q], self.tptr[q + 1]))
                   for q in range(q0, q1)]
            for q, val in zip(range(q0, q1), new):
                v[self.trow[q]] = val
        return v

    def apply(self, x_block, ldx, M, out_block, ldy, xh=None, yh_out=None):
        import torch
        from ._lib import check, lib, ptr, stream
        dev = x_block.device
        if self._dev is None:
            # never hand an empty tensor's (null) pointer to the kernel
            pad = lambda a: a if len(a) else np.zeros(1, dtype=a.dtype)
            self._dev = tuple(
                torch.from_numpy(pad(a)).to(dev)
                for a in (self.lev_ptr, self.trow, self.tptr, self.tcol,
                          self.tval))
        lev_ptr, trow, tptr, tcol, tval = self._dev
        check(lib().stk_time_chain(M, self.plan.n_loc, self.plan.n_halo,
                                   self.nlev, ptr(lev_ptr), ptr(trow),
                                   ptr(tptr), ptr(tcol), ptr(tval),
                                   len(self.trow), len(self.tcol),
                                   ptr(x_block), ldx, ptr(xh), ptr(out_block),
                                   ldy, ptr(yh_out), stream()))


class TimeOpPlan2:
    """out = (Ta (x) I) va + (Tb (x) I) vb in one pass: the two local CSRs
    stacked side by side, columns [va | vb | halo of va | halo of vb]."""
    def __init__(self, plan_a, plan_b):
        assert plan_a.n_loc == plan_b.n_loc
        self.a, self.b = plan_a, plan_b
        n, ha, hb = plan_a.n_loc, plan_a.n_halo, plan_b.n_halo

        def shifted(loc, col_shift, halo_shift):
            loc = loc.tocoo()
            cols = np.where(loc.col < n, loc.col + col_shift,
                            loc.col - n + halo_shift)
            return sp.coo_matrix((loc.data, (loc.row, cols)),
                                 shape=(n, 2 * n + ha + hb))

        self.local = (shifted(plan_a.local, 0, 2 * n)
                      + shifted(plan_b.local, n, 2 * n + ha)).tocsr()
        self.local.sort_indices()
        self._dev = None

    def apply(self, va, vb, out_block, alpha=1.0, beta=0.0, ldy=None):
        """out_block: a device tensor, or a raw device address of a sub-block
        with pitch ldy (va.ld columns are written)."""
        import torch
        from ._lib import check, lib, ptr, stream
        dev = va.data.device
        if self._dev is None:
            self._dev = tuple(
                torch.from_numpy(a).to(dev)
                for a in (self.local.indptr.astype(np.int32),
                          self.local.indices.astype(np.int32),
                          self.local.data.astype(np.float64)))
        indptr, indices, vals = self._dev
        ha = self.a.fetch(va) if self.a.n_halo else None
        hb = self.b.fetch(vb) if self.b.n_halo else None
        check(lib().stk_time_apply2(va.M, self.a.n_loc, ptr(indptr),
                                    ptr(indices), ptr(vals), ptr(va.data),
                                    ptr(vb.data), va.ld, self.a.n_loc,
                                    ptr(ha), self.a.n_halo, ptr(hb),
                                    float(alpha), float(beta),
                                    out_block if isinstance(out_block, int)
                                    else ptr(out_block),
                                    va.ld if ldy is None else ldy, va.ld,
                                    stream()))


def _now():
    from .comm import Wtime_device
    return Wtime_device()


_neighbour_plans = {}


def neighbour_plan(dofs_distr):
    """The +-1 halo of every tridiagonal time matrix (mpi_vector.py:140-187)."""
    key = id(dofs_distr)
    if key not in _neighbour_plans:
        N = dofs_distr.N
        T = sp.diags([np.ones(N - 1), np.ones(N), np.ones(N - 1)], [-1, 0, 1],
                     format='csr') if N > 1 else sp.identity(1, format='csr')
        _neighbour_plans[key] = (dofs_distr, TimeOpPlan(dofs_distr, T))
    return _neighbour_plans[key][1]
