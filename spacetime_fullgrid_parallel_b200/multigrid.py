"""Space multigrid V-cycle preconditioner, batched over time slices on the GPU.

Drop-in for `MultiGrid` of /root/reference/source/multigrid.py:130-197:
same constructor `MultiGrid(mat, hierarchy, smoothsteps=2, vcycles=1)`, same
Galerkin hierarchy (:140-145), lexicographic Gauss-Seidel pre/post smoothing
(PETSc MatSOR, :100-127), exact coarsest solve (:161-170), `vcycles` cycles
from a zero guess (:184-193).  `hierarchy` is duck-typed: `.J`, `.P_mats`,
`.R_mats` (what MeshHierarchy :14-80 provides).

Difference in kind, not in result: the reference runs one slice at a time
through Python; here one launch sequence serves every local time slice, and a
`MultiGridFamily` serves every matrix of the form  sum_k c_k B_k  (K_x = MG(A_x)
and all C_j = MG(2^j M_x + alpha A_x), heateq_mpi.py:143-153) from ONE
device-resident hierarchy with per-slice coefficients.
"""
import ctypes
import os

import numpy as np
import scipy.sparse as sp
import torch

from . import gs_program
from ._lib import check, lib, ptr, stream
from .linop import (ROW_ORDER_MIN_ROWS, _RowOrder, host_apply, locality_order,
                    row_order_enabled)
from .mpi_vector import _device

MAX_COARSE = 1024
# fused smoother (gs_program.py / csrc/stk_gsfused.cu): time values per CTA,
# shared memory the window + kind tables may take, smallest level worth it
FUSED_T = int(os.environ.get('STK_GS_T', '8'))
FUSED_SMEM = 224 * 1024
FUSED_MIN_ROWS = int(os.environ.get('STK_GS_FUSED_MIN_ROWS', '256'))
FUSED_NGRP = int(os.environ.get('STK_GS_NGRP', '128'))  # 512 threads / 4 lanes
MAX_GROUPS = 16  # value tables the fused kernel keeps in shared memory


def fused_enabled():
    return os.environ.get('STK_GS_FUSED', '1') != '0'


def _dev_bytes(a, device, keep):
    """Upload a NumPy array of any integer width as raw bytes."""
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(device)
    keep.append(t)
    return t


class FusedLevel:
    """Device-resident programs of one level (nu forward and nu backward
    sweeps), compiled once from the pattern and the BASE matrices' values; any
    set of group matrices on that pattern can run them (`values_for`)."""
    def __init__(self, indptr, indices, wave, nsweeps, base_values, device,
                 chunks=33, sms=148, T=None, generic=False, capacity=None,
                 ngrp=None, chunks_wide=None):
        self.ok = False
        self.handles = []
        self.handles_wide, self.programs_wide, self.wide_min_chunks = [], [], 0
        self._keep = []
        self.T = T = FUSED_T if T is None else T
        self.indptr = np.asarray(indptr, dtype=np.int64)
        self.maxnnz = int(np.diff(indptr).max())
        self.canon = gs_program.canonical_order(indptr, indices, base_values)
        # the kernel's branch-free row product holds <= 8 entries per row
        kinds = None if generic or self.maxnnz > 8 else gs_program.row_kinds(
            indptr, base_values, self.canon)
        if kinds is None:
            self.kind_of_row, self.rep = None, None
            self.nkinds, self.bulk_kind, tab_bytes = 0, -1, 0
        else:
            self.kind_of_row, self.rep = kinds
            self.nkinds = len(self.rep)
            self.bulk_kind = int(np.argmax(np.bincount(self.kind_of_row)))
            # room for the value tables of up to MAX_GROUPS groups
            tab_bytes = 8 * self.nkinds * (MAX_GROUPS * (self.maxnnz + 3) + T
                                           + 2) + 16
        ngrp = FUSED_NGRP if ngrp is None else ngrp
        if capacity is None:
            # what the window may take: shared memory minus the value tables,
            # the record ring and the f ring (csrc/stk_gsfused.cu)
            recw = 8 + 4 * max(0, (self.maxnnz - 8 + 7) // 8)
            rings = 8 * ngrp * recw * 4 + 64 + gs_program.PREFETCH * ngrp * T * 8
            capacity = min(65535, (FUSED_SMEM - tab_bytes - rings) // (8 * T))
        n = len(indptr) - 1
        emb = gs_program.graph_embedding(
            np.ascontiguousarray(indptr, dtype=np.int32),
            np.ascontiguousarray(indices, dtype=np.int32), n)
        if emb is None:
            return
        progs = []
        for backward in (False, True):
            pg = gs_program.compile_program(
                indptr, indices, wave, nsweeps, backward, capacity,
                embedding=emb, kind_of_row=self.kind_of_row, canon=self.canon,
                chunks=chunks, sms=sms, ngrp=ngrp,
                tiling=progs[0].tiling if progs else None)
            if pg is None:
                return
            progs.append(pg)
        self.programs = progs
        # The same level tiled for blocks twice as wide (the Schur operator
        # solves for both of its brackets in one block): with more time chunks
        # per item, fewer and longer items are better.  Kept only if the
        # compiler indeed chooses another tiling.
        if chunks_wide and chunks_wide > chunks and os.environ.get(
                'STK_GS_WIDE', '1') != '0':
            wide = []
            for backward in (False, True):
                pg = gs_program.compile_program(
                    indptr, indices, wave, nsweeps, backward, capacity,
                    embedding=emb, kind_of_row=self.kind_of_row,
                    canon=self.canon, chunks=chunks_wide, sms=sms, ngrp=ngrp,
                    tiling=wide[0].tiling if wide else None)
                if pg is None or pg.tiling == progs[0].tiling:
                    wide = []
                    break
                wide.append(pg)
            self.programs_wide = wide
            self.wide_min_chunks = (chunks + chunks_wide + 1) // 2

        def upload(pg):
            d = [_dev_bytes(x, device, self._keep)
                 for x in (pg.item_step, pg.item_pass, pg.step_info, pg.op,
                           pg.ld)]
            h = ctypes.c_void_p(lib().stk_gs_prog_create(
                pg.nitems, pg.nslots, pg.maxnnz, int(pg.generic), pg.recw,
                pg.ngrp, *[ptr(t) for t in d]))
            assert h.value, 'stk_gs_prog_create failed'
            return h

        self.handles = [upload(pg) for pg in progs]
        self.handles_wide = [upload(pg) for pg in self.programs_wide]
        self.device = device
        # row kinds on the device: the grouped residual SpMM reads the groups'
        # matrices from the kind table (stk_mg_set_fused)
        self.d_kind = self.d_cidx = None
        if self.kind_of_row is not None:
            self.d_kind = _dev_bytes(np.asarray(self.kind_of_row, dtype=np.int32),
                                     device, self._keep)
            self.d_cidx = _dev_bytes(
                np.asarray(indices, dtype=np.int32)[self.canon], device,
                self._keep)
        self.ok = True

    def values_for(self, group_values):
        """Device arrays that let the programs run with these G matrices
        (CSR-order value arrays on the level's pattern): (ktab, cvals), one of
        them None, or None if the programs cannot (they were compiled with row
        kinds and some group's values differ inside a kind)."""
        G = len(group_values)
        canon_vals = [np.ascontiguousarray(np.asarray(v)[self.canon])
                      for v in group_values]
        if self.rep is None:  # generic programs: values at the CSR offsets
            t = torch.from_numpy(np.stack(canon_vals)).to(self.device)
            return None, t
        if G > MAX_GROUPS or not gs_program.kinds_hold(
                self.indptr, self.kind_of_row, self.rep, canon_vals):
            return None
        tab = np.zeros((G, self.nkinds, self.maxnnz + 2))
        for k, r in enumerate(self.rep):
            p0, p1 = self.indptr[r], self.indptr[r + 1]
            for g, v in enumerate(canon_vals):
                tab[g, k, :p1 - p0] = v[p0:p1]
        return torch.from_numpy(tab).to(self.device), None

    def sweeps(self, backward, vals, grp, f, u_in, u_out):
        """u_out <- the level's nu sweeps (device blocks) with the matrices
        `vals` = values_for(...); tests / microbench."""
        ktab, cvals = vals
        G = (ktab if ktab is not None else cvals).shape[0]
        check(lib().stk_gs_fused(self.handles[int(bool(backward))], G, self.T,
                                 ptr(ktab), self.nkinds, self.bulk_kind,
                                 ptr(cvals), int(self.indptr[-1]), ptr(grp),
                                 ptr(f), ptr(u_in), ptr(u_out), f.shape[1],
                                 stream()))

    def attach(self, mg_handle, level, group_values, keep):
        """Returns False if these values cannot use the programs."""
        vals = self.values_for(group_values)
        if vals is None:
            return False
        keep.extend(t for t in vals if t is not None)
        check(lib().stk_mg_set_fused(mg_handle, level, self.handles[0],
                                     self.handles[1], ptr(vals[0]),
                                     self.nkinds, self.bulk_kind, ptr(vals[1]),
                                     self.T, self.maxnnz + 2, ptr(self.d_kind),
                                     ptr(self.d_cidx)))
        if self.handles_wide:
            check(lib().stk_mg_set_fused_wide(mg_handle, level,
                                              self.handles_wide[0],
                                              self.handles_wide[1],
                                              self.wide_min_chunks))
        return True

    def __del__(self):
        try:
            for h in self.handles + self.handles_wide:
                lib().stk_gs_prog_destroy(h)
            self.handles, self.handles_wide = [], []
        except Exception:
            pass


def _project(pattern_keys, mat, ncols, pattern=None):
    """Values of `mat` laid out on the (sorted) union pattern."""
    mat = sp.csr_matrix(mat)
    mat.sort_indices()
    if (pattern is not None and mat.nnz == pattern.nnz
            and np.array_equal(mat.indptr, pattern.indptr)
            and np.array_equal(mat.indices, pattern.indices)):
        return mat.data.astype(np.float64)
    rows = np.repeat(np.arange(mat.shape[0], dtype=np.int64),
                     np.diff(mat.indptr))
    keys = rows * ncols + mat.indices
    pos = np.searchsorted(pattern_keys, keys)
    vals = np.zeros(len(pattern_keys))
    vals[pos] = mat.data
    return vals


def gauss_seidel_schedule(indptr, indices, return_wave=False):
    """Wavefronts of the lexicographic sweep: (rows ordered wavefront by
    wavefront, phase_ptr).  Host, setup only."""
    n = len(indptr) - 1
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    wave = np.zeros(max(n, 1), dtype=np.int32)
    # A row must come after every lower-numbered row it is coupled to IN EITHER
    # DIRECTION (row j > i reads u_i, and row i reads u_j: neither may overtake
    # the other), so the wavefronts come from the symmetrised pattern; for the
    # structurally symmetric FE matrices that is the pattern itself.
    pat = sp.csr_matrix((np.ones(len(indices), dtype=np.int8), indices, indptr),
                        shape=(n, n))
    sym = (pat + pat.T).tocsr()
    if sym.nnz != pat.nnz:
        sym.sort_indices()
        wp = np.ascontiguousarray(sym.indptr, dtype=np.int32)
        wi = np.ascontiguousarray(sym.indices, dtype=np.int32)
    else:
        wp, wi = indptr, indices
    depth = lib().stk_gs_wavefronts(n, wp.ctypes.data, wi.ctypes.data,
                                    wave.ctypes.data)
    wave = wave[:n]
    # Rows of a wavefront are independent, so their order is free: sort them by
    # their highest-numbered neighbour.  With hierarchical numberings the
    # coarse ("old") vertices of a level are scattered in index space, but
    # their highest neighbours are the new vertices, which are numbered along
    # the mesh -- the sweep then walks every wavefront in mesh order and
    # neighbouring rows are re-read from L2 instead of HBM (measured: -40 %
    # DRAM reads in the first wavefront).
    maxcol = np.maximum.reduceat(
        indices, indptr[:-1].astype(np.int64)) if n and len(indices) else \
        np.zeros(n, dtype=np.int32)
    order = np.lexsort((maxcol, wave)).astype(np.int32)
    counts = np.bincount(wave, minlength=max(depth, 1))
    phase_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    if return_wave:
        return order, phase_ptr, wave
    return order, phase_ptr


class MGContext:
    """The groups of a block with pitch ld: slices with equal coefficients
    share a group; device arrays for the kernels, the handle that holds the
    groups' hierarchies."""
    def __init__(self, family, coefs_per_slice, ld):
        n = len(coefs_per_slice)
        assert 1 <= n <= ld
        distinct = []
        group = np.zeros(ld, dtype=np.int32)
        for t in range(ld):
            c = tuple(float(v) for v in coefs_per_slice[min(t, n - 1)])
            assert len(c) == family.K
            if c not in distinct:
                distinct.append(c)
            group[t] = distinct.index(c)
        self.groups = tuple(distinct)
        self.handle, self.inv = family.handle_for(self.groups)
        self.group = (torch.from_numpy(group).to(family.device)
                      if len(distinct) > 1 else None)
        self.ld = ld


_ws = {}


def _workspace(device, doubles):
    cur = _ws.get(device)
    if cur is None or cur.numel() < doubles:
        _ws[device] = None
        cur = _ws[device] = torch.empty(int(doubles), dtype=torch.float64,
                                        device=device)
    return cur


class MultiGridFamily:
    """All V-cycle preconditioners MG(sum_k c_k base_mats[k]) on one mesh
    hierarchy.  Pattern, Gauss-Seidel schedule, transfer operators and the
    fused smoother programs are built once; a set of coefficient tuples (the
    GROUPS of a block's time slices) gets a handle whose level matrices are the
    Galerkin products of each combined matrix, formed exactly as the reference
    forms them (heateq_mpi.py:97-98,143-153; multigrid.py:140-145)."""
    def __init__(self, base_mats, hierarchy, smoothsteps=2, vcycles=1,
                 device=None, ld_hint=None):
        self.K = K = len(base_mats)
        assert K >= 1
        self.base_mats = [sp.csr_matrix(B, dtype=np.float64) for B in base_mats]
        self.hierarchy = hierarchy
        self.smoothsteps, self.vcycles = smoothsteps, vcycles
        self.device = dev = _device() if device is None else device
        self.J = J = hierarchy.J
        mats = [self._galerkin(B) for B in self.base_mats]
        self.level_mats = mats
        self.shape = mats[0][-1].shape
        n0 = mats[0][0].shape[0]
        if n0 > MAX_COARSE:
            raise ValueError('coarsest level has %d dofs (> %d): the dense '
                             'coarse solve needs a coarser mesh' %
                             (n0, MAX_COARSE))
        self._ctx_cache = {}
        self._handles = {}  # groups -> (handle, inverses, kept tensors)
        self._keep = []  # device tensors shared by the handles
        self.num_phases = []
        self._levels = []
        self._orders = []  # registered row schedules (kept alive)

        def up(a, dt):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
            self._keep.append(t)
            return t

        for l in range(J + 1):
            n = mats[0][l].shape[0]
            pat = mats[0][l].copy()
            pat.data = np.ones_like(pat.data)
            for k in range(1, K):
                q = mats[k][l].copy()
                q.data = np.ones_like(q.data)
                pat = pat + q
            pat = pat.tocsr()
            pat.sort_indices()
            rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(pat.indptr))
            keys = rows * n + pat.indices
            base_vals = [_project(keys, mats[k][l], n, pat) for k in range(K)]
            order, phase_ptr, wave = gauss_seidel_schedule(
                pat.indptr, pat.indices, return_wave=True)
            self.num_phases.append(len(phase_ptr) - 1)
            fused = None
            if (l >= 1 and fused_enabled() and smoothsteps > 0
                    and n >= FUSED_MIN_ROWS and len(phase_ptr) - 1 <= 8):
                chunks = -(-(ld_hint or 256) // FUSED_T)
                fused = FusedLevel(pat.indptr, pat.indices, wave, smoothsteps,
                                   base_vals, dev, chunks=chunks,
                                   chunks_wide=-(-2 * (ld_hint or 256) // FUSED_T),
                                   sms=torch.cuda.get_device_properties(
                                       dev).multi_processor_count)
                if not fused.ok:
                    fused = None
            # row schedules of the level's SpMMs (residual; prolongation walks
            # the fine rows, restriction the coarse rows, in their levels'
            # locality orders)
            loc = (locality_order(pat)
                   if n >= ROW_ORDER_MIN_ROWS and row_order_enabled() else None)
            lv = {'n': n, 'nnz': int(pat.nnz), 'pattern': pat, 'keys': keys,
                  'locality': loc,
                  'indptr': up(pat.indptr, np.int32),
                  'indices': up(pat.indices, np.int32),
                  'order': up(order, np.int32), 'phase_ptr': phase_ptr,
                  'transfer': None, 'fused': fused}
            if l >= 1:
                P = sp.csr_matrix(hierarchy.P_mats[l - 1], dtype=np.float64)
                R = sp.csr_matrix(hierarchy.R_mats[l - 1], dtype=np.float64)
                P.sort_indices()
                R.sort_indices()
                lv['transfer'] = [up(P.indptr, np.int32), up(P.indices, np.int32),
                                  up(P.data, np.float64), up(R.indptr, np.int32),
                                  up(R.indices, np.int32), up(R.data, np.float64)]
                self._orders.append(_RowOrder(lv['transfer'][0], n, loc, dev))
                self._orders.append(_RowOrder(lv['transfer'][3], R.shape[0],
                                              self._levels[l - 1]['locality'],
                                              dev))
            self._orders.append(_RowOrder(lv['indptr'], n, loc, dev))
            self._levels.append(lv)

    def _galerkin(self, mat):
        """Level matrices, coarse from fine (multigrid.py:140-145)."""
        lv = [sp.csr_matrix(mat, dtype=np.float64)]
        for j in reversed(range(self.J)):
            lv.insert(0, (self.hierarchy.R_mats[j] @ lv[0]
                          @ self.hierarchy.P_mats[j]).tocsr())
        return lv

    def combined(self, coefs):
        """sum_k coefs[k] * base_mats[k], the expression of heateq_mpi.py:97-98."""
        mat = coefs[0] * self.base_mats[0]
        for c, B in zip(coefs[1:], self.base_mats[1:]):
            mat = mat + c * B
        return sp.csr_matrix(mat)

    def __del__(self):
        try:
            for h, _inv, _keep in list(getattr(self, '_handles', {}).values()):
                lib().stk_mg_destroy(h)
            self._handles = {}
        except Exception:
            pass

    def handle_for(self, groups):
        """(stk_mg handle, device coarse inverses) for a tuple of coefficient
        tuples; built on first use."""
        groups = tuple(tuple(float(c) for c in g) for g in groups)
        if groups not in self._handles:
            G = len(groups)
            chains = [self._galerkin(self.combined(g)) for g in groups]
            h = ctypes.c_void_p(lib().stk_mg_create(
                len(self._levels), self.smoothsteps, self.vcycles, G))
            assert h.value, 'stk_mg_create failed'
            keep = []
            for l, lv in enumerate(self._levels):
                vals = [_project(lv['keys'], ch[l], lv['n'], lv['pattern'])
                        for ch in chains]
                diag = [ch[l].diagonal() for ch in chains]
                dv = torch.from_numpy(np.stack(vals)).to(self.device)
                dd = torch.from_numpy(np.stack(diag)).to(self.device)
                keep += [dv, dd]
                check(lib().stk_mg_set_level(
                    h, l, lv['n'], lv['nnz'], ptr(lv['indptr']),
                    ptr(lv['indices']), ptr(dv), ptr(dd), ptr(lv['order']),
                    lv['phase_ptr'].ctypes.data, len(lv['phase_ptr']) - 1))
                if lv['transfer'] is not None:
                    check(lib().stk_mg_set_transfer(
                        h, l, *[ptr(t) for t in lv['transfer']]))
                if lv['fused'] is not None:
                    lv['fused'].attach(h, l, vals, keep)
            # coarsest level: exact inverse (the reference factorises it with
            # SuperLU, multigrid.py:161-165; it is 1x1 for `square`)
            inv = np.stack([np.linalg.inv(ch[0].toarray()) for ch in chains])
            dinv = torch.from_numpy(np.ascontiguousarray(inv)).to(self.device)
            self._handles[groups] = (h, dinv, keep)
        h, dinv, _ = self._handles[groups]
        return h, dinv

    def coarse_inverse(self, coefs):
        return np.linalg.inv(self._galerkin(self.combined(coefs))[0].toarray())

    def context(self, coefs_per_slice, ld):
        key = (tuple(tuple(float(v) for v in c) for c in coefs_per_slice), ld)
        if key not in self._ctx_cache:
            self._ctx_cache[key] = MGContext(self, coefs_per_slice, ld)
        return self._ctx_cache[key]

    def apply_uniform(self, coefs, b, x):
        """x <- MG(sum_k coefs[k] B_k) b, the same matrix for every slice."""
        return self.apply_block(b, x, self.context([coefs], b.shape[1]))

    def apply_block(self, b, x, ctx):
        """x <- MG(A_{group(t)}) b for every slice t of the block."""
        ld = b.shape[1]
        assert ctx.ld == ld and b.shape[0] == self.shape[0]
        ws = _workspace(b.device, lib().stk_mg_workspace(ctx.handle, ld))
        check(lib().stk_mg_apply(ctx.handle, ptr(ctx.group), ptr(ctx.inv),
                                 ptr(b), ptr(x), ld, ptr(ws), stream()))

    def member(self, coefs):
        """The MultiGrid operator of sum_k coefs[k] * base_mats[k]."""
        return MultiGrid(None, self.hierarchy, self.smoothsteps, self.vcycles,
                         _family=self, _coefs=coefs)


class Smoother:
    """Lexicographic Gauss-Seidel on one level matrix (the semantics of
    multigrid.py:83-97 / PETScSMoother :100-127), on the device: host arrays
    u, f of shape (n,) or (n, k); `PreSmooth` sweeps forward, `PostSmooth`
    backward, `its` times, updating u in place."""
    def __init__(self, mat, its=1):
        self.its = its
        mat = sp.csr_matrix(mat, dtype=np.float64)
        mat.sort_indices()
        self.n = mat.shape[0]
        self.device = dev = _device()
        order, phase_ptr = gauss_seidel_schedule(mat.indptr, mat.indices)
        self._keep = [
            torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
            for a, dt in ((mat.indptr, np.int32), (mat.indices, np.int32),
                          (mat.data, np.float64), (mat.diagonal(), np.float64),
                          (order, np.int32))
        ]
        self.handle = ctypes.c_void_p(lib().stk_mg_create(2, its, 1, 1))
        ip, ix, dv, dd, od = self._keep
        check(lib().stk_mg_set_level(self.handle, 1, self.n, int(mat.nnz),
                                     ptr(ip), ptr(ix), ptr(dv), ptr(dd),
                                     ptr(od), phase_ptr.ctypes.data,
                                     len(phase_ptr) - 1))

    def __del__(self):
        try:
            if self.handle and self.handle.value:
                lib().stk_mg_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _sweep(self, u, f, backward):
        from .mpi_vector import pitch
        u2 = np.asarray(u, dtype=np.float64).reshape(self.n, -1)
        f2 = np.asarray(f, dtype=np.float64).reshape(self.n, -1)
        k = u2.shape[1]
        ld = pitch(k)
        dev = self.device
        du = torch.zeros((self.n, ld), dtype=torch.float64, device=dev)
        df = torch.zeros((self.n, ld), dtype=torch.float64, device=dev)
        du[:, :k] = torch.from_numpy(np.ascontiguousarray(u2)).to(dev)
        df[:, :k] = torch.from_numpy(np.ascontiguousarray(f2)).to(dev)
        check(lib().stk_mg_smooth(self.handle, 1, self.its, int(backward),
                                  None, ptr(df), ptr(du), ld, stream()))
        u[...] = du[:, :k].cpu().numpy().reshape(np.shape(u))

    def PreSmooth(self, u, f):
        self._sweep(u, f, False)

    def PostSmooth(self, u, f):
        self._sweep(u, f, True)


class MultiGrid:
    """V-cycle preconditioner for one matrix (multigrid.py:130-197)."""
    def __init__(self, mat, hierarchy, smoothsteps=2, vcycles=1, _family=None,
                 _coefs=None):
        self.num_applies = 0
        self.time_applies = 0
        self.hierarchy = hierarchy
        self.smoothsteps = smoothsteps
        self.vcycles = vcycles
        if _family is None:
            _family = MultiGridFamily([mat], hierarchy, smoothsteps, vcycles)
            _coefs = (1.0, )
        self.family = _family
        self.coefs = tuple(float(c) for c in _coefs)
        assert len(self.coefs) == _family.K
        self.shape = _family.shape
        self.dtype = np.dtype(np.float64)

    @property
    def mats(self):
        """Level matrices, coarse to fine (multigrid.py:140-154)."""
        return self.family._galerkin(self.family.combined(self.coefs))

    def apply_block(self, x, out, ctx=None):
        if ctx is None:
            self.family.apply_uniform(self.coefs, x, out)
        else:
            self.family.apply_block(x, out, ctx)
        self.num_applies += 1

    # SciPy LinearOperator-style host interface (M,) / (M, k)
    def __matmul__(self, B):
        return host_apply(self, B)

    matvec = matmat = dot = _matvec = __matmul__

    def time_per_apply(self):
        assert self.time_applies
        return self.time_applies / self.num_applies
