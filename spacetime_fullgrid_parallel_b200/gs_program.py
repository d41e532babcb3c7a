"""Setup-time compiler of the fused Gauss-Seidel smoother (host, once per level).

The reference smoother is nu lexicographic Gauss-Seidel sweeps
(/root/reference/source/multigrid.py:89-97 and the PETSc MatSOR form :113-127).
Launching one kernel per wavefront of that sweep reproduces it exactly but
streams the level's block through HBM ~6 times per sweep (profiles/traffic.json,
round 1).  This module turns `nu` sweeps over one level into a *program* for
`k_gs_fused` (csrc/stk_gsfused.cu) that touches HBM about once:

  * the rows are cut into spatial **items** (strip x segment of a graph
    embedding computed from BFS distances -- no geometry is needed);
  * for one item and one chunk of T time slices, a CTA marches along the
    embedding's first coordinate and runs ALL nu * D stages (D = number of
    wavefronts of the sweep) as a skewed pipeline on a sliding **window** of
    rows held in shared memory: stage k works LAG columns behind stage k-1;
  * rows outside the item that its updates depend on (the dependency closure,
    <= 1 hop per stage) are recomputed redundantly, so items never wait on one
    another -- no flags, no grid barriers -- and every (row, sweep) update sees
    exactly the operands of the sequential sweep: the iterates are those of the
    reference up to the rounding of the row sums.

A program is a list of macro-steps; each holds the window loads to issue, and
the row updates ("ops") that may run concurrently.  Between macro-steps the CTA
synchronises.  Correctness never depends on the embedding or the tiling: every
op is placed after the ops that produce its operands (`_schedule_item`), and a
window slot is recycled only after its last reader (`stk_gs_alloc_slots`).
`emulate()` in tests/gs_emulator.py interprets the same arrays with NumPy.
"""
import ctypes

import numpy as np

from ._lib import lib

LOOKAHEAD = 2          # macro-steps between issuing a window load and using it
LAG = 2                # columns between consecutive stages of the pipeline
PREFETCH = 4           # the record names the row of the op this many passes later
MAX_PASSES = 3         # passes per macro-step (the kernel's record ring holds 8)
MAX_STAGES = 48
MAX_KINDS = 64


def _expand(indptr, rows):
    """Positions in `indices` of the entries of `rows`, concatenated, and the
    segment starts (len(rows) + 1)."""
    starts = indptr[rows].astype(np.int64)
    lens = indptr[rows + 1].astype(np.int64) - starts
    seg = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=seg[1:])
    pos = np.arange(seg[-1], dtype=np.int64) + np.repeat(starts - seg[:-1], lens)
    return pos, seg


def bfs_distance(indptr, indices, source, n):
    dist = np.full(n, -1, dtype=np.int32)
    frontier = np.atleast_1d(np.asarray(source, dtype=np.int64))
    dist[frontier] = 0
    d = 0
    while len(frontier):
        pos, _ = _expand(indptr, frontier)
        nb = indices[pos]
        nb = np.unique(nb[dist[nb] < 0])
        d += 1
        dist[nb] = d
        frontier = nb.astype(np.int64)
    return dist


def graph_embedding(indptr, indices, n):
    """Two integer coordinates per row from BFS distances (a two-axis
    'high-dimensional embedding'): x = d(a,.) - d(a',.) for a pseudo-peripheral
    pair (a, a'), y the same for the two ends (b, b') of the middle level set of
    x.  Returns None for a disconnected graph."""
    if n == 0:
        return None
    d0 = bfs_distance(indptr, indices, 0, n)
    if (d0 < 0).any():
        return None
    a = int(np.argmax(d0))
    da = bfs_distance(indptr, indices, a, n)
    a2 = int(np.argmax(da))
    da2 = bfs_distance(indptr, indices, a2, n)
    x = da.astype(np.int64) - da2
    mid = np.nonzero(np.abs(x - np.median(x)) <= 1)[0]
    dm = bfs_distance(indptr, indices, int(mid[0]), n)
    b = int(mid[np.argmax(dm[mid])])
    db = bfs_distance(indptr, indices, b, n)
    b2 = int(mid[np.argmax(db[mid])])
    db2 = bfs_distance(indptr, indices, b2, n)
    y = db.astype(np.int64) - db2
    return x, y


def canonical_order(indptr, indices, value_arrays):
    """Per row, the order in which a program lists the row's entries: the
    diagonal first, then the off-diagonal entries sorted by their values in
    the base matrices (ties: by column).  Rows of a uniformly refined mesh
    whose stencils differ only by the numbering of the neighbours then carry
    the same value sequence and share a *kind* (row_kinds).  Returns `canon`:
    canon[indptr[i] + e] = CSR position of entry e of row i."""
    n = len(indptr) - 1
    nnz = np.diff(indptr).astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), nnz)
    cols = np.asarray(indices, dtype=np.int64)
    keys = [cols]
    for v in reversed(list(value_arrays)):
        keys.append(np.asarray(v, dtype=np.float64))
    keys.append(cols != rows)  # the diagonal entry first
    keys.append(rows)
    return np.lexsort(keys).astype(np.int64)


def row_kinds(indptr, value_arrays, canon=None):
    """Rows with bitwise identical value sequences (all arrays, in canonical
    entry order) share a *kind*: on a uniformly refined mesh a level has a
    handful.  Returns (kind_of_row, representative_row_of_kind) or None if
    there are more than MAX_KINDS (unstructured mesh: generic path)."""
    n = len(indptr) - 1
    if n == 0:
        return None
    nnz = np.diff(indptr).astype(np.int64)
    if (nnz == 0).any():
        return None
    if canon is None:
        canon = np.arange(int(nnz.sum()), dtype=np.int64)
    rng = np.random.RandomState(12345)
    maxnnz = int(nnz.max())
    mult = rng.randint(1, 2**62, size=(len(value_arrays), maxnnz),
                       dtype=np.int64).astype(np.uint64) * np.uint64(2) + np.uint64(1)
    h = nnz.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    starts = indptr[:-1].astype(np.int64)
    within = np.arange(len(canon), dtype=np.int64) - np.repeat(starts, nnz)
    vals = [np.ascontiguousarray(np.asarray(v, dtype=np.float64)[canon])
            for v in value_arrays]
    with np.errstate(over='ignore'):
        for k, v in enumerate(vals):
            contrib = v.view(np.uint64) * mult[k][within]
            h += np.add.reduceat(contrib, starts) * np.uint64(k * 2 + 3)
    uniq, first, inv = np.unique(h, return_index=True, return_inverse=True)
    if len(uniq) > MAX_KINDS:
        return None
    # exact verification: every row equals its representative bit for bit
    if not kinds_hold(indptr, inv, first, vals):
        return None
    return inv.astype(np.int32), first.astype(np.int64)


def kinds_hold(indptr, kind_of_row, rep, canon_value_arrays):
    """True if every row carries exactly (bit for bit) the value sequence of
    its kind's representative row, for every array (already in canonical
    entry order)."""
    nnz = np.diff(indptr).astype(np.int64)
    rep_row = np.asarray(rep)[np.asarray(kind_of_row)]
    if not np.array_equal(nnz, nnz[rep_row]):
        return False
    starts = indptr[:-1].astype(np.int64)
    within = np.arange(int(nnz.sum()), dtype=np.int64) - np.repeat(starts, nnz)
    rep_pos = np.repeat(starts[rep_row], nnz) + within
    for v in canon_value_arrays:
        v = np.ascontiguousarray(v, dtype=np.float64).view(np.uint64)
        if not np.array_equal(v, v[rep_pos]):
            return False
    return True


def _alloc_slots(start, end):
    """Interval colouring: slot per row so that two rows share a slot only if
    one's [start, end] ends before the other's starts.  Returns (slot, nslots)."""
    n = len(start)
    start = np.ascontiguousarray(start, dtype=np.int32)
    end = np.ascontiguousarray(end, dtype=np.int32)
    slot = np.zeros(n, dtype=np.int32)
    used = lib().stk_gs_alloc_slots(n, start.ctypes.data, end.ctypes.data,
                                    slot.ctypes.data)
    assert used >= 0
    return slot, int(used)


class _Scratch:
    """n-sized work arrays shared by the items of one level."""
    def __init__(self, n):
        self.g2l = np.full(n, -1, dtype=np.int64)
        self.mark = np.zeros(n, dtype=bool)


def _schedule_item(indptr, indices, wave, stage_colour, own, colidx, scratch):
    """Ops, loads and slots of one item.  Returns a dict of local arrays."""
    S = len(stage_colour)
    # ---- candidate region: S-hop neighbourhood of the own rows ----
    mark = scratch.mark
    mark[own] = True
    region = [own]
    frontier = own
    for _ in range(S):
        pos, _seg = _expand(indptr, frontier)
        nb = indices[pos]
        nb = np.unique(nb[~mark[nb]])
        if len(nb) == 0:
            break
        mark[nb] = True
        region.append(nb.astype(np.int64))
        frontier = nb.astype(np.int64)
    region = np.sort(np.concatenate(region))
    mark[region] = False
    g2l = scratch.g2l
    nl = len(region)
    g2l[region] = np.arange(nl)
    try:
        lwave = wave[region]
        lcol = colidx[region].astype(np.int64)
        # ---- dependency closure, last stage first ----
        need = np.zeros(nl, dtype=bool)
        need[g2l[own]] = True
        comp = [None] * S
        for k in range(S - 1, -1, -1):
            ck = np.nonzero(need & (lwave == stage_colour[k]))[0]
            comp[k] = ck
            if len(ck):
                pos, _seg = _expand(indptr, region[ck])
                need[g2l[indices[pos]]] = True
        loaded = np.nonzero(need)[0]  # rows whose initial value enters the window
        # ---- macro-steps: after every producer, never before the column ----
        BIG = np.iinfo(np.int64).max // 4
        ready = np.full(nl, BIG, dtype=np.int64)  # step after which the value is valid
        ready[loaded] = lcol[loaded] - 2
        first_use = np.full(nl, BIG, dtype=np.int64)
        last_use = np.full(nl, -BIG, dtype=np.int64)
        op_row, op_step, op_stage = [], [], []
        for k in range(S):
            ck = comp[k]
            if len(ck) == 0:
                continue
            pos, seg = _expand(indptr, region[ck])
            nbl = g2l[indices[pos]]
            dep = np.maximum.reduceat(ready[nbl], seg[:-1])
            m = np.maximum(lcol[ck] + LAG * k, dep + 1)
            ready[ck] = m
            mm = np.repeat(m, np.diff(seg))
            np.minimum.at(first_use, nbl, mm)
            np.maximum.at(last_use, nbl, mm)
            op_row.append(ck)
            op_step.append(m)
            op_stage.append(np.full(len(ck), k, dtype=np.int64))
        op_row = np.concatenate(op_row)
        op_step = np.concatenate(op_step)
        op_stage = np.concatenate(op_stage)
        # final update of an own row -> stored to global memory
        is_own = np.zeros(nl, dtype=bool)
        is_own[g2l[own]] = True
        last_stage = np.full(nl, -1, dtype=np.int64)
        np.maximum.at(last_stage, op_row, op_stage)
        op_store = is_own[op_row] & (last_stage[op_row] == op_stage)
        assert op_store.sum() == len(own), 'every own row has a final update'
        # ---- window slots ----
        landed = lcol[loaded] - 2
        issue = landed - LOOKAHEAD
        shift = -int(issue.min())
        issue = issue + shift
        op_step = op_step + shift
        endl = last_use[loaded] + shift
        assert (endl >= issue + LOOKAHEAD).all()
        slot_l, nslots = _alloc_slots(issue, endl)
        slot = np.full(nl, -1, dtype=np.int64)
        slot[loaded] = slot_l
        order = np.lexsort((op_row, op_stage, op_step))
        ld_order = np.lexsort((loaded, issue))
        return {
            'region': region, 'nslots': nslots,
            'nsteps': int(op_step.max()) + 1,
            'op_row': op_row[order], 'op_step': op_step[order],
            'op_store': op_store[order], 'slot': slot,
            'ld_row': loaded[ld_order], 'ld_step': issue[ld_order],
        }
    finally:
        g2l[region] = -1


class GSProgram:
    """Device-ready arrays of one (level, direction) program.

    items:    item_step[nitems + 1], item_pass[nitems + 1]: step / pass ranges
    steps:    step_info[nsteps, 2] = (end pass, end load) of every macro-step
    records:  op: npasses x (recw / 4) x ngrp x 4 uint32 (per pass: word
              quadruples 0 of all records, then quadruples 1, ...;
              records_aos() gives [npasses * ngrp, recw]),
              recw = 8 + 4 * ceil((maxnnz - 8) / 8);
              record pass * ngrp + g is run by thread group g in that pass:
                [0] row | store << 31
                [1] row of the record PREFETCH passes later (same g): its f
                    value is fetched while this op computes
                [2] kind of the row, or the row's offset in the CSR arrays
                [3] window slot | nnz << 16   (nnz = 0: padding, no-op)
                [4:] uint16 window slot of every entry of the row in the
                    program's entry order (`canon`: the diagonal first);
                    unused entries name the row's own slot
    loads:    ld[nloads, 2] int32 = (row, window slot)
    """
    def __init__(self):
        self.stats = {}

    def records_aos(self):
        q = self.recw // 4
        return (self.op.reshape(-1, q, self.ngrp, 4).transpose(0, 2, 1, 3)
                .reshape(-1, self.recw))


def stage_colours(D, nsweeps, backward):
    one = list(range(D))
    if backward:
        one = one[::-1]
    return one * nsweeps


def _dense_rank(v):
    return np.unique(v, return_inverse=True)[1].astype(np.int64)


def _tilings(embedding, per_col, S):
    """Candidate (columns, strip coordinate) pairs: the BFS embedding as it is
    and rotated by 45 degrees (on a square whose pseudo-peripheral pair is a
    diagonal the raw embedding is a diamond: twice the columns, half the rows
    per column).  For each, the smallest number of equal-width strips whose
    fullest (strip, column) cell leaves room for the halo; the cheapest
    (fewest macro-steps x window rows) comes first."""
    ex, ey = embedding
    rx, ry = _dense_rank(ex), _dense_rank(ey)
    cands = [(rx, ry), (_dense_rank(rx + ry), _dense_rank(ry - rx)),
             (ry, rx), (_dense_rank(ry - rx), _dense_rank(rx + ry))]
    out = []
    target = max(4.0, per_col - S)
    for col, v in cands:
        ncols, vmax = int(col.max()) + 1, int(v.max()) + 1
        nstrips = 1
        while True:
            sid = (v * nstrips) // vmax
            cell = np.bincount(sid * ncols + col, minlength=nstrips * ncols)
            cell = cell.reshape(nstrips, ncols)
            if cell.max() <= target or nstrips >= vmax:
                break
            nstrips = max(nstrips + 1, int(nstrips * cell.max() / target))
        cost = float(((cell > 0).sum(axis=1) *
                      (cell.max(axis=1) + S)).sum())
        out.append((cost, len(out), col, v, nstrips))
    out.sort(key=lambda c: c[:2])
    return out


def compile_program(indptr, indices, wave, nsweeps, backward, capacity,
                    embedding=None, kind_of_row=None, chunks=33, sms=148,
                    max_redundancy=1.7, ngrp=128, tiling=None, canon=None,
                    verbose=False):
    """Program for `nsweeps` sweeps (forward or backward) of the level whose
    sparsity pattern is (indptr, indices) and whose rows have wavefront numbers
    `wave`.  `capacity` = window slots available to a CTA, `ngrp` = row updates
    a CTA runs side by side (threads / lanes per row), `chunks` = time chunks
    per item the launch is expected to have (items x chunks CTAs share `sms`
    SMs: the tiling is chosen for the shortest makespan; `tiling` = (strips,
    segments) of an earlier program of the same level skips the search).  Returns None if the
    level does not tile well (deep wavefront DAG, 3-D-like connectivity,
    disconnected graph): the caller keeps the per-wavefront kernels then."""
    n = len(indptr) - 1
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    wave = np.ascontiguousarray(wave, dtype=np.int32)
    D = int(wave.max()) + 1 if n else 0
    S = D * nsweeps
    if n == 0 or S == 0 or S > MAX_STAGES or capacity > 65535:
        return None
    nnz_row = np.diff(indptr)
    if nnz_row.max() > 64:
        return None
    rows_all = np.repeat(np.arange(n, dtype=np.int64), nnz_row)
    if np.count_nonzero(indices == rows_all) != n:
        return None  # the kernel takes u_i from the row's diagonal entry
    if canon is None:  # diagonal first, then CSR order
        canon = np.lexsort((indices, indices != rows_all, rows_all))
    canon = np.asarray(canon, dtype=np.int64)
    canon_cols = indices[canon]  # columns of the rows' entries, program order
    assert np.array_equal(canon_cols[indptr[:-1]], np.arange(n))
    if embedding is None:
        embedding = graph_embedding(indptr, indices, n)
        if embedding is None:
            return None
    colours = stage_colours(D, nsweeps, backward)
    scratch = _Scratch(n)
    per_col = capacity / (LAG * S + LOOKAHEAD + 3.0)
    if n <= capacity // 2:
        tilings = [(0.0, 0, _dense_rank(embedding[0]),
                    _dense_rank(embedding[1]), 1)]
    else:
        tilings = _tilings(embedding, per_col, S)
    _cost, _k, colidx, vcoord, nstrips = tilings[0]
    ncols, vmax = int(colidx.max()) + 1, int(vcoord.max()) + 1

    def items_for(nstrips, nseg):
        sid = (vcoord * nstrips) // vmax
        order = np.lexsort((colidx, sid))
        bounds = np.searchsorted(sid[order], np.arange(nstrips + 1))
        out = []
        for s in range(nstrips):
            rows = order[bounds[s]:bounds[s + 1]]
            for g in range(nseg):
                part = rows[g * len(rows) // nseg:(g + 1) * len(rows) // nseg]
                if len(part):
                    out.append(np.sort(part).astype(np.int64))
        return out

    max_seg = max(1, ncols // (4 * LAG * S + 1))  # shorter segments are all fill and drain

    def probe_tiling(nstrips, nseg):
        """Schedule only the largest item: (makespan estimate, items) or None
        if it does not fit the window."""
        its = items_for(nstrips, nseg)
        res = _schedule_item(indptr, indices, wave, colours, max(its, key=len),
                             colidx, scratch)
        if res['nslots'] > capacity:
            return None
        per_step = (np.bincount(res['op_step']) + ngrp - 1) // ngrp
        if per_step.max() > MAX_PASSES:
            return None
        # CTA rounds x passes of the longest item: what one launch costs when
        # every SM holds one CTA (items x chunks CTAs on `sms` SMs)
        longest = int(per_step.sum()) + res['nsteps'] // 2
        return -(-len(its) * chunks // sms) * longest, its

    def full_tiling(its):
        sch = [
            _schedule_item(indptr, indices, wave, colours, own, colidx,
                           scratch) for own in its
        ]
        if max(r['nslots'] for r in sch) > capacity:
            return None
        return sch

    sched = None
    if tiling is not None:
        sched = full_tiling(items_for(*tiling))
        nstrips = tiling[0]
    else:
        first = None
        for _attempt in range(60):
            first = probe_tiling(nstrips, 1)
            if first is not None or nstrips >= vmax:
                break
            nstrips += 1 + nstrips // 16
        if first is None:
            return None
        # more strips / segments: fuller passes of `ngrp` ops, fewer idle SMs
        # in the last round of CTAs.  A CTA's run time is its number of steps
        # and passes, whatever the slab width: on narrow time slabs (few
        # chunks, e.g. 33 slices per GPU) and on the small levels the SMs are
        # filled by cutting the level into more items -- shorter pipelines at
        # the price of recomputed fill columns -- as long as all CTAs still
        # run in about one round.
        cands = [(first[0], nstrips, 1, first[1])]
        strip_opts = sorted(set(
            list(range(nstrips, nstrips + (4 if nstrips > 1 else 1))) +
            [min(vmax, nstrips * k) for k in (2, 3, 4, 6, 8)
             if nstrips * k * chunks <= 2 * sms]))
        seg_opts = [g for g in (1, 2, 3, 4, 5, 6, 7, 8, 10) if g <= max(1, ncols // 6)]
        probes = 0
        for ns in strip_opts:
            for nseg in seg_opts:
                if (ns, nseg) == (nstrips, 1):
                    continue
                ctas = ns * nseg * chunks
                if ns > nstrips + 3 and ctas > 2 * sms:
                    break  # extra strips only to fill idle SMs
                if ctas > 12 * sms and nseg > 1:
                    break
                if probes >= 64:
                    break
                probes += 1
                got = probe_tiling(ns, nseg)
                if got is not None:
                    cands.append((got[0], ns, nseg, got[1]))
        cands.sort(key=lambda c: c[:3])
        for _cost, ns, nseg, its in cands[:16]:
            sched = full_tiling(its)
            if sched is not None:
                nops_c = sum(len(r['op_row']) for r in sched)
                if (nops_c / float(n * nsweeps) > max_redundancy
                        and (ns, nseg) != (nstrips, 1)):
                    sched = None  # too much recomputation: next candidate
                    continue
                tiling = (ns, nseg)
                nstrips = ns
                break
        if sched is None:  # the plain tiling, whatever its redundancy
            sched = full_tiling(first[1])
            tiling = (nstrips, 1)
    if sched is None:
        return None
    nops = sum(len(r['op_row']) for r in sched)
    redundancy = nops / float(n * nsweeps)
    if redundancy > max_redundancy:
        if verbose:
            print('gs_program: redundancy %.2f too high' % redundancy)
        return None

    # ---- flatten: static pass layout ----
    # A CTA runs `ngrp` row updates side by side; the ops of a macro-step are
    # laid out as ceil(count / ngrp) passes of exactly ngrp records (padded
    # with no-ops, nnz = 0), so that thread group g executes records
    # pass * ngrp + g for consecutive passes: the record and the right-hand
    # side of the op PREFETCH passes ahead have addresses known in advance.
    prog = GSProgram()
    maxnnz = int(nnz_row.max())
    recw = 8 + 4 * max(0, (maxnnz - 8 + 7) // 8)
    item_step, item_pass = [0], [0]
    step_pass, step_ld = [], []   # END offsets per step
    recs, ld_rows, ld_slots = [], [], []
    npass = nld = 0
    for r in sched:
        region, slot = r['region'], r['slot']
        nst = r['nsteps']
        cnt = np.bincount(r['op_step'], minlength=nst)
        ptr_loc = np.concatenate([[0], np.cumsum(cnt)])
        passes = (cnt + ngrp - 1) // ngrp
        pass_loc = np.concatenate([[0], np.cumsum(passes)])
        step_pass.extend((npass + pass_loc[1:]).tolist())
        cntl = np.bincount(r['ld_step'], minlength=nst)
        step_ld.extend((nld + np.cumsum(cntl)).tolist())
        item_step.append(item_step[-1] + nst)
        grow = region[r['op_row']]
        nop = len(grow)
        rnnz = nnz_row[grow].astype(np.int64)
        pos, seg = _expand(indptr, grow)
        nb_slot = slot[np.searchsorted(region, canon_cols[pos])]
        assert (nb_slot >= 0).all()
        tot = int(pass_loc[-1]) * ngrp
        rec = np.zeros((tot, recw), dtype=np.uint32)
        j = np.arange(nop, dtype=np.int64) - ptr_loc[r['op_step']]
        dst = pass_loc[r['op_step']] * ngrp + j   # record index inside the item
        rec[dst, 0] = grow.astype(np.uint32) | (r['op_store'].astype(np.uint32) << 31)
        if kind_of_row is not None:
            rec[dst, 2] = kind_of_row[grow].astype(np.uint32)
        else:
            rec[dst, 2] = indptr[grow].astype(np.uint32)
        rec[dst, 3] = slot[r['op_row']].astype(np.uint32) | (rnnz.astype(np.uint32) << 16)
        slots16 = rec[:, 4:].view(np.uint16)  # (tot, 2 * (recw - 4))
        within = np.arange(len(pos), dtype=np.int64) - np.repeat(seg[:-1], rnnz)
        # entries in the program's (canonical) order, the diagonal first; unused
        # entries name the row's own slot (their table value is 0)
        slots16[dst] = slot[r['op_row']].astype(np.uint16)[:, None]
        slots16[np.repeat(dst, rnnz), within] = nb_slot.astype(np.uint16)
        # f prefetch hint: the row of the record PREFETCH passes later
        ahead = np.arange(tot, dtype=np.int64) + PREFETCH * ngrp
        rec[:, 1] = rec[np.minimum(ahead, tot - 1), 0] & np.uint32(0x7fffffff)
        rec[ahead >= tot, 1] = 0
        if passes.max(initial=0) > MAX_PASSES:
            return None
        # device layout: per pass, all headers, then all slot words (SoA), so
        # that one bulk copy brings a pass and the reads are conflict-free
        recs.append(rec.reshape(-1, ngrp, recw // 4, 4).transpose(0, 2, 1, 3)
                    .reshape(-1, recw))
        ld_rows.append(region[r['ld_row']].astype(np.int32))
        ld_slots.append(slot[r['ld_row']].astype(np.uint16))
        npass += int(pass_loc[-1])
        nld += len(r['ld_row'])
        item_pass.append(npass)
    prog.item_step = np.asarray(item_step, dtype=np.int32)
    prog.item_pass = np.asarray(item_pass, dtype=np.int32)
    # per step: (end pass, end load), both global; one int2 per step
    prog.step_info = np.ascontiguousarray(
        np.stack([np.asarray(step_pass, dtype=np.int32),
                  np.asarray(step_ld, dtype=np.int32)], axis=1))
    prog.op = np.ascontiguousarray(np.concatenate(recs))
    prog.ld = np.ascontiguousarray(
        np.stack([np.concatenate(ld_rows).astype(np.int32),
                  np.concatenate(ld_slots).astype(np.int32)], axis=1))
    prog.nitems = len(sched)
    prog.nslots = max(r['nslots'] for r in sched)
    prog.nrows = n
    prog.nsweeps, prog.backward, prog.D = nsweeps, bool(backward), D
    prog.maxnnz, prog.recw, prog.ngrp = maxnnz, recw, ngrp
    prog.generic = kind_of_row is None
    prog.tiling = tiling
    prog.canon = canon
    prog.stats = {
        'rows': n, 'items': len(sched), 'stages': S,
        'columns': ncols, 'ops': int(nops), 'redundancy': redundancy,
        'records': int(len(prog.op)),
        'padding': len(prog.op) / float(nops),
        'slots': prog.nslots,
        'steps_max': int(max(r['nsteps'] for r in sched)),
        'steps_total': int(item_step[-1]),
        'passes_total': int(npass),
        'loads': int(len(prog.ld)),
        'load_redundancy': len(prog.ld) / float(n),
        'bytes': int(prog.op.nbytes + prog.ld.nbytes + prog.step_info.nbytes),
    }
    return prog


def wavefronts(indptr, indices):
    """wave[i] of the lexicographic sweep's dependency DAG and the number of
    wavefronts (stk_gs_wavefronts, host)."""
    n = len(indptr) - 1
    indptr = np.ascontiguousarray(indptr, dtype=np.int32)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    wave = np.zeros(max(n, 1), dtype=np.int32)
    depth = lib().stk_gs_wavefronts(n, indptr.ctypes.data, indices.ctypes.data,
                                    wave.ctypes.data)
    return wave[:n], int(depth)
