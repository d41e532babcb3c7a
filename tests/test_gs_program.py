"""CPU tests of the fused Gauss-Seidel program compiler
(spacetime_fullgrid_parallel_b200/gs_program.py): the program, interpreted by
tests/gs_emulator.py with the device kernel's semantics (window slots, load
latency, concurrent ops inside a macro-step), must reproduce the sequential
lexicographic sweep of the oracle (oracle/gs.c, multigrid.py:89-97) whatever
the tiling."""
import numpy as np
import pytest
import scipy.sparse as sp

from gs_emulator import emulate
from spacetime_fullgrid_parallel_b200 import gs_program as gp


def _seq(A, f, u0, nsw, backward):
    from oracle import cgs
    u = np.ascontiguousarray(u0.T.copy())
    cgs.gauss_seidel(A.indptr.astype(np.int32), A.indices.astype(np.int32),
                     A.data, np.ascontiguousarray(f.T), u, nsw,
                     backward=backward)
    return u.T


def _square_level(Js):
    from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
    prob = SquareProblem(Js, 1)
    A = sp.csr_matrix(prob.A_x + 3.0 * prob.M_x)
    A.sort_indices()
    return A


def _check(A, nsw, backward, capacity, min_items, zero, kinds=False):
    n = A.shape[0]
    wave, depth = gp.wavefronts(A.indptr, A.indices)
    kind, canon = None, None
    if kinds:
        canon = gp.canonical_order(A.indptr, A.indices, [A.data])
        kind = gp.row_kinds(A.indptr, [A.data], canon)
        assert kind is not None
        kind = kind[0]
    prog = gp.compile_program(A.indptr, A.indices, wave, nsw, backward,
                              capacity, chunks=1, sms=min_items,
                              kind_of_row=kind, canon=canon,
                              max_redundancy=50.0)
    assert prog is not None
    assert prog.nslots <= capacity
    rng = np.random.RandomState(3)
    f = rng.rand(n, 2)
    u0 = np.zeros((n, 2)) if zero else rng.rand(n, 2)
    out = emulate(prog, A.indptr, A.data, A.diagonal(), f,
                  None if zero else u0, indices=A.indices)
    ref = _seq(A, f, u0, nsw, backward)
    assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max()
    return prog


@pytest.mark.parametrize('backward', [False, True])
@pytest.mark.parametrize('Js,capacity,min_items', [(3, 400, 1), (5, 700, 1),
                                                   (5, 900, 6), (6, 1500, 9)])
def test_program_equals_sequential_sweeps(Js, capacity, min_items, backward):
    A = _square_level(Js)
    prog = _check(A, 3, backward, capacity, min_items, zero=not backward,
                  kinds=(Js == 5))
    if Js >= 5:
        assert prog.nitems > 1
    _check(A, 1, backward, capacity, 1, zero=False)


def test_row_kinds_uniform_mesh():
    """A uniformly refined mesh has a handful of distinct stencil rows, found
    exactly (bit for bit); a perturbed matrix has none to share."""
    A = _square_level(5)
    canon = gp.canonical_order(A.indptr, A.indices, [A.data])
    assert np.array_equal(A.indices[canon][A.indptr[:-1]],
                          np.arange(A.shape[0]))  # the diagonal first
    kind, rep = gp.row_kinds(A.indptr, [A.data], canon)
    # entries sorted by value: all interior rows are ONE kind, whatever the
    # numbering of their neighbours; the others are boundary variants
    assert len(rep) <= 16
    assert np.bincount(kind).max() > 0.9 * A.shape[0]
    nnz = np.diff(A.indptr)
    cv = A.data[canon]
    for k, r in enumerate(rep):
        rows = np.nonzero(kind == k)[0]
        assert (nnz[rows] == nnz[r]).all()
        for i in rows[:5]:
            assert np.array_equal(cv[A.indptr[i]:A.indptr[i + 1]],
                                  cv[A.indptr[r]:A.indptr[r + 1]])
    B = A.copy()
    B.data = B.data * (1 + 1e-9 * np.random.RandomState(0).rand(B.nnz))
    assert gp.row_kinds(B.indptr, [B.data], canon) is None


def test_unstructured_mesh_generic_path():
    """No geometry, no stencil: a Delaunay mesh of random points, numbered by
    greedy colour class (a multicolour Gauss-Seidel).  The BFS embedding tiles
    it, values come from the CSR arrays."""
    from scipy.spatial import Delaunay
    rng = np.random.RandomState(11)
    pts = rng.rand(3000, 2)
    tri = Delaunay(pts).simplices
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])
    n = len(pts)
    G = sp.coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(n, n))
    G = ((G + G.T) > 0).astype(np.float64).tocsr()
    colour = np.full(n, -1)
    for i in range(n):
        used = set(colour[G.indices[G.indptr[i]:G.indptr[i + 1]]])
        c = 0
        while c in used:
            c += 1
        colour[i] = c
    perm = np.argsort(colour, kind='stable')
    G = G[perm][:, perm]
    L = sp.diags(np.asarray(G.sum(axis=1)).ravel() + 1.0) - 0.5 * G
    L = sp.csr_matrix(L)
    L.sort_indices()
    wave, depth = gp.wavefronts(L.indptr, L.indices)
    assert depth <= 8
    prog = _check(L, 2, False, 900, 4, zero=False)
    assert prog.generic and prog.nitems >= 2
    _check(L, 2, True, 900, 1, zero=True)


def test_deep_wavefront_numbering_is_rejected():
    """Lexicographic numbering: O(sqrt n) wavefronts, no bounded dependency
    closure; the caller keeps the per-wavefront kernels."""
    from spacetime_fullgrid_parallel_b200.assembly import SquareProblem
    prob = SquareProblem(4, 1, order='lex')
    A = sp.csr_matrix(prob.A_x)
    A.sort_indices()
    wave, depth = gp.wavefronts(A.indptr, A.indices)
    assert depth > 8
    assert gp.compile_program(A.indptr, A.indices, wave, 3, False,
                              3000) is None


def test_narrow_slabs_get_more_items():
    """A CTA's run time is its number of steps and passes whatever the slab
    width, so with few time chunks per item the compiler cuts the level into
    more items (shorter pipelines) until one round of CTAs fills the SMs -- and
    the program still reproduces the sequential sweeps."""
    A = _square_level(6)
    n = A.shape[0]
    wave, _ = gp.wavefronts(A.indptr, A.indices)
    progs = {}
    for chunks in (33, 5):
        progs[chunks] = gp.compile_program(A.indptr, A.indices, wave, 3, False,
                                           2300, chunks=chunks, sms=148)
        assert progs[chunks] is not None

    def longest(pg):
        return int((np.diff(pg.item_pass) + np.diff(pg.item_step) // 2).max())

    assert progs[5].nitems > progs[33].nitems
    assert longest(progs[5]) < longest(progs[33])
    assert progs[5].stats['redundancy'] <= 1.7
    rng = np.random.RandomState(4)
    f, u0 = rng.rand(n, 2), rng.rand(n, 2)
    out = emulate(progs[5], A.indptr, A.data, A.diagonal(), f, u0,
                  indices=A.indices)
    ref = _seq(A, f, u0, 3, False)
    assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max()
