"""Host-side assembly of the synthetic `square` heat-equation problem.

In the reference this is NGSolve/Netgen's job, done once on the host
(heateq_mpi.py:63-104, source/mesh.py:4-30, source/problem.py:7-18,
source/ngsolve_helper.py:38-46, source/multigrid.py:19-60).  NGSolve is not
available in this image, so the same objects are produced here with vectorised
numpy/scipy: P1 mass/stiffness matrices restricted to the free (interior)
dofs, the 1-D P1 time matrices, the load vectors and the prolongation
matrices of the uniformly refined mesh hierarchy.  Everything downstream (the
operator graph, the solve) consumes only these CSR matrices and vectors, as
the reference does.

The unit square is split into 2 triangles, refined once (mesh.py:21-30:
`ngmesh.Refine()`), then `J_space` more times; all refinements are uniform
(red), so level j is the (2^(j+1)) x (2^(j+1)) Friedrichs-Keller triangulation
with (2^(j+1)-1)^2 interior vertices.  Vertices are numbered hierarchically
(coarse vertices keep their numbers, new vertices are appended), which is what
`MeshHierarchy` relies on (multigrid.py:26-35).

The order in which the *new* vertices of a level are appended is the one
degree of freedom NGSolve would fix and we cannot observe; `order` selects it:
  'class'  edge-midpoint classes one after the other (horizontal, vertical,
           diagonal edges), lexicographic inside a class.  Lexicographic
           Gauss-Seidel is then exactly a <=4-wavefront sweep on every level.
  'lex'    all new vertices lexicographic by (y, x): O(2^j) wavefronts.
  'random' a seeded random order (stress test of the general wavefront path).
"""
import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------
# time direction: 1-D P1 elements on [0, T] with 2^J_time elements
# ----------------------------------------------------------------------------
def _assemble_1d(n_el, T, local, scale):
    """Sum `scale * local` (2x2) over the n_el intervals."""
    n = n_el + 1
    k = np.arange(n_el)
    rows = np.stack([k, k, k + 1, k + 1], axis=1).reshape(-1)
    cols = np.stack([k, k + 1, k, k + 1], axis=1).reshape(-1)
    vals = np.tile(np.asarray(local, dtype=np.float64).reshape(-1) * scale,
                   n_el)
    mat = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    mat.eliminate_zeros()  # ngsolve_helper.py:44
    mat.sort_indices()
    return mat


def time_matrices(J_time, T=1.0):
    """A_t, L_t, M_t, G_t and u0_t of heateq_mpi.py:78-88,102.

    Row index = test function, column index = trial function, as in
    `BilForm` (ngsolve_helper.py:21-36).  No Dirichlet condition in time.
    """
    n_el = 2**J_time
    h = T / n_el
    A_t = _assemble_1d(n_el, T, [[1, -1], [-1, 1]], 1.0 / h)
    M_t = _assemble_1d(n_el, T, [[2, 1], [1, 2]], h / 6.0)
    # L[i, j] = int phi_j * phi_i'  (u * grad(v) * dx, u trial, v test)
    L_t = _assemble_1d(n_el, T, [[-1, -1], [1, 1]], 0.5)
    N = n_el + 1
    G_t = sp.csr_matrix(([1.0], ([0], [0])), shape=(N, N))
    u0_t = np.zeros(N)
    u0_t[0] = 1.0
    return A_t, L_t, M_t, G_t, u0_t


# ----------------------------------------------------------------------------
# space direction: uniformly refined triangulation of the unit square
# ----------------------------------------------------------------------------
def square_dof_numbering(J_space, order='class', seed=0):
    """Hierarchical numbering of the interior vertices of level J_space.

    Returns (dof, nverts): `dof` is an (n+1, n+1) int64 array indexed [y, x] in
    finest-grid coordinates (n = 2^(J_space+1)), -1 on the boundary;
    nverts[j] is the number of interior vertices of level j.
    """
    n = 2**(J_space + 1)
    dof = np.full((n + 1, n + 1), -1, dtype=np.int64)
    nverts = []
    counter = 0
    rng = np.random.RandomState(seed)
    for j in range(J_space + 1):
        s = 2**(J_space - j)  # fine-grid spacing of level j
        nj = 2**(j + 1)
        Y, X = np.meshgrid(np.arange(1, nj), np.arange(1, nj), indexing='ij')
        Y, X = Y.reshape(-1), X.reshape(-1)  # lexicographic by (y, x)
        if j == 0:
            groups = [np.ones(len(X), dtype=bool)]
        else:
            px, py = X % 2, Y % 2
            new = (px + py) > 0
            if order == 'class':
                groups = [(px == 1) & (py == 0), (px == 0) & (py == 1),
                          (px == 1) & (py == 1)]
            else:
                groups = [new]
        for g in groups:
            idx = np.nonzero(g)[0]
            if order == 'random':
                idx = rng.permutation(idx)
            dof[Y[idx] * s, X[idx] * s] = counter + np.arange(len(idx))
            counter += len(idx)
        nverts.append(counter)
    assert counter == (n - 1)**2
    return dof, nverts


def _assemble_p1(nj, vertex_dof, n_dofs, kind):
    """P1 mass ('mass') or stiffness ('stiff') matrix on the level with nj
    cells per side, restricted to the free dofs (ngsolve_helper.py:38-46).

    Generic element-by-element assembly (vectorised over triangles), not a
    hard-coded stencil: the 5-point stencil of the stiffness matrix appears
    because the hypotenuse couplings are exact zeros and are eliminated.
    """
    h = 1.0 / nj
    cy, cx = np.meshgrid(np.arange(nj), np.arange(nj), indexing='ij')
    cy, cx = cy.reshape(-1), cx.reshape(-1)
    # two triangles per cell, diagonal (x,y)-(x+1,y+1)
    tx = np.concatenate([
        np.stack([cx, cx + 1, cx + 1], axis=1),
        np.stack([cx, cx + 1, cx], axis=1)
    ])
    ty = np.concatenate([
        np.stack([cy, cy, cy + 1], axis=1),
        np.stack([cy, cy + 1, cy + 1], axis=1)
    ])
    px, py = tx * h, ty * h
    det = ((px[:, 1] - px[:, 0]) * (py[:, 2] - py[:, 0]) -
           (px[:, 2] - px[:, 0]) * (py[:, 1] - py[:, 0]))
    area = 0.5 * np.abs(det)
    if kind == 'mass':
        loc = (area / 12.0)[:, None, None] * (np.ones((3, 3)) + np.eye(3))
    else:
        # grad phi_a = (y_b - y_c, x_c - x_b) / det for (a, b, c) cyclic
        gx = np.stack([py[:, 1] - py[:, 2], py[:, 2] - py[:, 0],
                       py[:, 0] - py[:, 1]], axis=1) / det[:, None]
        gy = np.stack([px[:, 2] - px[:, 1], px[:, 0] - px[:, 2],
                       px[:, 1] - px[:, 0]], axis=1) / det[:, None]
        loc = area[:, None, None] * (gx[:, :, None] * gx[:, None, :] +
                                     gy[:, :, None] * gy[:, None, :])
    d = vertex_dof[ty, tx]  # (ntri, 3) dof ids, -1 on the boundary
    rows = np.repeat(d[:, :, None], 3, axis=2).reshape(-1)
    cols = np.repeat(d[:, None, :], 3, axis=1).reshape(-1)
    vals = loc.reshape(-1)
    keep = (rows >= 0) & (cols >= 0)
    mat = sp.coo_matrix((vals[keep], (rows[keep], cols[keep])),
                        shape=(n_dofs, n_dofs)).tocsr()
    mat.eliminate_zeros()
    mat.sort_indices()
    return mat


def _prolongation(dof_f, nf_dofs, nc_dofs, njf):
    """P: level j -> j+1 (multigrid.py:39-59): identity on the coarse
    vertices, 1/2-1/2 from the two parent vertices for each new vertex,
    restricted to free dofs.  `dof_f` is the dof array in level-(j+1) grid
    coordinates ((njf+1) x (njf+1))."""
    Y, X = np.meshgrid(np.arange(1, njf), np.arange(1, njf), indexing='ij')
    Y, X = Y.reshape(-1), X.reshape(-1)
    px, py = X % 2, Y % 2
    me = dof_f[Y, X]
    old = (px == 0) & (py == 0)
    rows = [me[old]]
    cols = [me[old]]  # hierarchical numbering: same id on the coarse level
    vals = [np.ones(old.sum())]
    # parents of an edge midpoint = the two endpoints of its edge
    for sel, dx, dy in [((px == 1) & (py == 0), 1, 0),
                        ((px == 0) & (py == 1), 0, 1),
                        ((px == 1) & (py == 1), 1, 1)]:
        for sgn in (-1, 1):
            par = dof_f[Y[sel] + sgn * dy, X[sel] + sgn * dx]
            ok = par >= 0  # boundary parents are not dofs (multigrid.py:57-59)
            rows.append(me[sel][ok])
            cols.append(par[ok])
            vals.append(np.full(ok.sum(), 0.5))
    rows, cols, vals = map(np.concatenate, (rows, cols, vals))
    assert cols.max() < nc_dofs
    P = sp.coo_matrix((vals, (rows, cols)), shape=(nf_dofs, nc_dofs)).tocsr()
    P.sort_indices()
    return P


class SquareMeshHierarchy:
    """Duck-typed stand-in for `MeshHierarchy` (multigrid.py:14-80): carries
    exactly the attributes `MultiGrid.__init__` consumes."""
    def __init__(self, J_space, order='class', seed=0):
        self.J = J_space
        self.order = order
        self.dof, self.nverts = square_dof_numbering(J_space, order, seed)
        self.shared_comm = None
        self.P_mats = []
        for j in range(J_space):
            s = 2**(J_space - (j + 1))
            self.P_mats.append(
                _prolongation(self.dof[::s, ::s], self.nverts[j + 1],
                              self.nverts[j], 2**(j + 2)))
        self.R_mats = [P.T.tocsr() for P in self.P_mats]
        for R in self.R_mats:
            R.sort_indices()

    def level_dof(self, j):
        """dof array of level j in level-j grid coordinates."""
        s = 2**(self.J - j)
        return self.dof[::s, ::s]

    def assemble(self, kind, j=None):
        j = self.J if j is None else j
        return _assemble_p1(2**(j + 1), self.level_dof(j), self.nverts[j],
                            kind)

    def nodal_values(self, fn):
        """fn(x, y) at the interior vertices, in dof order."""
        n = 2**(self.J + 1)
        Y, X = np.nonzero(self.dof >= 0)
        out = np.empty(self.nverts[-1])
        out[self.dof[Y, X]] = fn(X / n, Y / n)
        return out


# ----------------------------------------------------------------------------
# space direction, 3-D: uniformly refined Kuhn (Freudenthal) triangulation of
# the unit cube (problem='cube', problem.py:21-32, mesh.py:33-43)
# ----------------------------------------------------------------------------
_KUHN_PERMS = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
# parity classes of the new vertices of a level = directions of the Kuhn edges
# whose midpoints they are; (dx, dy, dz)
_CUBE_CLASSES = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1),
                 (0, 1, 1), (1, 1, 1)]


def cube_dof_numbering(J_space, order='class', seed=0):
    """Hierarchical numbering of the interior vertices of the level-J_space
    grid of the unit cube; `dof` is indexed [z, y, x] in finest-grid
    coordinates, -1 on the boundary (3-D analogue of square_dof_numbering)."""
    n = 2**(J_space + 1)
    dof = np.full((n + 1, n + 1, n + 1), -1, dtype=np.int64)
    nverts, counter = [], 0
    rng = np.random.RandomState(seed)
    for j in range(J_space + 1):
        s = 2**(J_space - j)
        nj = 2**(j + 1)
        r = np.arange(1, nj)
        Z, Y, X = [a.reshape(-1) for a in np.meshgrid(r, r, r, indexing='ij')]
        if j == 0:
            groups = [np.ones(len(X), dtype=bool)]
        else:
            px, py, pz = X % 2, Y % 2, Z % 2
            if order == 'class':
                groups = [(px == a) & (py == b) & (pz == c)
                          for a, b, c in _CUBE_CLASSES]
            else:
                groups = [(px + py + pz) > 0]
        for g in groups:
            idx = np.nonzero(g)[0]
            if order == 'random':
                idx = rng.permutation(idx)
            dof[Z[idx] * s, Y[idx] * s, X[idx] * s] = counter + np.arange(
                len(idx))
            counter += len(idx)
        nverts.append(counter)
    assert counter == (n - 1)**3
    return dof, nverts


def _assemble_p1_3d(nj, vertex_dof, n_dofs, kind):
    """P1 mass / stiffness matrix on the Kuhn triangulation with nj cells per
    side (six tetrahedra per cell, all sharing the cell's main diagonal),
    restricted to the free dofs; generic element assembly vectorised over the
    tetrahedra (ngsolve_helper.py:38-46)."""
    h = 1.0 / nj
    r = np.arange(nj)
    cz, cy, cx = [a.reshape(-1) for a in np.meshgrid(r, r, r, indexing='ij')]
    base = np.stack([cx, cy, cz], axis=1)  # (ncell, 3) as (x, y, z)
    eye = np.eye(3, dtype=np.int64)
    verts = []  # (ntet, 4, 3) integer vertex coordinates
    for perm in _KUHN_PERMS:
        v = [base]
        for k in perm:
            v.append(v[-1] + eye[k])
        verts.append(np.stack(v, axis=1))
    verts = np.concatenate(verts, axis=0)
    P = verts * h
    E = P[:, 1:, :] - P[:, :1, :]  # edge matrix rows = v_k - v_0
    det = np.linalg.det(E)
    vol = np.abs(det) / 6.0
    if kind == 'mass':
        loc = (vol / 20.0)[:, None, None] * (np.ones((4, 4)) + np.eye(4))
    else:
        # gradients of the barycentric functions: rows of inv(E) for 1..3,
        # minus their sum for 0
        Einv = np.linalg.inv(E)  # (ntet, 3, 3): columns = grad phi_k, k=1..3
        g = np.transpose(Einv, (0, 2, 1))  # (ntet, 3 funcs, 3 comps)
        g0 = -g.sum(axis=1, keepdims=True)
        G = np.concatenate([g0, g], axis=1)  # (ntet, 4, 3)
        loc = vol[:, None, None] * np.einsum('tac,tbc->tab', G, G)
    d = vertex_dof[verts[:, :, 2], verts[:, :, 1], verts[:, :, 0]]  # (ntet, 4)
    rows = np.repeat(d[:, :, None], 4, axis=2).reshape(-1)
    cols = np.repeat(d[:, None, :], 4, axis=1).reshape(-1)
    vals = loc.reshape(-1)
    keep = (rows >= 0) & (cols >= 0)
    mat = sp.coo_matrix((vals[keep], (rows[keep], cols[keep])),
                        shape=(n_dofs, n_dofs)).tocsr()
    # exact zeros of the stiffness matrix (orthogonal gradients) appear as
    # round-off of size 1e-17: drop them as NGSolve's eliminate_zeros would
    mat.data[np.abs(mat.data) < 1e-13 * np.abs(mat.data).max()] = 0.0
    mat.eliminate_zeros()
    mat.sort_indices()
    return mat


def _prolongation_3d(dof_f, nf_dofs, nc_dofs, njf):
    """P: level j -> j+1 on the cube (multigrid.py:39-59): new vertices are
    midpoints of Kuhn edges, their parents the two end points."""
    r = np.arange(1, njf)
    Z, Y, X = [a.reshape(-1) for a in np.meshgrid(r, r, r, indexing='ij')]
    px, py, pz = X % 2, Y % 2, Z % 2
    me = dof_f[Z, Y, X]
    old = (px + py + pz) == 0
    rows, cols, vals = [me[old]], [me[old]], [np.ones(old.sum())]
    for dx, dy, dz in _CUBE_CLASSES:
        sel = (px == dx) & (py == dy) & (pz == dz)
        for sgn in (-1, 1):
            par = dof_f[Z[sel] + sgn * dz, Y[sel] + sgn * dy, X[sel] + sgn * dx]
            ok = par >= 0
            rows.append(me[sel][ok])
            cols.append(par[ok])
            vals.append(np.full(ok.sum(), 0.5))
    rows, cols, vals = map(np.concatenate, (rows, cols, vals))
    assert cols.max() < nc_dofs
    P = sp.coo_matrix((vals, (rows, cols)), shape=(nf_dofs, nc_dofs)).tocsr()
    P.sort_indices()
    return P


class CubeMeshHierarchy:
    """3-D counterpart of SquareMeshHierarchy: level j is the Kuhn
    triangulation of the 2^(j+1) x 2^(j+1) x 2^(j+1) grid, (2^(j+1)-1)^3
    interior vertices, hierarchical numbering.  With order='class' the eight
    parity classes of a level are independent sets of the 15-point mass-matrix
    graph, i.e. lexicographic Gauss-Seidel is an 8-wavefront sweep."""
    def __init__(self, J_space, order='class', seed=0):
        self.J = J_space
        self.order = order
        self.dof, self.nverts = cube_dof_numbering(J_space, order, seed)
        self.shared_comm = None
        self.P_mats = []
        for j in range(J_space):
            s = 2**(J_space - (j + 1))
            self.P_mats.append(
                _prolongation_3d(self.dof[::s, ::s, ::s], self.nverts[j + 1],
                                 self.nverts[j], 2**(j + 2)))
        self.R_mats = [P.T.tocsr() for P in self.P_mats]
        for R in self.R_mats:
            R.sort_indices()

    def level_dof(self, j):
        s = 2**(self.J - j)
        return self.dof[::s, ::s, ::s]

    def assemble(self, kind, j=None):
        j = self.J if j is None else j
        return _assemble_p1_3d(2**(j + 1), self.level_dof(j), self.nverts[j],
                               kind)

    def nodal_values(self, fn):
        n = 2**(self.J + 1)
        Z, Y, X = np.nonzero(self.dof >= 0)
        out = np.empty(self.nverts[-1])
        out[self.dof[Z, Y, X]] = fn(X / n, Y / n, Z / n)
        return out


class SquareProblem:
    """Everything heateq_mpi.py:63-104 obtains from NGSolve for
    problem='square' (problem.py:7-18): u(t,x,y)=exp(-2pi^2 t)sin(pi x)sin(pi y).
    """
    dim = 2

    def _mesh(self, J_space, order, seed):
        return SquareMeshHierarchy(J_space, order, seed)

    def _u0(self):
        return self.hierarchy.nodal_values(
            lambda x, y: np.sin(np.pi * x) * np.sin(np.pi * y))

    def __init__(self, J_space, J_time=None, alpha=0.3, order='class', seed=0):
        if J_time is None:
            J_time = J_space
        self.J_space, self.J_time, self.alpha = J_space, J_time, alpha
        self.hierarchy = self._mesh(J_space, order, seed)
        self.A_t, self.L_t, self.M_t, self.G_t, self.u0_t = time_matrices(
            J_time)
        self.M_x = self.hierarchy.assemble('mass')
        self.A_x = self.hierarchy.assemble('stiff')
        self.N = self.A_t.shape[0]
        self.M = self.M_x.shape[0]
        # u0_x = int u0 phi_i, with u0 replaced by its nodal interpolant
        # (NGSolve's quadrature is not observable).  u0
        # vanishes on the boundary, so the free-dof mass matrix suffices.
        u0 = self._u0()
        self.u0_x = self.M_x @ u0
        self._Cinv_j = None

    @property
    def Cinv_j(self):
        """2^j M_x + alpha A_x, j = 0..J_time (heateq_mpi.py:97-98); built on
        first use -- the device path works from M_x and A_x directly."""
        if self._Cinv_j is None:
            self._Cinv_j = [
                2**j * self.M_x + self.alpha * self.A_x
                for j in range(self.J_time + 1)
            ]
        return self._Cinv_j


class CubeProblem(SquareProblem):
    """problem='cube' (problem.py:21-32): u = exp(-3 pi^2 t) sin(pi x) sin(pi y)
    sin(pi z) on the unit cube, Kuhn triangulation."""
    dim = 3

    def _mesh(self, J_space, order, seed):
        return CubeMeshHierarchy(J_space, order, seed)

    def _u0(self):
        return self.hierarchy.nodal_values(
            lambda x, y, z: np.sin(np.pi * x) * np.sin(np.pi * y) * np.sin(
                np.pi * z))
