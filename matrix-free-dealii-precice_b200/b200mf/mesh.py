"""Synthetic structured hex/quad meshes for the benchmarks and parity tests.

The library itself consumes plain index / vertex arrays (include/b200mf.h) —
in production they are extracted from deal.II's MatrixFree.
This module only *generates* such arrays for the synthetic configurations of
BASELINE.json (box, mapped box, cylinder shell as in the reference's
include/cases/elasticity/tube_3d.h:23-33) plus the uniformly refined
multigrid hierarchy that `parallel::distributed::Triangulation` +
`refine_global` would produce (mf_elasticity.h:638).

Numbering: nodes are lexicographic on the global tensor grid (x fastest);
vector DoF = ncomp*node + component (interleaved).  A random permutation of the
nodes can be requested to test that kernels do not rely on the numbering.
"""
from __future__ import annotations

import numpy as np

NODE_MASK = np.uint32((1 << 29) - 1)


def _gll(p):
    # local import keeps the package free of oracle/ dependencies
    from numpy.polynomial import legendre as L

    if p == 1:
        return np.array([0.0, 1.0])
    c = np.zeros(p + 1)
    c[p] = 1.0
    r = L.legroots(L.legder(c))
    for _ in range(3):
        r = r - L.legval(r, L.legder(c)) / L.legval(r, L.legder(c, 2))
    x = np.concatenate([[-1.0], np.sort(r), [1.0]])
    x = 0.5 * (x - x[::-1])
    return 0.5 * (x + 1.0)


class StructuredMesh:
    """ncells = (nx,ny[,nz]) cells of degree p on the image of the unit
    box under `mapping` (a vectorised callable X[...,dim] -> X'[...,dim]
    applied to the *vertices*; geometry inside a cell is the Q1 interpolant,
    like deal.II's StaticMappingQ1)."""

    def __init__(self, ncells, degree, mapping=None, lower=None, upper=None, permute_seed=None,
                 periodic_x=False):
        self.dim = len(ncells)
        self.ncells = tuple(int(c) for c in ncells)
        self.p = int(degree)
        self.n = self.p + 1
        dim, p = self.dim, self.p
        self.periodic_x = periodic_x
        self.nnodes_dir = tuple(c * p + 1 for c in self.ncells)
        if periodic_x:
            self.nnodes_dir = (self.ncells[0] * p,) + self.nnodes_dir[1:]
        self.n_nodes = int(np.prod(self.nnodes_dir))
        self.n_cells = int(np.prod(self.ncells))
        lower = np.zeros(dim) if lower is None else np.asarray(lower, float)
        upper = np.ones(dim) if upper is None else np.asarray(upper, float)
        self.lower, self.upper = lower, upper
        self.mapping = mapping

        # --- cell -> vertex coordinates -------------------------------------
        axes = [np.linspace(lower[d], upper[d], self.ncells[d] + 1) for d in range(dim)]
        cidx = np.meshgrid(*[np.arange(c) for c in self.ncells[::-1]], indexing="ij")
        cidx = [a.ravel() for a in cidx[::-1]]  # cidx[d][cell], x fastest
        self.cell_ijk = np.stack(cidx, axis=1)
        nv = 1 << dim
        verts = np.empty((self.n_cells, nv, dim))
        for m in range(nv):
            for d in range(dim):
                verts[:, m, d] = axes[d][cidx[d] + ((m >> d) & 1)]
        if mapping is not None:
            verts = mapping(verts)
        self.vertices = verts

        # --- cell -> node index (lexicographic local order) -----------------
        loc = np.meshgrid(*[np.arange(self.n)] * dim, indexing="ij")
        loc = [a.ravel() for a in loc[::-1]]  # loc[d][a], x fastest
        self.loc_ijk = np.stack(loc, axis=1)
        node = np.zeros((self.n_cells, self.n**dim), dtype=np.int64)
        stride = 1
        for d in range(dim):
            g = cidx[d][:, None] * p + loc[d][None, :]
            if d == 0 and periodic_x:
                g = g % self.nnodes_dir[0]
            node += g * stride
            stride *= self.nnodes_dir[d]
        if permute_seed is not None:
            perm = np.random.default_rng(permute_seed).permutation(self.n_nodes)
            node = perm[node]
            self.perm = perm
        else:
            self.perm = None
        self.cell_nodes = node
        self.material_id = np.zeros(self.n_cells, dtype=np.uint8)
        self.constrained_nodes = None  # [n_nodes, ncomp] bool, set by clamp()

    # ------------------------------------------------------------------
    def renumber_cellwise_morton(self):
        """deal.II-like enumeration instead of the lexicographic tensor grid: cells in Morton (z-curve) order - the order
        p4est / parallel::distributed::Triangulation walks them (mf_elasticity.h:416-425) - and nodes numbered cell by
        cell in the order they are first touched (DoFHandler::distribute_dofs before Cuthill-McKee, :667-669).  Benchmarks
        use it to show that the gather does not live off the lexicographic numbering.  Single-level meshes only
        (children_of assumes lexicographic cells)."""
        assert self.perm is None and self.dim == 3
        key = np.zeros(self.n_cells, dtype=np.int64)
        ijk = self.cell_ijk.astype(np.int64)
        for b in range(21):
            for d in range(3):
                key |= ((ijk[:, d] >> b) & 1) << (3 * b + d)
        order = np.argsort(key, kind="stable")
        self.cell_ijk = self.cell_ijk[order]
        self.vertices = self.vertices[order]
        self.material_id = self.material_id[order]
        nodes = self.cell_nodes[order]
        flat = nodes.ravel()
        _, first = np.unique(flat, return_index=True)  # first touch of every lexicographic node
        touch_order = np.argsort(first, kind="stable")   # lexicographic node ids sorted by first touch
        uniq = np.unique(flat)
        perm = np.empty(self.n_nodes, dtype=np.int64)
        perm[uniq[touch_order]] = np.arange(uniq.size)
        self.perm = perm
        self.cell_nodes = perm[nodes]
        if self.constrained_nodes is not None:
            out = np.empty_like(self.constrained_nodes)
            out[perm] = self.constrained_nodes
            self.constrained_nodes = out
        return self

    def node_grid_index(self):
        """[n_nodes, dim] integer grid coordinates of every node (in the
        mesh's own numbering)."""
        g = np.meshgrid(*[np.arange(m) for m in self.nnodes_dir[::-1]], indexing="ij")
        g = np.stack([a.ravel() for a in g[::-1]], axis=1)
        if self.perm is not None:
            out = np.empty_like(g)
            out[self.perm] = g
            return out
        return g

    def node_coordinates(self):
        """Support points: Q1 map of the cell vertices at the GLL points."""
        dim = self.dim
        x1 = _gll(self.p)
        if self.mapping is None and not self.periodic_x:
            # axis-aligned box: tensor grid of (cell origin + h * GLL point)
            g = self.node_grid_index()
            coords = np.empty((self.n_nodes, dim))
            for d in range(dim):
                h = (self.upper[d] - self.lower[d]) / self.ncells[d]
                cell, loc = np.divmod(g[:, d], self.p)
                last = cell == self.ncells[d]
                cell = np.where(last, cell - 1, cell)
                loc = np.where(last, self.p, loc)
                coords[:, d] = self.lower[d] + h * (cell + x1[loc])
            return coords
        lin = np.stack([1.0 - x1, x1], axis=1)
        Nv = lin
        for _ in range(dim - 1):
            Nv = np.kron(lin, Nv)  # last factor fastest => x fastest
        # Nv[a, m]
        coords = np.zeros((self.n_nodes, dim))
        pts = np.einsum("am,cme->cae", Nv, self.vertices)
        coords[self.cell_nodes.ravel()] = pts.reshape(-1, dim)
        return coords

    def boundary_nodes(self, d, side):
        """bool[n_nodes]: nodes on the face x_d = lower (side 0) / upper (1)
        of the *parameter* box."""
        g = self.node_grid_index()
        target = 0 if side == 0 else self.nnodes_dir[d] - 1
        return g[:, d] == target

    def clamp(self, faces, ncomp):
        """Homogeneous Dirichlet on the listed (d, side) faces, all components
        (the clamped faces of the reference's cases, e.g. tube_3d.h:47-63)."""
        m = np.zeros(self.n_nodes, dtype=bool)
        for d, side in faces:
            m |= self.boundary_nodes(d, side)
        self.constrained_nodes = np.repeat(m[:, None], ncomp, axis=1)
        return self.constrained_nodes

    # ------------------------------------------------------------------
    def dof_index(self, ncomp):
        """Full [n_cells, ncomp*n^dim] map, component-major (what
        FEEvaluation::read_dof_values walks)."""
        nl = self.n**self.dim
        out = np.empty((self.n_cells, ncomp * nl), dtype=np.int64)
        for c in range(ncomp):
            out[:, c * nl:(c + 1) * nl] = ncomp * self.cell_nodes + c
        return out

    def packed_node_index(self, ncomp):
        """uint32 node index with the per-component constrained flags in the
        top 3 bits (bit 29+c) — the 4 B/node layout of SURVEY.md §8d."""
        assert self.n_nodes <= int(NODE_MASK)
        out = self.cell_nodes.astype(np.uint32)
        if self.constrained_nodes is not None:
            flags = np.zeros(self.n_nodes, dtype=np.uint32)
            for c in range(ncomp):
                flags |= self.constrained_nodes[:, c].astype(np.uint32) << np.uint32(29 + c)
            out |= flags[self.cell_nodes]
        return out

    def constrained_dofs(self, ncomp):
        if self.constrained_nodes is None:
            return np.zeros(self.n_nodes * ncomp, dtype=bool)
        return self.constrained_nodes.reshape(-1).copy()

    def is_affine(self):
        return self.mapping is None

    # ------------------------------------------------------------------
    def refine_parent_of(self):
        """The coarse mesh this one would be the uniform refinement of
        (needs even cell counts)."""
        assert all(c % 2 == 0 for c in self.ncells)
        assert self.perm is None
        return StructuredMesh(tuple(c // 2 for c in self.ncells), self.p, self.mapping,
                              self.lower, self.upper, periodic_x=self.periodic_x)


def children_of(coarse_mesh):
    """[n_coarse_cells, 8] cell indices of the uniformly refined mesh (child (cx,cy,cz) at cx + 2 cy + 4 cz), the
    parent/child relation `refine_global` produces (mf_elasticity.h:638) — input of b200mf_mg_transfer_create."""
    nc = coarse_mesh.ncells
    dim = len(nc)
    ijk = coarse_mesh.cell_ijk
    out = np.zeros((coarse_mesh.n_cells, 1 << dim), dtype=np.int64)
    for ch in range(1 << dim):
        idx = np.zeros(coarse_mesh.n_cells, dtype=np.int64)
        stride = 1
        for d in range(dim):
            idx += (2 * ijk[:, d] + ((ch >> d) & 1)) * stride
            stride *= 2 * nc[d]
        out[:, ch] = idx
    return out


def cylinder_shell_mapping(length, r_in, r_out, scale=1.0):
    """Unit box (s,t,r) -> cylinder shell: axis along x, angle from s.  Matches
    the shape of GridGenerator::cylinder_shell used by
    include/cases/elasticity/tube_3d.h:23-33 (scaled by 1e-2, :74).  Use with
    periodic_x=True and s as the *angular* direction."""

    def f(X):
        out = np.empty_like(X)
        theta = 2.0 * np.pi * X[..., 0]
        r = r_in + (r_out - r_in) * X[..., 2]
        out[..., 0] = scale * length * X[..., 1]
        out[..., 1] = scale * r * np.cos(theta)
        out[..., 2] = scale * r * np.sin(theta)
        return out

    return f


def smooth_mapping(amplitude=0.08):
    """A smooth non-affine perturbation of the unit box (general metric)."""

    def f(X):
        out = X.copy()
        dim = X.shape[-1]
        for d in range(dim):
            e = (d + 1) % dim
            out[..., d] += amplitude * np.sin(np.pi * X[..., e]) * np.sin(np.pi * X[..., d])
        return out

    return f


def synthetic_displacement(coords, amplitude=0.05):
    """u = a*(sin(pi x) cos(pi y), sin(pi y) cos(pi z), sin(pi z) cos(pi x))
    sampled at nodes — SURVEY.md §8d config 3 (det F > 0 guaranteed)."""
    dim = coords.shape[1]
    u = np.empty_like(coords)
    for d in range(dim):
        e = (d + 1) % dim
        u[:, d] = amplitude * np.sin(np.pi * coords[:, d]) * np.cos(np.pi * coords[:, e])
    return u.reshape(-1)


def tube_hierarchy(degree, n_levels, first_refinement=0):
    """Multigrid hierarchy of the tube of include/cases/elasticity/tube_3d.h:23-33, 74: GridGenerator::cylinder_shell(length 5,
    inner radius 0.5, outer radius 0.6, 5 cells around, 7 along the axis), scaled by 1e-2, one cell through the wall,
    refined globally; both ends clamped (:47-63).  Direction 0 is the (periodic) angle, 1 the axis, 2 the radius: the inner
    wall - the coupling interface - is the face (2, 0)."""
    meshes = []
    for r in range(first_refinement, first_refinement + n_levels):
        m = StructuredMesh((5 * 2**r, 7 * 2**r, 2**r), degree, mapping=cylinder_shell_mapping(5.0, 0.5, 0.6, 1e-2), periodic_x=True)
        m.clamp([(1, 0), (1, 1)], 3)
        meshes.append(m)
    return meshes


def tube_pressure_traction(pressure):
    """x -> traction of a uniform pressure on the inner wall (radially outward), the stand-in for the fluid load."""

    def t(x):
        r = np.hypot(x[1], x[2])
        return np.array([0.0, pressure * x[1] / r, pressure * x[2] / r])

    return t
