"""Domain decomposition of a structured box into one brick per rank (= per GPU).

Mirrors what the reference gets from p4est + deal.II:
one contiguous block of cells per rank, every shared DoF owned by the
lowest-rank adjacent brick, local vector = [owned | ghosts sorted by global
index] (LinearAlgebra::distributed::Vector / Utilities::MPI::Partitioner,
SURVEY.md A.9), and per neighbour an import list (owned entries another rank
ghosts) and a contiguous ghost range.

Everything is computed locally from (grid, rank): no global mesh is ever
formed, so 8 x 5e7 DoFs set up in the same time as one brick.
"""
from __future__ import annotations

import numpy as np

from .mesh import _gll


def rank_of(b, grid):
    return int(b[0] + grid[0] * (b[1] + grid[1] * b[2]))


def brick_of(rank, grid):
    return (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))


def weak_scaling_grid(world):
    """(1,2,2,2) x (1,1,2,2) x (1,1,1,2) bricks at 1/2/4/8 ranks."""
    g = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    if world not in g:
        raise ValueError("world size must be 1, 2, 4 or 8")
    return g[world]


class BrickPartition:
    """Local view of rank `rank` of a box of grid[0] x grid[1] x grid[2] bricks with nc^3 cells of degree p each
    (nc may be a 3-tuple).  The global box is [0, grid] scaled to `lower`/`upper`."""

    def __init__(self, grid, nc, p, rank, ncomp=3, lower=(0.0, 0.0, 0.0), upper=None, periodic0=False, mapping=None):
        """periodic0: direction 0 is periodic (grid[0] must be 1: the periodic direction is never split), e.g. the angle of
        the tube; mapping: vectorised map of the box vertices (b200mf.mesh.cylinder_shell_mapping), Q1 geometry inside a cell."""
        self.grid = tuple(int(g) for g in grid)
        self.nc = (int(nc),) * 3 if np.isscalar(nc) else tuple(int(c) for c in nc)
        self.p, self.rank, self.ncomp = int(p), int(rank), int(ncomp)
        self.world = int(np.prod(self.grid))
        self.b = brick_of(rank, self.grid)
        p, nc, b, grid = self.p, self.nc, self.b, self.grid
        self.n = p + 1
        self.periodic0, self.mapping = bool(periodic0), mapping
        assert not periodic0 or self.grid[0] == 1, "the periodic direction must not be split"
        self.nn_dir = tuple(c * p + (0 if (periodic0 and d == 0) else 1) for d, c in enumerate(nc))  # local node grid
        nn = self.nn_dir
        self.n_cells = int(np.prod(nc))
        self.n_local_nodes = int(np.prod(nn))
        self.lower = np.asarray(lower, float)
        self.upper = np.asarray(self.grid, float) if upper is None else np.asarray(upper, float)

        # ---- ownership: nodes on a low face that has a neighbour brick are ghosts ----------------
        I = np.meshgrid(*[np.arange(m) for m in nn[::-1]], indexing="ij")
        ijk = [a.ravel() for a in I[::-1]]  # ijk[d][local lexicographic node], x fastest
        self._ijk = ijk
        low = [(ijk[d] == 0) & (b[d] > 0) for d in range(3)]
        ghost = low[0] | low[1] | low[2]
        owned_lex = np.flatnonzero(~ghost)
        self.n_owned_nodes = owned_lex.size
        # owner brick offset code: bit d set <=> owner has b_d - 1
        code = low[0].astype(np.int64) | (low[1].astype(np.int64) << 1) | (low[2].astype(np.int64) << 2)
        # local numbering: owned (lexicographic) then ghosts sorted by (owner rank, owner-local owned index)
        local_of_lex = np.full(self.n_local_nodes, -1, dtype=np.int64)
        local_of_lex[owned_lex] = np.arange(owned_lex.size)
        self.neighbors = []  # dicts: rank, send_nodes (local owned idx), ghost_start, ghost_count (in nodes)
        ghost_blocks = []
        pos = owned_lex.size
        owners = []
        for c in range(1, 8):
            d = (c & 1, (c >> 1) & 1, (c >> 2) & 1)
            if any(d[k] and b[k] == 0 for k in range(3)):
                continue
            ob = tuple(b[k] - d[k] for k in range(3))
            owners.append((rank_of(ob, grid), c, ob))
        owners.sort()
        recv = {}
        for orank, c, ob in owners:
            sel = np.flatnonzero(code == c)  # my lexicographic nodes owned by ob, ascending lex
            if sel.size == 0:
                continue
            # index in the owner's local grid: low-face directions map to the owner's high face
            oi = [np.where(np.full(sel.size, (c >> k) & 1, dtype=bool), nn[k] - 1, ijk[k][sel]) for k in range(3)]
            oidx = self._owned_index_in(ob, oi)
            order = np.argsort(oidx, kind="stable")
            local_of_lex[sel[order]] = pos + np.arange(sel.size)
            recv[orank] = (pos, sel.size)
            pos += sel.size
        self.n_ghost_nodes = pos - owned_lex.size
        assert (local_of_lex >= 0).all()
        self.local_of_lex = local_of_lex

        # ---- import lists: my owned nodes that brick b + d ghosts --------------------------------
        send = {}
        for c in range(1, 8):
            d = (c & 1, (c >> 1) & 1, (c >> 2) & 1)
            tb = tuple(b[k] + d[k] for k in range(3))
            if any(tb[k] >= grid[k] for k in range(3)):
                continue
            # nodes of tb whose owner code (as seen from tb) is exactly c: on tb's low face in the d=1
            # directions, and in the d=0 directions NOT on a low face that has a neighbour
            rng = []
            for k in range(3):
                if d[k]:
                    rng.append(np.array([nn[k] - 1]))  # my high face
                else:
                    rng.append(np.arange(1 if tb[k] > 0 else 0, nn[k]))
            gx, gy, gz = np.meshgrid(rng[0], rng[1], rng[2], indexing="ij")
            lex = (gz.ravel() * nn[1] + gy.ravel()) * nn[0] + gx.ravel()
            loc = np.sort(local_of_lex[lex])
            assert loc.size == 0 or loc[-1] < self.n_owned_nodes
            if loc.size:
                send[rank_of(tb, grid)] = loc
        for r in sorted(set(send) | set(recv)):
            gs, gc = recv.get(r, (0, 0))
            self.neighbors.append(dict(rank=r, send_nodes=send.get(r, np.zeros(0, dtype=np.int64)), ghost_start=gs,
                                       ghost_count=gc))

        # ---- cells -> local nodes ----------------------------------------------------------------
        n = self.n
        cI = np.meshgrid(*[np.arange(c) for c in nc[::-1]], indexing="ij")
        cidx = [a.ravel() for a in cI[::-1]]
        lI = np.meshgrid(*[np.arange(n)] * 3, indexing="ij")
        loc = [a.ravel() for a in lI[::-1]]
        lex = np.zeros((self.n_cells, n**3), dtype=np.int64)
        stride = 1
        for d in range(3):
            g1 = cidx[d][:, None] * p + loc[d][None, :]
            if periodic0 and d == 0:
                g1 = g1 % nn[0]
            lex += g1 * stride
            stride *= nn[d]
        self.cell_nodes = local_of_lex[lex]
        # vertices (global box coordinates)
        h = (self.upper - self.lower) / (np.asarray(grid) * np.asarray(nc))
        self.h = h
        verts = np.empty((self.n_cells, 8, 3))
        for m in range(8):
            for d in range(3):
                verts[:, m, d] = self.lower[d] + h[d] * (b[d] * nc[d] + cidx[d] + ((m >> d) & 1))
        self.vertices = verts if mapping is None else mapping(verts)
        self.material_id = np.zeros(self.n_cells, dtype=np.uint8)
        self.constrained_nodes = np.zeros((self.n_local_nodes, ncomp), dtype=bool)
        # the attributes b200mf.mesh.children_of needs (local cells are lexicographic, x fastest, like StructuredMesh)
        self.ncells = self.nc
        self.cell_ijk = np.stack(cidx, axis=1)
        # position in the global (deal.II: rank-contiguous) enumeration: owned DoFs of the lower ranks come first
        def owned_of(r):
            ob = brick_of(r, grid)
            return int(np.prod([nn[k] - (1 if ob[k] > 0 else 0) for k in range(3)])) * ncomp
        self.first_global_dof = sum(owned_of(r) for r in range(self.rank))
        self.n_global_dofs = int(np.prod([grid[k] * nc[k] * p + (0 if (periodic0 and k == 0) else 1) for k in range(3)])) * ncomp

    # owned index (in brick ob's local numbering) of grid nodes oi=(ix,iy,iz) of that brick
    def _owned_index_in(self, ob, oi):
        nn = self.nn_dir
        # owned nodes of ob are those with i_k >= lo_k where lo_k = 1 if ob_k > 0 else 0, lexicographic
        lo = [1 if ob[k] > 0 else 0 for k in range(3)]
        ext = [nn[k] - lo[k] for k in range(3)]
        return ((oi[2] - lo[2]) * ext[1] + (oi[1] - lo[1])) * ext[0] + (oi[0] - lo[0])

    # ------------------------------------------------------------------
    @property
    def n_nodes(self):
        return self.n_local_nodes

    @property
    def n_dofs(self):
        return self.n_local_nodes * self.ncomp

    @property
    def n_owned(self):
        return self.n_owned_nodes * self.ncomp

    def global_grid_index(self):
        """[n_local_nodes, 3] global tensor-grid coordinates of every local node (local numbering)."""
        out = np.empty((self.n_local_nodes, 3), dtype=np.int64)
        for d in range(3):
            out[self.local_of_lex, d] = self.b[d] * self.nc[d] * self.p + self._ijk[d]
        return out

    def global_lex_node(self):
        """index of every local node in a global StructuredMesh(grid*nc, p) (lexicographic, unpermuted)."""
        g = self.global_grid_index()
        gn = [self.grid[d] * self.nc[d] * self.p + (0 if (self.periodic0 and d == 0) else 1) for d in range(3)]
        return (g[:, 2] * gn[1] + g[:, 1]) * gn[0] + g[:, 0]

    def node_coordinates(self):
        if self.mapping is not None or self.periodic0:
            # Q1 map of the (mapped) cell vertices at the GLL points, as StructuredMesh.node_coordinates
            x1 = _gll(self.p)
            lin = np.stack([1.0 - x1, x1], axis=1)
            Nv = lin
            for _ in range(2):
                Nv = np.kron(lin, Nv)
            coords = np.zeros((self.n_local_nodes, 3))
            pts = np.einsum("am,cme->cae", Nv, self.vertices)
            coords[self.cell_nodes.ravel()] = pts.reshape(-1, 3)
            return coords
        g = self.global_grid_index()
        x1 = _gll(self.p)
        out = np.empty((self.n_local_nodes, 3))
        for d in range(3):
            cell, loc = np.divmod(g[:, d], self.p)
            last = cell == self.grid[d] * self.nc[d]
            cell = np.where(last, cell - 1, cell)
            loc = np.where(last, self.p, loc)
            out[:, d] = self.lower[d] + self.h[d] * (cell + x1[loc])
        return out

    def clamp_global_face(self, d, side):
        """homogeneous Dirichlet on the global face x_d = lower/upper (all components)."""
        g = self.global_grid_index()
        target = 0 if side == 0 else self.grid[d] * self.nc[d] * self.p
        self.constrained_nodes |= (g[:, d] == target)[:, None]

    def constrained_dofs(self):
        return self.constrained_nodes.reshape(-1).copy()

    def dof_index(self):
        nl = self.n**3
        out = np.empty((self.n_cells, self.ncomp * nl), dtype=np.int64)
        for c in range(self.ncomp):
            out[:, c * nl:(c + 1) * nl] = self.ncomp * self.cell_nodes + c
        return out

    # partitioner arrays in DoF units, as handed to b200mf_ctx_set_partitioner ---------------------
    def partitioner_arrays(self):
        nc = self.ncomp
        ranks = np.array([nb["rank"] for nb in self.neighbors], dtype=np.int32)
        send_off = [0]
        send_idx = []
        ghost_off, ghost_cnt = [], []
        for nb in self.neighbors:
            s = (nc * nb["send_nodes"][:, None] + np.arange(nc)[None, :]).reshape(-1)
            send_idx.append(s)
            send_off.append(send_off[-1] + s.size)
            ghost_off.append(nc * nb["ghost_start"])
            ghost_cnt.append(nc * nb["ghost_count"])
        send_idx = np.concatenate(send_idx) if send_idx else np.zeros(0, dtype=np.int64)
        return dict(ranks=ranks, send_offsets=np.asarray(send_off, dtype=np.int64),
                    send_indices=send_idx.astype(np.uint32), ghost_offsets=np.asarray(ghost_off, dtype=np.int64),
                    ghost_counts=np.asarray(ghost_cnt, dtype=np.int64))


def tube_partitions(degree, first_refinement, n_levels, world, rank):
    """Level hierarchy of the tube (b200mf.mesh.tube_hierarchy: 5 x 7 x 1 coarse cells, direction 0 = periodic angle, 1 = axis,
    2 = radius) cut into `world` axial slabs, one per rank: refinement r has 7 * 2^r axial cells, so world = 8 needs
    first_refinement >= 3.  Both ends clamped (tube_3d.h:47-63)."""
    from .mesh import cylinder_shell_mapping

    parts = []
    for r in range(first_refinement, first_refinement + n_levels):
        axial = 7 * 2**r
        assert axial % world == 0, f"{axial} axial cells do not split into {world} slabs"
        pt = BrickPartition((1, world, 1), (5 * 2**r, axial // world, 2**r), degree, rank, upper=(1.0, 1.0, 1.0), periodic0=True,
                            mapping=cylinder_shell_mapping(5.0, 0.5, 0.6, 1e-2))
        pt.clamp_global_face(1, 0)
        pt.clamp_global_face(1, 1)
        parts.append(pt)
    return parts


def children_of_replicated(coarse_mesh, fine_part):
    """[n_coarse_cells, 8] children lists for a transfer between a REPLICATED coarse level (a whole b200mf.mesh.StructuredMesh
    on every rank) and a partitioned fine level (BrickPartition): local fine cell indices (child (cx,cy,cz) at cx + 2 cy + 4 cz)
    for the parents whose children lie in this rank's brick, -1 for all other parents.  The brick boundaries must coincide with
    coarse cell boundaries (every parent has all or none of its children here)."""
    nc, b = fine_part.nc, fine_part.b
    assert all((b[d] * nc[d]) % 2 == 0 and nc[d] % 2 == 0 for d in range(3)), "brick boundaries cut through coarse cells"
    ijk = coarse_mesh.cell_ijk
    assert all(coarse_mesh.ncells[d] * 2 == fine_part.grid[d] * nc[d] for d in range(3)), "levels are not one refinement apart"
    out = np.full((coarse_mesh.n_cells, 8), -1, dtype=np.int64)
    lo = np.array([b[d] * nc[d] for d in range(3)])
    for ch in range(8):
        g = 2 * ijk + np.array([(ch >> d) & 1 for d in range(3)])[None, :]  # global fine cell (i, j, k)
        loc = g - lo[None, :]
        here = np.all((loc >= 0) & (loc < np.array(nc)[None, :]), axis=1)
        out[here, ch] = (loc[here, 2] * nc[1] + loc[here, 1]) * nc[0] + loc[here, 0]
    n_here = (out >= 0).sum(axis=1)
    assert np.all((n_here == 0) | (n_here == 8))
    return out


def tube_hierarchy_distributed(degree, top_refinement, world, rank):
    """The reference's whole multigrid hierarchy of the tube (35 coarse cells, refinements 0..top_refinement) on `world` ranks:
    the levels fine enough to be cut into `world` axial slabs whose boundaries follow the cells of the level below are
    partitioned (BrickPartition); all coarser levels are kept whole on every rank (StructuredMesh) - p4est would leave most ranks
    without cells there, here every rank repeats that small amount of work and no rank waits for it.
    Returns (levels, children, first_partitioned_level)."""
    from .mesh import children_of, cylinder_shell_mapping, tube_hierarchy

    r_dist = next((r for r in range(1, top_refinement + 1) if (7 * 2 ** (r - 1)) % world == 0), None)
    if world == 1:
        r_dist = top_refinement + 1
    assert r_dist is not None, f"refinement {top_refinement} is too coarse for {world} axial slabs"
    levels = list(tube_hierarchy(degree, min(r_dist, top_refinement + 1)))  # refinements 0 .. r_dist-1, whole
    for r in range(r_dist, top_refinement + 1):
        axial = 7 * 2**r
        pt = BrickPartition((1, world, 1), (5 * 2**r, axial // world, 2**r), degree, rank, upper=(1.0, 1.0, 1.0), periodic0=True,
                            mapping=cylinder_shell_mapping(5.0, 0.5, 0.6, 1e-2))
        pt.clamp_global_face(1, 0)
        pt.clamp_global_face(1, 1)
        levels.append(pt)
    children = []
    for l in range(len(levels) - 1):
        if isinstance(levels[l], BrickPartition) or not isinstance(levels[l + 1], BrickPartition):
            children.append(children_of(levels[l]))
        else:
            children.append(children_of_replicated(levels[l], levels[l + 1]))
    return levels, children, r_dist
