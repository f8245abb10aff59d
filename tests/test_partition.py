"""CPU tests of the multi-GPU host logic: brick partition,
ownership, import lists and ghost ranges.  The exchange is emulated in-process
for arbitrary brick grids and run for real over torch.distributed/gloo with
world_size 2; the arithmetic is the CPU oracle applied to each rank's cells
(update_ghost_values -> cells -> compress(add) -> constrained identity,
mf_nh_operator.h:609-624)."""
import os

import numpy as np
import pytest

from b200mf.mesh import StructuredMesh, synthetic_displacement
from b200mf.partition import BrickPartition, weak_scaling_grid
from common import MAT, rel
from oracle import c_oracle as CO


def global_reference(grid, nc, p):
    m = StructuredMesh(tuple(g * c for g, c in zip(grid, nc)), p, lower=(0, 0, 0), upper=grid)
    m.clamp([(0, 0)], 3)
    u = synthetic_displacement(m.node_coordinates(), 0.05)
    invjac, jxw = CO.q1_metric(m.vertices, p + 1)
    op = CO.NeoHookOperator(CO.Cells(p, p + 1, 3, m.dof_index(3), m.constrained_dofs(3), invjac, jxw), formulation=1, **MAT)
    op.initialize(u, "tensor2")
    op.cache()
    src = np.random.default_rng(0).uniform(-1, 1, 3 * m.n_nodes)
    dst = np.zeros_like(src)
    op.vmult(dst, src)
    return m, u, src, dst


def local_operator(part, u_local):
    """oracle on the rank's cells WITHOUT the constrained-identity step (done after compress)."""
    invjac, jxw = CO.q1_metric(part.vertices, part.p + 1)
    nocon = np.zeros(part.n_dofs, dtype=bool)
    cells = CO.Cells(part.p, part.p + 1, 3, part.dof_index(), part.constrained_dofs(), invjac, jxw)
    op = CO.NeoHookOperator(cells, formulation=1, **MAT)
    op.initialize(u_local, "tensor2")
    op.cache()
    return op


def dofs_of_nodes(nodes):
    return (3 * np.asarray(nodes)[:, None] + np.arange(3)[None, :]).reshape(-1)


@pytest.mark.parametrize("grid,nc,p", [((2, 1, 1), (2, 2, 2), 2), ((2, 2, 1), (2, 1, 2), 1), ((2, 2, 2), (1, 2, 1), 2),
                                        ((3, 2, 1), (1, 1, 2), 3)])
def test_partition_emulated_exchange(grid, nc, p):
    world = int(np.prod(grid))
    gm, gu, gsrc, gdst = global_reference(grid, nc, p)
    parts = [BrickPartition(grid, nc, p, r) for r in range(world)]
    # every global node owned exactly once
    owner_count = np.zeros(gm.n_nodes, dtype=int)
    for pt in parts:
        owner_count[pt.global_lex_node()[:pt.n_owned_nodes]] += 1
    assert (owner_count == 1).all()
    for pt in parts:
        pt.clamp_global_face(0, 0)
        # coordinates agree with the global mesh
        assert np.abs(pt.node_coordinates() - gm.node_coordinates()[pt.global_lex_node()]).max() < 1e-14
        # neighbour lists are symmetric
        for nb in pt.neighbors:
            other = [x for x in parts[nb["rank"]].neighbors if x["rank"] == pt.rank][0]
            assert len(nb["send_nodes"]) == other["ghost_count"]
            g_send = pt.global_lex_node()[nb["send_nodes"]]
            g_recv = parts[nb["rank"]].global_lex_node()[other["ghost_start"]:other["ghost_start"] + other["ghost_count"]]
            assert np.array_equal(g_send, g_recv)
    # distributed vmult with emulated update_ghost_values / compress(add)
    src_loc = []
    for pt in parts:
        v = np.zeros(pt.n_dofs)
        g = dofs_of_nodes(pt.global_lex_node()[:pt.n_owned_nodes])
        v[:pt.n_owned] = gsrc[g]
        src_loc.append(v)
    # update_ghost_values
    for pt in parts:
        for nb in pt.neighbors:
            other = [x for x in parts[nb["rank"]].neighbors if x["rank"] == pt.rank][0]
            s, c = 3 * other["ghost_start"], 3 * other["ghost_count"]
            src_loc[nb["rank"]][s:s + c] = src_loc[pt.rank][dofs_of_nodes(nb["send_nodes"])]
    dst_loc = []
    for pt, v in zip(parts, src_loc):
        u_loc = gu[dofs_of_nodes(pt.global_lex_node())]
        op = local_operator(pt, u_loc)
        d = np.zeros(pt.n_dofs)
        op.vmult(d, v)
        con = pt.constrained_dofs()
        d[con] -= v[con]  # undo the oracle's constrained identity; re-applied once after compress
        dst_loc.append(d)
    # compress(add)
    for pt in parts:
        for nb in pt.neighbors:
            other = [x for x in parts[nb["rank"]].neighbors if x["rank"] == pt.rank][0]
            s, c = 3 * nb["ghost_start"], 3 * nb["ghost_count"]
            dst_loc[nb["rank"]][dofs_of_nodes(other["send_nodes"])] += dst_loc[pt.rank][s:s + c]
    for pt, d, v in zip(parts, dst_loc, src_loc):
        con = pt.constrained_dofs()[:pt.n_owned]
        d[:pt.n_owned][con] += v[:pt.n_owned][con]
        g = dofs_of_nodes(pt.global_lex_node()[:pt.n_owned_nodes])
        assert rel(d[:pt.n_owned], gdst[g]) < 1e-13


def test_weak_scaling_grids():
    assert [weak_scaling_grid(w) for w in (1, 2, 4, 8)] == [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2)]
    pt = BrickPartition((2, 2, 2), 2, 2, 7)
    arr = pt.partitioner_arrays()
    assert len(arr["ranks"]) == 7 and arr["send_offsets"][-1] == 0  # last brick owns its high faces: only receives
    pt0 = BrickPartition((2, 2, 2), 2, 2, 0)
    assert pt0.n_ghost_nodes == 0 and len(pt0.partitioner_arrays()["ranks"]) == 7


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid, nc, p = (2, 1, 1), (2, 2, 2), 2
    gm, gu, gsrc, gdst = global_reference(grid, nc, p)
    pt = BrickPartition(grid, nc, p, rank)
    pt.clamp_global_face(0, 0)
    v = np.zeros(pt.n_dofs)
    g_owned = dofs_of_nodes(pt.global_lex_node()[:pt.n_owned_nodes])
    v[:pt.n_owned] = gsrc[g_owned]

    def exchange(vec, reverse):
        reqs, bufs = [], []
        for nb in pt.neighbors:
            s, c = 3 * nb["ghost_start"], 3 * nb["ghost_count"]
            send_idx = dofs_of_nodes(nb["send_nodes"])
            if not reverse:  # update_ghost_values: owner -> ghost
                if send_idx.size:
                    reqs.append(dist.isend(torch.from_numpy(vec[send_idx].copy()), nb["rank"]))
                if c:
                    buf = torch.empty(c, dtype=torch.float64)
                    reqs.append(dist.irecv(buf, nb["rank"]))
                    bufs.append(("ghost", s, c, buf))
            else:  # compress(add): ghost -> owner
                if c:
                    reqs.append(dist.isend(torch.from_numpy(vec[s:s + c].copy()), nb["rank"]))
                if send_idx.size:
                    buf = torch.empty(send_idx.size, dtype=torch.float64)
                    reqs.append(dist.irecv(buf, nb["rank"]))
                    bufs.append(("owned", send_idx, None, buf))
        for r in reqs:
            r.wait()
        for kind, a, b, buf in bufs:
            if kind == "ghost":
                vec[a:a + b] = buf.numpy()
            else:
                np.add.at(vec, a, buf.numpy())
        if reverse:
            vec[pt.n_owned:] = 0.0

    exchange(v, False)
    op = local_operator(pt, gu[dofs_of_nodes(pt.global_lex_node())])
    d = np.zeros(pt.n_dofs)
    op.vmult(d, v)
    con = pt.constrained_dofs()
    d[con] -= v[con]
    exchange(d, True)
    d[:pt.n_owned][con[:pt.n_owned]] += v[:pt.n_owned][con[:pt.n_owned]]
    err = rel(d[:pt.n_owned], gdst[g_owned])
    # distributed dot product (allreduce over owned entries)
    t = torch.tensor([float(d[:pt.n_owned] @ d[:pt.n_owned])], dtype=torch.float64)
    dist.all_reduce(t)
    q.put((rank, err, float(t.item()), float(gdst @ gdst)))
    dist.destroy_process_group()


def test_gloo_world2_vmult():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p_ in procs:
        p_.join(60)
    for rank, err, dot, ref in res:
        assert err < 1e-13
        assert abs(dot - ref) < 1e-12 * ref


@pytest.mark.parametrize("world,p", [(2, 2), (4, 1)])
def test_tube_partition_matches_global_periodic_mesh(world, p):
    """Axial slabs of the tube (cases/elasticity/tube_3d.h geometry; direction 0 periodic inside every rank): local nodes,
    coordinates and cell connectivity agree with the undivided periodic mesh, every global node is owned exactly once, and
    what a rank sends is what its neighbour expects as ghosts."""
    from b200mf.mesh import tube_hierarchy
    from b200mf.partition import tube_partitions

    r = 2 if world == 4 else 1
    g = tube_hierarchy(p, 1, first_refinement=r)[0]
    gc = g.node_coordinates()
    owner = np.full(g.n_nodes, -1)
    parts = [tube_partitions(p, r, 1, world, rank)[0] for rank in range(world)]
    for rank, pt in enumerate(parts):
        gl = pt.global_lex_node()
        assert np.allclose(pt.node_coordinates(), gc[gl])
        assert (owner[gl[:pt.n_owned_nodes]] == -1).all()
        owner[gl[:pt.n_owned_nodes]] = rank
        axial = pt.nc[1]
        for c in range(pt.n_cells):
            i, j, k = pt.cell_ijk[c]
            gcell = (k * g.ncells[1] + (j + axial * rank)) * g.ncells[0] + i
            assert (gl[pt.cell_nodes[c]] == g.cell_nodes[gcell]).all()
        # clamped ends: first and last rank only
        con = pt.constrained_nodes[:, 0]
        assert con.any() == (rank in (0, world - 1))
        assert (g.constrained_nodes[gl, 0] == con).all()
    assert (owner >= 0).all()
    for rank, pt in enumerate(parts):
        gl = pt.global_lex_node()
        for nb in pt.neighbors:
            other = parts[nb["rank"]]
            back = [o for o in other.neighbors if o["rank"] == rank][0]
            sent = gl[nb["send_nodes"]]
            expect = other.global_lex_node()[back["ghost_start"]:back["ghost_start"] + back["ghost_count"]]
            assert (np.sort(sent) == np.sort(expect)).all() and (sent == expect).all()
    assert sum(pt.n_owned for pt in parts) == 3 * g.n_nodes == parts[0].n_global_dofs


@pytest.mark.parametrize("world,top", [(2, 2), (2, 3), (8, 4)])
def test_tube_hierarchy_with_replicated_coarse_levels(world, top):
    """tube_hierarchy_distributed: the levels below the first one that can be cut into `world` aligned slabs stay whole on every
    rank; the children lists of the transfer between the last whole level and the first partitioned one name, per rank, exactly
    the local copies of the global children, and every parent is claimed by exactly one rank."""
    from b200mf.mesh import children_of, tube_hierarchy
    from b200mf.partition import BrickPartition, tube_hierarchy_distributed

    p = 1
    full = tube_hierarchy(p, top + 1)
    claimed = None
    for rank in range(world):
        levels, children, r_dist = tube_hierarchy_distributed(p, top, world, rank)
        assert len(levels) == top + 1 and len(children) == top
        assert [isinstance(l, BrickPartition) for l in levels] == [r >= r_dist for r in range(top + 1)]
        assert (7 * 2 ** (r_dist - 1)) % world == 0
        for r in range(r_dist - 1):
            assert np.array_equal(children[r], children_of(full[r]))
        ch, pt = children[r_dist - 1], levels[r_dist]
        here = (ch >= 0).all(axis=1)
        assert np.array_equal(here, (ch >= 0).any(axis=1)) and here.sum() * 8 == pt.n_cells
        gi = pt.cell_ijk + np.array([pt.b[d] * pt.nc[d] for d in range(3)])[None, :]
        nf = full[r_dist].ncells
        glob = (gi[:, 2] * nf[1] + gi[:, 1]) * nf[0] + gi[:, 0]
        assert np.array_equal(glob[ch[here]], children_of(full[r_dist - 1])[here])
        assert np.allclose(pt.vertices[ch[here][:, 0]], full[r_dist].vertices[children_of(full[r_dist - 1])[here][:, 0]])
        claimed = here.astype(int) if claimed is None else claimed + here
    assert np.all(claimed == 1)
