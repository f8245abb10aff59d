"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the distributed vmult
update_ghost_values -> cell loop -> compress(add) -> constrained identity (mf_nh_operator.h:609-624)
over the library's NCCL partitioner, one process per GPU, against the CPU oracle applied to the
undivided global mesh.  Tolerance 1e-12 relative l2 (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, grid, nc, p, q, p2p=1):
    os.environ["B200MF_P2P"] = str(p2p)
    import faulthandler
    import sys as _sys

    faulthandler.enable()
    faulthandler.dump_traceback_later(60, exit=True, file=_sys.stderr)  # a hung rank reports where and dies
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from b200mf import Communicator, DeviceVector, MatrixFree, NeoHookOperator
        from b200mf.partition import BrickPartition
        from test_partition import dofs_of_nodes, global_reference
        from common import MAT, rel

        comm = Communicator.from_torch_distributed()
        gm, gu, gsrc, gdst = global_reference(grid, nc, p)
        pt = BrickPartition(grid, nc, p, rank)
        pt.clamp_global_face(0, 0)
        mf = MatrixFree.from_partition(pt, comm)
        g_all = dofs_of_nodes(pt.global_lex_node())
        g_owned = g_all[:pt.n_owned]
        op = NeoHookOperator()
        op.set_material(mf, dict(formulation=1, **MAT))
        # displacement: owned entries only; cache() must update the ghosts itself
        u = np.zeros(pt.n_dofs)
        u[:pt.n_owned] = gu[g_owned]
        disp = DeviceVector(mf, u)
        op.initialize(disp, "tensor2")
        op.cache()
        v = np.zeros(pt.n_dofs)
        v[:pt.n_owned] = gsrc[g_owned]
        v[pt.n_owned:] = 123.0  # stale ghosts must be overwritten by update_ghost_values
        s, d = DeviceVector(mf, v), DeviceVector(mf)
        errs = []
        # 1: exchanges hidden behind the interior cells (boundary cells first, second stream); 0: the reference's
        # serial order update_ghost_values -> cells -> compress(add)
        for overlap in (1, 0, 1):  # 0 is the library default
            mf.set_overlap(overlap)
            s.upload(v)  # stale ghosts again
            op.vmult(d, s)
            out = d.download()
            errs.append(rel(out[:pt.n_owned], gdst[g_owned]))
            assert np.all(out[pt.n_owned:] == 0.0)  # compress(add) zeroes the ghost entries
        # vmult_add into a vector whose ghost entries hold stale values (they must not reach the owners)
        d.upload(np.concatenate([gdst[g_owned], np.full(pt.n_dofs - pt.n_owned, 7.0)]))
        op.vmult_add(d, s)
        errs.append(rel(d.download()[:pt.n_owned], 2 * gdst[g_owned]))
        assert mf.exchange_is_p2p() == (os.environ.get("B200MF_P2P", "1") != "0")
        op.compute_diagonal()
        diag = op.get_matrix_diagonal()
        q.put((rank, max(errs), diag[:pt.n_owned].min()))
    except BaseException:
        import traceback
        q.put((rank, traceback.format_exc(), 0.0))
    finally:
        dist.destroy_process_group()


def _spawn(world, target, args):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=target, args=(r, world, port) + tuple(args) + (q,)) for r in range(world)]
    for pr in procs:
        pr.start()
    import queue as _queue

    res = []
    try:
        for _ in range(world):
            res.append(q.get(timeout=200))
    except _queue.Empty:
        pass  # a rank that raised leaves its peers waiting: report the ranks that did answer
    for pr in procs:
        pr.join(5)
        if pr.is_alive():
            pr.kill()
    for r in res:
        assert not isinstance(r[1], str), r[1]
    assert len(res) == world, f"only {len(res)} of {world} ranks answered"
    return sorted(res, key=lambda r: r[0])


def _newton_worker(rank, world, port, grid, p, levels, q):
    """Two Newmark steps of the solid solver (b200mf.solid.Solid) on one brick per rank: distributed residual, cache,
    diagonals, interpolate_to_mg, Chebyshev estimates, GMG-preconditioned CG (mf_elasticity.h:1400-1491, 1981-2070)."""
    import faulthandler
    import sys as _sys

    faulthandler.enable()
    faulthandler.dump_traceback_later(150, exit=True, file=_sys.stderr)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from b200mf import Communicator, MatrixFree
        from b200mf.mesh import children_of
        from b200mf.partition import BrickPartition
        from b200mf.solid import Solid
        from test_gpu_newton import MATERIAL

        comm = Communicator.from_torch_distributed()
        upper = tuple(float(g) for g in grid)
        parts = []
        for l in range(levels):
            pt = BrickPartition(grid, 2**l, p, rank, upper=upper)
            pt.clamp_global_face(0, 0)
            parts.append(pt)
        level_mf = [MatrixFree.from_partition(pt, comm, dtype="f32") for pt in parts]
        sys_mf = MatrixFree.from_partition(parts[-1], comm)
        sol = Solid(parts, delta_t=0.01, body_force=(0.0, -800.0, 200.0), n_cheb_cg=30, matrix_frees=(sys_mf, level_mf),
                    children=[children_of(pt) for pt in parts[:-1]], **MATERIAL)
        sol.run(2)
        fine = parts[-1]
        g_owned = (3 * fine.global_lex_node()[:, None] + np.arange(3)[None, :]).reshape(-1)[:fine.n_owned]
        q.put((rank, (g_owned, sol.total_displacement()[:fine.n_owned]), list(sol.newton_iterations), list(sol.cg_history)))
    except BaseException:
        import traceback
        q.put((rank, traceback.format_exc(), None, None))
    finally:
        dist.destroy_process_group()


def _tube_worker(rank, world, port, p, first_refinement, levels, q):
    """The tube of cases/elasticity/tube_3d.h split into `world` axial slabs (periodic around the axis inside every rank)."""
    import faulthandler
    import sys as _sys

    faulthandler.enable()
    faulthandler.dump_traceback_later(150, exit=True, file=_sys.stderr)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from b200mf import Communicator, MatrixFree
        from b200mf.mesh import children_of, tube_pressure_traction
        from b200mf.partition import tube_partitions
        from b200mf.solid import Solid

        comm = Communicator.from_torch_distributed()
        parts = tube_partitions(p, first_refinement, levels, world, rank)
        level_mf = [MatrixFree.from_partition(pt, comm, dtype="f32") for pt in parts]
        sys_mf = MatrixFree.from_partition(parts[-1], comm)
        sol = Solid(parts, mu=115384.0, nu=0.3, rho=1200.0, delta_t=1e-4, n_cheb_cg=30, matrix_frees=(sys_mf, level_mf),
                    children=[children_of(pt) for pt in parts[:-1]], traction=tube_pressure_traction(5.0e4), traction_faces=((2, 0),))
        sol.run(1)
        fine = parts[-1]
        g_owned = (3 * fine.global_lex_node()[:, None] + np.arange(3)[None, :]).reshape(-1)[:fine.n_owned]
        q.put((rank, (g_owned, sol.total_displacement()[:fine.n_owned]), list(sol.newton_iterations), list(sol.cg_history)))
    except BaseException:
        import traceback
        q.put((rank, traceback.format_exc(), None, None))
    finally:
        dist.destroy_process_group()


def _tube_full_worker(rank, world, port, p, top, q):
    """The tube with the reference's whole hierarchy on `world` ranks: coarse levels replicated, fine levels in axial slabs."""
    import faulthandler
    import sys as _sys

    faulthandler.enable()
    faulthandler.dump_traceback_later(150, exit=True, file=_sys.stderr)
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from b200mf import Communicator, MatrixFree
        from b200mf.mesh import tube_pressure_traction
        from b200mf.partition import BrickPartition, tube_hierarchy_distributed
        from b200mf.solid import Solid

        comm = Communicator.from_torch_distributed()
        levels, children, r_dist = tube_hierarchy_distributed(p, top, world, rank)
        level_mf = [MatrixFree.from_partition(lv, comm, dtype="f32") if isinstance(lv, BrickPartition) else MatrixFree.from_mesh(lv, 3, dtype="f32")
                    for lv in levels]
        sys_mf = MatrixFree.from_partition(levels[-1], comm)
        sol = Solid(levels, mu=115384.0, nu=0.3, rho=1200.0, delta_t=1e-4, n_cheb_cg=30, matrix_frees=(sys_mf, level_mf),
                    children=children, traction=tube_pressure_traction(5.0e4), traction_faces=((2, 0),))
        sol.run(1)
        fine = levels[-1]
        g_owned = (3 * fine.global_lex_node()[:, None] + np.arange(3)[None, :]).reshape(-1)[:fine.n_owned]
        q.put((rank, (g_owned, sol.total_displacement()[:fine.n_owned]), list(sol.newton_iterations), list(sol.cg_history) + [sol.mg.graph_info()["state"]]))
    except BaseException:
        import traceback
        q.put((rank, traceback.format_exc(), None, None))
    finally:
        dist.destroy_process_group()


def test_distributed_tube_full_hierarchy():
    """The reference's whole 35-coarse-cell hierarchy on 2 GPUs: refinements 0 and 1 are kept whole on both ranks (replicated
    coarse levels - restriction sums the ranks' shares, interpolation merges them), refinement 2 is cut into two axial slabs.
    Same Newton / CG counts and displacement as the one-GPU run of the same three levels."""
    world, p, top = 2, 2, 2
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    from b200mf.mesh import tube_hierarchy, tube_pressure_traction
    from b200mf.solid import Solid

    res = _spawn(world, _tube_full_worker, (p, top))
    meshes = tube_hierarchy(p, top + 1)
    ref = Solid(meshes, mu=115384.0, nu=0.3, rho=1200.0, delta_t=1e-4, n_cheb_cg=30, traction=tube_pressure_traction(5.0e4),
                traction_faces=((2, 0),))
    ref.run(1)
    u_ref = ref.total_displacement()
    got = np.full_like(u_ref, np.nan)
    for rank, (idx, val), newton, cg in res:
        got[idx] = val
        assert newton == ref.newton_iterations
        assert cg[-1] == 1  # lower levels (replicated ones and the all-reduce of the restriction included) ran from the recorded graph
        assert max(abs(a - b) for a, b in zip(cg[:-1], ref.cg_history)) <= 1, (cg, ref.cg_history)
    assert not np.isnan(got).any()
    assert np.abs(got - u_ref).max() <= 1e-7 * np.abs(u_ref).max()


def test_distributed_tube_newton():
    """BASELINE configs[3] in miniature on 2 GPUs: tube geometry (curved cells, periodic around the axis), interface traction
    on the inner wall, GMG-preconditioned Newton step, against the one-GPU run of the same hierarchy."""
    world, p, r0, levels = 2, 2, 1, 2
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    from b200mf.mesh import tube_hierarchy, tube_pressure_traction
    from b200mf.solid import Solid

    res = _spawn(world, _tube_worker, (p, r0, levels))
    meshes = tube_hierarchy(p, levels, first_refinement=r0)
    ref = Solid(meshes, mu=115384.0, nu=0.3, rho=1200.0, delta_t=1e-4, n_cheb_cg=30, traction=tube_pressure_traction(5.0e4),
                traction_faces=((2, 0),))
    ref.run(1)
    u_ref = ref.total_displacement()
    got = np.full_like(u_ref, np.nan)
    for rank, (idx, val), newton, cg in res:
        got[idx] = val
        assert newton == ref.newton_iterations
        assert max(abs(a - b) for a, b in zip(cg, ref.cg_history)) <= 1, (cg, ref.cg_history)
    assert not np.isnan(got).any()
    assert np.abs(got - u_ref).max() <= 1e-7 * np.abs(u_ref).max()


@pytest.mark.parametrize("world,grid,p,levels", [(2, (2, 1, 1), 2, 2), (2, (2, 1, 1), 1, 3)])
def test_distributed_newton_gmg(world, grid, p, levels):
    """BASELINE config 4 in miniature: the GMG-preconditioned Newton step on more than one GPU (per-level brick partitions,
    level vectors exchanged through the same partitioner) reproduces the one-GPU run of the same global mesh."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    from b200mf.mesh import StructuredMesh
    from b200mf.solid import Solid
    from test_gpu_newton import MATERIAL

    res = _spawn(world, _newton_worker, (grid, p, levels))
    meshes = []
    for l in range(levels):
        m = StructuredMesh(tuple(g * 2**l for g in grid), p, upper=tuple(float(g) for g in grid))
        m.clamp([(0, 0)], 3)
        meshes.append(m)
    ref = Solid(meshes, delta_t=0.01, body_force=(0.0, -800.0, 200.0), n_cheb_cg=30, **MATERIAL)
    ref.run(2)
    u_ref = ref.total_displacement()
    got = np.full_like(u_ref, np.nan)
    for rank, (idx, val), newton, cg in res:
        got[idx] = val
        assert newton == ref.newton_iterations
        # the deterministic Chebyshev start vector follows the (rank-contiguous) global enumeration, which differs from the
        # one-rank numbering: eigenvalue estimates and hence CG counts may move by one
        assert max(abs(a - b) for a, b in zip(cg, ref.cg_history)) <= 1, (cg, ref.cg_history)
    assert not np.isnan(got).any()
    assert np.abs(got - u_ref).max() <= 1e-7 * np.abs(u_ref).max()


@pytest.mark.parametrize("p2p", [1, 0], ids=["p2p", "nccl"])
@pytest.mark.parametrize("world,grid,nc,p", [(2, (2, 1, 1), (3, 2, 2), 2), (2, (1, 2, 1), (2, 2, 3), 3), (2, (1, 1, 2), (4, 4, 5), 4),
                                             (4, (2, 2, 1), (4, 3, 3), 2)])
def test_distributed_vmult_two_gpus(world, grid, nc, p, p2p):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, nc, p, q, p2p)) for r in range(world)]
    for pr in procs:
        pr.start()
    import queue as _queue

    res = []
    try:
        for _ in range(world):
            res.append(q.get(timeout=100))
    except _queue.Empty:
        pass  # a rank that raised leaves its peers waiting in NCCL: report the ranks that did answer
    for pr in procs:
        pr.join(5)
        if pr.is_alive():
            pr.kill()
    for rank, err, dmin in res:
        assert not isinstance(err, str), err
    assert len(res) == world, f"only {len(res)} of {world} ranks answered"
    for rank, err, dmin in res:
        assert not isinstance(err, str), err
        assert err < 1e-12, (rank, err)
        assert dmin > 0.0
