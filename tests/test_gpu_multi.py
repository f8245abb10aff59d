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


def _worker(rank, world, port, grid, nc, p, q):
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
        op.compute_diagonal()
        diag = op.get_matrix_diagonal()
        q.put((rank, max(errs), diag[:pt.n_owned].min()))
    except BaseException:
        import traceback
        q.put((rank, traceback.format_exc(), 0.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,grid,nc,p", [(2, (2, 1, 1), (3, 2, 2), 2), (2, (1, 2, 1), (2, 2, 3), 3), (2, (1, 1, 2), (4, 4, 5), 4),
                                             (4, (2, 2, 1), (4, 3, 3), 2)])
def test_distributed_vmult_two_gpus(world, grid, nc, p):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, nc, p, q)) for r in range(world)]
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
