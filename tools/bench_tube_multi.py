#!/usr/bin/env python
"""BASELINE.json configs[3]: tube3d-geometry neo-Hookean Newton step with GMG-Chebyshev on N GPUs (one axial slab per rank,
NVLink ghost exchange).  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
      tools/bench_tube_multi.py --degree 3 --first 3 --levels 4
The reference's hierarchy goes down to 35 cells (p4est leaves ranks empty on coarse levels); the brick partitioner keeps
every level on every rank, so the coarsest level is refinement `--first` (>= 3 for 8 ranks) and the coarse-grid Chebyshev
solver works on that level."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "matrix-free-dealii-precice_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--degree", type=int, default=3)
ap.add_argument("--first", type=int, default=3)
ap.add_argument("--levels", type=int, default=3)
ap.add_argument("--pressure", type=float, default=5.0e3)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--coarse", default="chebyshev", choices=["chebyshev", "cg"])
ap.add_argument("--full-hierarchy", type=int, default=-1, metavar="TOP",
                help="the reference's whole hierarchy, refinements 0..TOP: levels too coarse for aligned slabs are kept whole on every rank")
a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from b200mf import Communicator, MatrixFree, lib as L_
from b200mf.mesh import children_of, tube_pressure_traction
from b200mf.partition import tube_partitions
from b200mf.solid import Solid

comm = Communicator.from_torch_distributed()
if a.full_hierarchy >= 0:
    from b200mf.partition import BrickPartition, tube_hierarchy_distributed

    parts, children, r_dist = tube_hierarchy_distributed(a.degree, a.full_hierarchy, world, rank)
    level_mf = [MatrixFree.from_partition(lv, comm, dtype="f32") if isinstance(lv, BrickPartition) else MatrixFree.from_mesh(lv, 3, dtype="f32")
                for lv in parts]
    a.first, a.levels = 0, len(parts)
    hierarchy = f"levels r=0..{a.full_hierarchy} (r<{r_dist} whole on every rank)"
else:
    parts = tube_partitions(a.degree, a.first, a.levels, world, rank)
    children = [children_of(pt) for pt in parts[:-1]]
    level_mf = [MatrixFree.from_partition(pt, comm, dtype="f32") for pt in parts]
    hierarchy = f"levels r={a.first}..{a.first + a.levels - 1}"
sys_mf = MatrixFree.from_partition(parts[-1], comm)
sol = Solid(parts, mu=115384.0, nu=0.3, rho=1200.0, delta_t=1e-4, n_cheb_cg=30, matrix_frees=(sys_mf, level_mf),
            children=children, traction=tube_pressure_traction(a.pressure), traction_faces=((2, 0),),
            coarse_solver=a.coarse)
dist.barrier(); torch.cuda.synchronize()
launches0 = L_.load().b200mf_launch_count()
steps = []
for step in range(a.steps):
    t0 = time.perf_counter()
    n0 = len(sol.cg_history)
    try:
        its = sol.solve_nonlinear_timestep()
        sol.advance_time_window()
        ok = True
    except RuntimeError:  # "No convergence in nonlinear solver!" (mf_elasticity.h:1486-1489): report what was done
        its, ok = sol.max_newton, False
    sol.mf.synchronize(); dist.barrier()
    t1 = time.perf_counter()
    steps.append({"seconds": t1 - t0, "newton_iterations": int(its), "converged": ok, "cg_iterations": list(map(int, sol.cg_history[n0:])),
                  "seconds_per_newton_iteration": (t1 - t0) / max(1, len(sol.cg_history) - n0)})
    if not ok:
        break
umax = torch.tensor([float(np.abs(sol.total_displacement()[:parts[-1].n_owned]).max())], device="cuda", dtype=torch.float64)
dist.all_reduce(umax, op=dist.ReduceOp.MAX)
if rank == 0:
    fine = parts[-1]
    print(json.dumps({"workload": f"tube3d Newmark steps, neo-Hookean tensor2, Q{a.degree}, {fine.grid[1]} axial slabs of {fine.nc} curved cells, "
                                  f"{fine.n_global_dofs} DoFs, {hierarchy}, GMG-Chebyshev CG, coarse={a.coarse}, "
                                  f"pressure {a.pressure} on the inner wall",
                      "n_gpus": world, "n_dofs": fine.n_global_dofs, "steps": steps, "level0": sol.eig[0], "level_top": sol.eig[-1],
                      "gpu_launches_rank0": int(L_.load().b200mf_launch_count() - launches0), "max_displacement": float(umax.item()),
                      "exchange_path": "p2p" if sys_mf.exchange_is_p2p() else "nccl", "mg_graph": sol.mg.graph_info(),
                      "note": "step 1 includes the one-off creation of smoother / solver objects"}), flush=True)
dist.barrier()
dist.destroy_process_group()
