#!/usr/bin/env python
"""Benchmark of the hot path: neo-Hookean tangent `vmult` (FP64, tensor2 caching)
on a synthetic 3-D box mesh (BASELINE.json configs[2]/[4]: ~5e7 DoF per GPU).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the host cores

One JSON line on stdout (rank 0).  A "step" is one vmult over the whole mesh.
`value`  = DoFs/s with vectors and tables resident in HBM (CUDA events),
`e2e`    = DoFs/s through b200mf_nh_vmult_host: pinned host src -> H2D -> vmult -> D2H -> host dst,
`roofline.achieved` = algorithmic bytes (tensor2 with symmetric tau) / kernel time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "matrix-free-dealii-precice_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

CELLS_PER_DIR = {1: 254, 2: 127, 3: 85, 4: 64, 5: 51, 6: 42}  # SURVEY.md §8d config 3 (~5e7 DoF)
MAT = dict(mu=0.5e6, nu=0.4, rho=1000.0, alpha_1=1.0 / (0.25 * 0.01**2), formulation=1)
METRIC = "nh_tangent_vmult_dofs_per_s"


# q-point values the kernel of each variant streams per quadrature point (csrc/ops.cuh): metric (9+1) plus
#   tensor2: c1, c2, symmetric tau/J (6)            scalar_referential: c + 3 cached displacement values
#   tensor4: c2, tau/J (6), 6x6 tensor (36)         tensor4_ns: 81      form0: J^(-1/3), F (9)
QPOINT_VALUES = {"tensor2": 18, "scalar_referential": 14, "tensor4": 53, "tensor4_ns": 91, "form0": 20}


# roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum per launch of the cell-loop kernel, read from the committed
# `ncu --set full` raw page of the same workload (it cannot be measured inside a timed run); None if no capture is committed
NCU_RAW = "profiles/r02_ncu_q{p}_raw.csv"


def ncu_traffic(p, variant, dtype, default_mesh):
    if variant != "tensor2" or dtype != "f64" or not default_mesh:
        return None, None
    import csv

    path = os.path.join(ROOT, NCU_RAW.format(p=p))
    try:
        rows = list(csv.reader(open(path)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(key)
            tot += float(vals[i]) * scale[units[i]]
        return tot, NCU_RAW.format(p=p)
    except Exception:
        return None, None


def algorithmic_bytes_per_cell(p, esize=8, variant="tensor2"):
    """src read + dst write per unique DoF, one uint32 per node, and the q-point
    table the kernel of this variant actually streams."""
    n = p + 1
    return 2 * esize * 3 * p**3 + 4 * n**3 + QPOINT_VALUES[variant] * esize * n**3


def build_problem(p, ncell_dir, dtype="f64", comm=None, caching="tensor2", formulation=1, numbering="lexicographic", curved=False):
    """One brick of ncell_dir^3 cells per rank (1x1x1, 2x1x1, 2x2x1, 2x2x2 bricks);
    shared faces are exchanged through the library's partitioner (update_ghost_values / compress(add))."""
    from b200mf import MatrixFree, NeoHookOperator
    from b200mf.mesh import StructuredMesh, synthetic_displacement

    if comm is not None and comm.world > 1:
        from b200mf.partition import BrickPartition, weak_scaling_grid

        m = BrickPartition(weak_scaling_grid(comm.world), ncell_dir, p, comm.rank)
        m.clamp_global_face(0, 0)
        u = synthetic_displacement(m.node_coordinates(), 0.05)
        mf = MatrixFree.from_partition(m, comm, dtype=dtype)
    else:
        from b200mf.mesh import smooth_mapping

        m = StructuredMesh((ncell_dir,) * 3, p, mapping=smooth_mapping(0.05) if curved else None)
        m.clamp([(0, 0)], 3)
        if numbering == "cellwise_morton":
            m.renumber_cellwise_morton()
        u = synthetic_displacement(m.node_coordinates(), 0.05)
        mf = MatrixFree.from_mesh(m, 3, dtype=dtype)
    op = NeoHookOperator()
    op.set_material(mf, dict(MAT, formulation=formulation))
    op.initialize(u, caching)
    op.cache()
    return m, mf, op


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML from a thread, ~10 ms period - a 1 ms poll measurably slows the launch path of multi-rank steps (driver locks);
    the timed region of the default run is only tens of ms, too short for `nvidia-smi -lms`)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.rows, self.index, self.stop, self.thread, self.h, self.max_mhz = [], index, False, None, None, None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _sample(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        return mhz, bits

    def _run(self):
        while not self.stop:
            try:
                self.rows.append(self._sample())
            except Exception:
                break
            time.sleep(0.010)

    def __enter__(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        sm = [r[0] for r in self.rows]
        bits = 0
        for r in self.rows:
            bits |= r[1]
        reasons = sorted(k for k, m in self.REASONS.items() if bits & m)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "how": "NVML clock + clocks-event-reasons polled every ~10 ms inside the timed region"}


def host_stream_gbs():
    """STREAM triad of the host (all threads) through the oracle library: what the CPU arm's algorithmic GB/s is to be read against."""
    from oracle import c_oracle as CO

    return CO.host_stream_triad_gbs()


def cpu_oracle_rate(p, ncell_dir, min_seconds=10.0, max_reps=100000):
    """DoFs/s of the CPU oracle port (oracle/mf_oracle.c, OpenMP over cells, all host threads)."""
    from b200mf.mesh import StructuredMesh, synthetic_displacement
    from oracle import c_oracle as CO

    m = StructuredMesh((ncell_dir,) * 3, p)
    m.clamp([(0, 0)], 3)
    invjac, jxw = CO.q1_metric(m.vertices, p + 1)
    cells = CO.Cells(p, p + 1, 3, m.dof_index(3), m.constrained_dofs(3), invjac, jxw)
    mat = {k: v for k, v in MAT.items() if k != "formulation"}
    op = CO.NeoHookOperator(cells, formulation=1, **mat)
    op.initialize(synthetic_displacement(m.node_coordinates(), 0.05), "tensor2")
    op.cache()
    n = 3 * m.n_nodes
    src = np.random.default_rng(0).uniform(-1, 1, n)
    dst = np.zeros(n)
    apply = cpu_apply(op, dst, src)  # warm-up inside
    reps, t0 = 0, time.perf_counter()
    while True:
        apply()
        reps += 1
        el = time.perf_counter() - t0
        if el > min_seconds or reps >= max_reps:
            break
    return n * reps / el, n, reps, CO.num_threads(), el / reps


def cpu_apply(op, dst, src):
    """The CPU arm's vmult: SIMD over cells (8 cells per batch, one lane per cell - deal.II's VectorizedArray scheme;
    AVX-512 / AVX2 picked at run time) when the oracle has a batched form of the level, else the scalar loop."""
    if op.vmult_simd(dst, src):
        return lambda: op.vmult_simd(dst, src)
    op.vmult(dst, src)
    return lambda: op.vmult(dst, src)


def cpu_arm_description():
    from oracle import c_oracle as CO

    return f"oracle/mf_oracle.c, OpenMP over cell batches, SIMD over cells ({CO.simd_isa()}), {CO.num_threads()} threads"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for every worker; the CPU arm is one process that should use all host cores
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or os.environ.get("TORCHELASTIC_RUN_ID"):
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    p = args.degree
    nc = args.cpu_cells
    times = []
    from b200mf.mesh import StructuredMesh, synthetic_displacement
    from oracle import c_oracle as CO

    m = StructuredMesh((nc,) * 3, p)
    m.clamp([(0, 0)], 3)
    invjac, jxw = CO.q1_metric(m.vertices, p + 1)
    cells = CO.Cells(p, p + 1, 3, m.dof_index(3), m.constrained_dofs(3), invjac, jxw)
    mat = {k: v for k, v in MAT.items() if k != "formulation"}
    op = CO.NeoHookOperator(cells, formulation=1, **mat)
    op.initialize(synthetic_displacement(m.node_coordinates(), 0.05), "tensor2")
    op.cache()
    n = 3 * m.n_nodes
    src = np.random.default_rng(0).uniform(-1, 1, n)
    dst = np.zeros(n)
    apply = cpu_apply(op, dst, src)
    for _ in range(args.warmup):
        apply()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        apply()
    el = time.perf_counter() - t0
    val = n * args.steps / el
    sample = (f"Q{p} box {nc}^3 cells = {n} DoFs per step (bounded sample of the {CELLS_PER_DIR[p]}^3-cell workload: the full mesh "
              f"needs ~16 GB of host tables and minutes of set-up); {cpu_arm_description()}")
    cfg = workload_config(p, CELLS_PER_DIR[p] if args.cells is None else args.cells, args.gpus)
    cfg["workload"] += f" [CPU arm measured on a {nc}^3-cell sample of it]"
    alg_gbs = val * algorithmic_bytes_per_cell(p) / (3 * p**3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "DoFs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "DoFs/s", "cores": CO.num_threads(), "kind": "port", "sample": sample,
                         "algorithmic_GB/s": alg_gbs, "host_stream_triad_GB/s": CO.host_stream_triad_gbs()},
        "e2e": {"value": val, "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference needs deal.II>=9.6 + preCICE>=3 (absent): CPU arm = oracle/mf_oracle.c port, vectorised over cells like "
                "deal.II's MatrixFree (VectorizedArray), OpenMP over cell batches",
    }
    print(json.dumps(line), flush=True)


def workload_config(p, nc, ngpus, caching="tensor2", formulation=1):
    return {"workload": f"neo-Hookean tangent vmult, {caching} caching, formulation {formulation}, Q{p}, box {nc}^3 cells per GPU",
            "degree": p, "cells_per_gpu": nc**3, "dofs_per_gpu": 3 * (nc * p + 1) ** 3, "caching": caching,
            "parallelism": f"dd{ngpus}", "l2": "vectors+tables per step >> 126 MB L2 (no flush needed)"}


def roofline_of(p, es, variant, n_cells, kernel_ms, step_ms, peak):
    alg = algorithmic_bytes_per_cell(p, es, variant) * n_cells
    return alg, alg / (kernel_ms * 1e-3) / 1e9 / peak, alg / (step_ms * 1e-3) / 1e9 / peak


def timed_vmults(torch, op, mf, src, dst, steps, stream, barrier):
    """`steps` vmults between CUDA events on `stream`; the library brackets every cell-loop launch with its own event pair
    (b200mf_ctx_time_kernels) -> (ms per step, cell-loop kernel ms per step, launches timed, per-step ms)."""
    mf.time_kernels(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    # an untimed ~2 ms spin kernel in front of the first event lets the host run ahead, so that the timed region
    # measures K steps of device work and not the host latency of enqueueing the first one after the barrier
    with torch.cuda.stream(stream):
        torch.cuda._sleep(4_000_000)
    ev[0].record(stream)
    for i in range(steps):
        op.vmult(dst, src)
        ev[i + 1].record(stream)
    barrier()
    kernel_ms_total, kernel_launches = mf.kernel_time()
    mf.time_kernels(False)
    per_step = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    return ev[0].elapsed_time(ev[-1]) / steps, kernel_ms_total / steps, int(kernel_launches), per_step


def graphed_vmults(torch, op, src, dst, steps, stream, barrier):
    """The same `steps` vmults captured once into a CUDA graph and replayed inside the timed region: with several ranks
    per host the Python/driver enqueue path of a rank is what occasionally starves its GPU (one 3.7 ms step among 1.57 ms
    ones at 8 ranks in round 1); a graph replay takes the host out of the timed region.  Returns ms per step or None if the
    capture is not possible."""
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
            for _ in range(steps):
                op.vmult(dst, src)
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"[bench] CUDA graph capture failed ({e}); timing the plain loop\n")
        return None
    with torch.cuda.stream(stream):
        g.replay()  # warm-up replay
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        torch.cuda._sleep(4_000_000)
        e0.record(stream)
        g.replay()
        e1.record(stream)
    barrier()
    return e0.elapsed_time(e1) / steps


def degree_line(torch, q, nc, dtype, caching, formulation, variant, peak, stream_ptr, numbering="lexicographic", curved=False, steps=10):
    """One (degree, mesh flavour) point of the metric on one GPU: device-resident DoFs/s, kernel-alone and whole-step
    fraction of the measured HBM bandwidth."""
    from b200mf import DeviceVector

    es = 8 if dtype == "f64" else 4
    mq, mfq, opq = build_problem(q, nc, dtype, None, caching, formulation, numbering, curved)
    mfq.set_stream(stream_ptr)
    nq = 3 * mq.n_nodes
    sh = np.random.default_rng(q).uniform(-1, 1, nq)
    sh[mq.constrained_dofs(3)] = 0.0
    s, d = DeviceVector(mfq, sh), DeviceVector(mfq)
    for _ in range(3):
        opq.vmult(d, s)
    torch.cuda.synchronize()
    ms, kms, _, _ = timed_vmults(torch, opq, mfq, s, d, steps, torch.cuda.current_stream(), torch.cuda.synchronize)
    alg, frac_k, frac_s = roofline_of(q, es, variant, mq.n_cells, kms, ms, peak)
    out = {"dofs": nq, "cells_per_dir": nc, "ms_per_step": ms, "kernel_ms": kms, "dofs_per_s": nq / (ms * 1e-3),
           "GB/s_kernel": alg / (kms * 1e-3) / 1e9, "frac_kernel": frac_k, "frac_step": frac_s,
           "bytes_per_dof": algorithmic_bytes_per_cell(q, es, variant) / (3 * q**3)}
    del s, d, opq, mfq, mq
    return out


def parity_block(torch, dist, comm, world, rank, p):
    """Checker, not product: the (distributed) vmult on a reduced mesh of the same brick grid against the C oracle
    (oracle/mf_oracle.c) applied to the undivided global mesh, relative l2 over all owned rows + two global checksums."""
    from b200mf import DeviceVector, MatrixFree, NeoHookOperator
    from b200mf.mesh import StructuredMesh, synthetic_displacement
    from b200mf.partition import BrickPartition, weak_scaling_grid

    grid = weak_scaling_grid(world)
    nc = 5 if p <= 3 else 3
    if world > 1:
        pt = BrickPartition(grid, nc, p, rank)
        pt.clamp_global_face(0, 0)
        mf = MatrixFree.from_partition(pt, comm)
        glob = pt.global_lex_node()
        n_owned, n_loc = pt.n_owned, pt.n_dofs
        coords = pt.node_coordinates()
    else:
        pt = StructuredMesh((nc,) * 3, p)
        pt.clamp([(0, 0)], 3)
        mf = MatrixFree.from_mesh(pt, 3)
        glob = np.arange(pt.n_nodes)
        n_owned = n_loc = 3 * pt.n_nodes
        coords = pt.node_coordinates()
    gdof = (3 * glob[:, None] + np.arange(3)[None, :]).reshape(-1)
    n_glob_nodes = int(np.prod([g * nc * p + 1 for g in grid]))
    gsrc = np.random.default_rng(7).uniform(-1, 1, 3 * n_glob_nodes)  # the same global vector on every rank
    op = NeoHookOperator()
    op.set_material(mf, dict(MAT))
    op.initialize(DeviceVector(mf, synthetic_displacement(coords, 0.05)), "tensor2")
    op.cache()
    v = np.zeros(n_loc)
    v[:n_owned] = gsrc[gdof[:n_owned]]
    con = pt.constrained_dofs() if world > 1 else pt.constrained_dofs(3)
    s, d = DeviceVector(mf, v), DeviceVector(mf)
    op.vmult(d, s)
    out = d.download()[:n_owned]
    box = [None] * world
    if dist is not None:
        dist.all_gather_object(box, (gdof[:n_owned], out))
    else:
        box = [(gdof[:n_owned], out)]
    if rank != 0:
        return None
    from oracle import c_oracle as CO

    m = StructuredMesh(tuple(g * nc for g in grid), p, lower=(0, 0, 0), upper=grid)
    m.clamp([(0, 0)], 3)
    invjac, jxw = CO.q1_metric(m.vertices, p + 1)
    mat = {k: v_ for k, v_ in MAT.items() if k != "formulation"}
    ref_op = CO.NeoHookOperator(CO.Cells(p, p + 1, 3, m.dof_index(3), m.constrained_dofs(3), invjac, jxw), formulation=1, **mat)
    ref_op.initialize(synthetic_displacement(m.node_coordinates(), 0.05), "tensor2")
    ref_op.cache()
    ref = np.zeros_like(gsrc)
    ref_op.vmult(ref, gsrc)
    got = np.full_like(gsrc, np.nan)
    for idx, val in box:
        got[idx] = val
    assert not np.isnan(got).any(), "the owned rows of all ranks do not cover the global vector"
    return {"rel_l2": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)), "norm_dst": float(np.linalg.norm(got)),
            "norm_dst_oracle": float(np.linalg.norm(ref)), "dst_dot_src": float(got @ gsrc), "dst_dot_src_oracle": float(ref @ gsrc),
            "mesh": f"Q{p}, {grid[0]}x{grid[1]}x{grid[2]} bricks of {nc}^3 cells, {gsrc.size} DoFs, tensor2",
            "checker": "oracle/mf_oracle.c on the undivided global mesh (test infrastructure; not on the measured path)"}


def tube_block(stream_ptr, degree=3, levels=5, pressure=5.0e3):
    """BASELINE.json configs[3] on ONE GPU: a Newmark step of the solid solver on the tube3d geometry (tube_3d.h:23-33, 74:
    cylinder shell, curved cells -> full per-q-point metric; ends clamped; pressure-like traction on the inner wall pulled back
    in every Newton iteration) with the reference's whole 35-coarse-cell multigrid hierarchy; every Newton iteration =
    residual (cells + interface faces) + update_matrix_free + GMG-CG.  (The 8-GPU, 4e8-DoF size of the config needs the
    coarse levels agglomerated below 8 axial cells per rank, which the brick partitioner does not do yet.)"""
    from b200mf import lib as L_
    from b200mf.mesh import tube_hierarchy, tube_pressure_traction
    from b200mf.solid import Solid

    meshes = tube_hierarchy(degree, levels)
    sol = Solid(meshes, mu=115384.0, nu=0.3, rho=1200.0, delta_t=1e-4, n_cheb_cg=30,  # examples/tube3d/solid/tube3d.prm
                traction=tube_pressure_traction(pressure), traction_faces=((2, 0),))
    for mf in [sol.mf] + sol.level_mf:
        mf.set_stream(stream_ptr)
    m = meshes[-1]
    sol.mf.synchronize()
    steps = []
    for _ in range(2):  # the first step from rest also creates the smoother / solver objects; the second one is the steady state
        n_solves, launches0 = len(sol.cg_history), L_.load().b200mf_launch_count()
        t0 = time.perf_counter()
        its = sol.solve_nonlinear_timestep()
        sol.advance_time_window()
        sol.mf.synchronize()
        t1 = time.perf_counter()
        steps.append({"time_step_s": t1 - t0, "newton_iterations": int(its), "cg_iterations": list(map(int, sol.cg_history[n_solves:])),
                      "gpu_launches": int(L_.load().b200mf_launch_count() - launches0)})
    last = steps[-1]
    return {"workload": f"tube3d Newmark step: neo-Hookean, tensor2, Q{degree}, cylinder shell {m.ncells[0]}x{m.ncells[1]}x{m.ncells[2]} curved cells, "
                        f"{3 * m.n_nodes} DoFs, {levels} levels, GMG-Chebyshev CG, interface traction on {sol.face_op.cells.size} faces",
            "n_dofs": 3 * m.n_nodes, "time_step_s": last["time_step_s"], "newton_iterations": last["newton_iterations"],
            "cg_iterations": last["cg_iterations"],
            "seconds_per_newton_iteration": last["time_step_s"] / max(1, len(last["cg_iterations"])),
            "gpu_launches": last["gpu_launches"], "first_step_from_rest": steps[0], "mg_graph": mg_graph_note(sol.mg),
            "max_displacement": float(np.abs(sol.total_displacement()).max()),
            "note": "second time step (steady state); first_step_from_rest additionally creates the smoother / solver objects"}


def mg_graph_note(mg):
    """How the V-cycle's lower levels ran: kernels replayed from the recorded CUDA graph count in `gpu_launches` like single launches."""
    if mg is None:
        return None
    gi = mg.graph_info()
    gi["note"] = "levels 0..top_level of every V-cycle are one cudaGraphLaunch of launches_per_replay kernels; state 1 = replaying"
    return gi


def newton_block(stream_ptr, degree=3, coarse=5, levels=5, amplitude=0.01):
    """BASELINE.json metric, second half: Newton-step time = update_matrix_free (cache + diagonals + interpolate_to_mg +
    Chebyshev eigenvalue estimates on every level) + one GMG-CG solve to 1e-6 * ||rhs|| (mf_elasticity.h:1289-1353,
    1981-2070), device-resident, wall clock between stream synchronisations."""
    from b200mf import DeviceVector, lib as L_
    from b200mf.mesh import StructuredMesh, synthetic_displacement
    from b200mf.solid import Solid

    meshes = []
    for l in range(levels):
        m = StructuredMesh((coarse * 2**l,) * 3, degree)
        m.clamp([(0, 0)], 3)
        meshes.append(m)
    mat = {k: v for k, v in MAT.items() if k not in ("formulation", "alpha_1")}
    sol = Solid(meshes, delta_t=0.01, **mat)
    sol.mf.set_stream(stream_ptr)
    for mf in sol.level_mf:
        mf.set_stream(stream_ptr)
    m = meshes[-1]
    sol.total.upload(synthetic_displacement(m.node_coordinates(), amplitude))
    b = np.random.default_rng(0).uniform(-1, 1, 3 * m.n_nodes)
    b[m.constrained_dofs(3)] = 0.0
    sol.rhs.upload(b)
    nrm = sol.rhs.l2_norm()
    sol.update_matrix_free()          # warm-up: allocations, first launches
    sol.solve_linear_system(nrm)
    sol.mf.synchronize()
    launches0 = L_.load().b200mf_launch_count()
    t0 = time.perf_counter()
    sol.update_matrix_free()
    sol.mf.synchronize()
    t1 = time.perf_counter()
    its = sol.solve_linear_system(nrm)
    sol.mf.synchronize()
    t2 = time.perf_counter()
    return {"workload": f"neo-Hookean Newton step (update_matrix_free + GMG-CG to 1e-6), tensor2, Q{degree}, box {m.ncells[0]}^3 cells, "
                        f"{3 * m.n_nodes} DoFs, {levels} levels (FP32 levels under FP64 CG)",
            "n_dofs": 3 * m.n_nodes, "newton_step_s": t2 - t0, "update_matrix_free_s": t1 - t0, "cg_solve_s": t2 - t1,
            "cg_iterations": int(its), "gpu_launches": int(L_.load().b200mf_launch_count() - launches0),
            "level0": sol.eig[0], "level_top": sol.eig[-1], "mg_graph": mg_graph_note(sol.mg)}


FP64_PEAK_TFLOPS = 33.9  # measured on this pool's B200: independent DFMA chains, tools/microbench/fp64_pipes.cu -> profiles/r01_fp64_pipes.jsonl


def heat_block(torch, stream_ptr, degree=4, coarse=7, levels=5, dt=0.1):
    """BASELINE.json configs[1]: backward-Euler heat step, Q4 cube, ~9e7 DoF, GMG-CG on one B200 (heat_transfer.h:577-812)
    through b200mf.heat.HeatProblem, plus the FP64 system operator alone: on affine cells it streams ~41 B/DoF and does
    ~293 flop/DoF, i.e. it is bound by the FP64 pipe, reported against the measured DFMA peak."""
    from b200mf.heat import HeatProblem
    from b200mf.mesh import StructuredMesh

    faces = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)]
    meshes = []
    for l in range(levels):
        m = StructuredMesh((coarse * 2**l,) * 3, degree)
        m.clamp(faces, 1)
        meshes.append(m)
    hp = HeatProblem(meshes, dt, source_const=1.0)
    for mf in [hp.mf] + hp.level_mf:
        mf.set_stream(stream_ptr)
    m = meshes[-1]
    x = m.node_coordinates()
    u0 = np.prod(np.sin(np.pi * x), axis=1)
    u0[m.constrained_dofs(1)] = 0.0
    hp.solution.upload(u0)
    hp.time_step()  # warm-up: allocations, smoother objects, first launches
    hp.mf.synchronize()
    t0 = time.perf_counter()
    hp.solution_old.sadd(0.0, 1.0, hp.solution)
    hp.assemble_rhs()
    hp.mf.synchronize()
    t1 = time.perf_counter()
    hp.setup_preconditioner()
    hp.mf.synchronize()
    t2 = time.perf_counter()
    res = hp.solve()
    hp.mf.synchronize()
    t3 = time.perf_counter()
    # the FP64 system operator alone
    n = m.n_nodes
    y = hp.mf.initialize_dof_vector()
    for _ in range(3):
        hp.system_matrix.vmult(y, hp.solution)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        hp.system_matrix.vmult(y, hp.solution)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nn, p = degree + 1, degree
    flop_per_dof = (24 * nn**4 + 30 * nn**3) / p**3
    tflops = n / (ms * 1e-3) * flop_per_dof / 1e12
    return {"workload": f"heat: backward-Euler (M + dt K), Q{degree}, cube {m.ncells[0]}^3 affine cells, {n} DoFs, {levels} levels, dt={dt}",
            "n_dofs": n, "time_step_s": t3 - t0, "assemble_rhs_s": t1 - t0, "diagonals_and_eigenvalues_s": t2 - t1,
            "cg_solve_s": t3 - t2, "cg_iterations": int(res["iterations"]), "residual": res["residual"],
            "mg_graph": mg_graph_note(getattr(hp, "mg", None)),
            "system_vmult": {"ms": ms, "dofs_per_s": n / (ms * 1e-3),
                             "roofline": {"bound": "fp64", "achieved": tflops, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                                          "frac": tflops / FP64_PEAK_TFLOPS, "flop_per_dof": flop_per_dof,
                                          "peak_source": "measured DFMA peak, profiles/r01_fp64_pipes.jsonl",
                                          "kernel": f"scalar_stack_kernel<double, {nn}, LaplaceAffineOp>"}}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--degree", type=int, default=3)
    ap.add_argument("--cells", type=int, default=None, help="cells per direction per GPU (default: ~5e7 DoF)")
    ap.add_argument("--cpu-cells", type=int, default=None, help="cells per direction of the CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="N > 1: time the plain Python loop instead of a CUDA-graph replay of it")
    ap.add_argument("--no-heat", action="store_true", help="skip the heat_step block (BASELINE configs[1])")
    ap.add_argument("--no-extras", action="store_true", help="skip the sweep / numbering / newton_step / parity blocks (tuning runs)")
    ap.add_argument("--sweep", action="store_true", help="report Q1..Q6 in the sweep block (default: Q2, Q3, Q4 = the metric)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--caching", default="tensor2", choices=["scalar_referential", "tensor2", "tensor4", "tensor4_ns"])
    ap.add_argument("--formulation", type=int, default=1, choices=[0, 1])
    args = ap.parse_args()
    p = args.degree
    if args.cpu_cells is None:
        args.cpu_cells = {1: 128, 2: 64, 3: 48, 4: 36, 5: 28, 6: 24}[p]
    if args.impl == "reference":
        return run_reference(args)

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import b200mf
    from b200mf import DeviceVector

    L = b200mf.load()
    nc = CELLS_PER_DIR[p] if args.cells is None else args.cells
    es = 8 if args.dtype == "f64" else 4
    comm = b200mf.Communicator.from_torch_distributed() if world > 1 else None
    variant = "form0" if args.formulation == 0 else args.caching
    m, mf, op = build_problem(p, nc, args.dtype, comm, args.caching, args.formulation)
    n = 3 * m.n_nodes  # local vector: owned + ghost entries
    n_owned = m.n_owned if world > 1 else n
    src_h = np.random.default_rng(rank).uniform(-1, 1, n)
    src_h[m.constrained_dofs(3) if world == 1 else m.constrained_dofs()] = 0.0
    src, dst = DeviceVector(mf, src_h), DeviceVector(mf)
    # N > 1 runs on a side stream (the legacy default stream cannot be captured into a CUDA graph)
    stream = torch.cuda.Stream() if world > 1 else torch.cuda.current_stream()
    mf.set_stream(stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:  # connection set-up: NCCL opens its peer channels lazily during the first exchanges
        for _ in range(10):
            op.vmult(dst, src)
        barrier()
    for _ in range(max(args.warmup, 3)):
        op.vmult(dst, src)
    barrier()
    launches0 = L.b200mf_launch_count()
    with ClockSampler(local_rank) as clocks:
        ms_rank, kernel_ms, kernel_launches, step_ms = timed_vmults(torch, op, mf, src, dst, args.steps, stream, barrier)
    launches = L.b200mf_launch_count() - launches0
    graph_ms = None
    if world > 1 and not args.no_graph:
        graph_ms = graphed_vmults(torch, op, src, dst, args.steps, stream, barrier)
        flags = torch.tensor([0.0 if graph_ms is None else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)  # all ranks or none
        if float(flags.item()) == 0.0:
            graph_ms = None
    loop_ms_rank = ms_rank
    use_graph = False
    if graph_ms is not None:
        # both are K vmults of the same library calls; report the faster execution mode (decided on the max over ranks): the
        # graph wins when NCCL's proxy thread jitters (35 ms outlier steps at 8 ranks), the plain loop with the p2p exchange
        t2 = torch.tensor([graph_ms, ms_rank], device="cuda", dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        use_graph = float(t2[0].item()) < float(t2[1].item())
        if use_graph:
            ms_rank = graph_ms
    total_ms = ms_rank * args.steps
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms = total_ms / args.steps
    dofs_total = n_owned  # global DoFs = sum of the owned entries of all ranks (weak scaling: one brick per rank)
    if dist is not None:
        t = torch.tensor([n_owned], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        dofs_total = int(t.item())
    value = dofs_total / (ms * 1e-3)

    # ---- roofline of the dominant kernel (the cell loop; memset + constrained fix-up are the rest of the step)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes, frac_kernel, frac_step = roofline_of(p, es, variant, m.n_cells, kernel_ms, ms, peak)
    traffic, traffic_src = ncu_traffic(p, variant, args.dtype, args.cells is None)
    roofline = {"bound": "hbm", "kernel": f"cell_loop_kernel<{'double' if es == 8 else 'float'}, {p + 1}, NHOp<{variant}>>",
                "achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": frac_kernel,
                "traffic": traffic, "traffic_source": (f"ncu --set full capture committed as {traffic_src} (dram__bytes_read.sum + "
                                                       "dram__bytes_write.sum of one launch of this workload; not measured in this run)") if traffic_src else None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                "algorithmic_bytes_per_launch": alg_bytes,
                "algorithmic_bytes_per_dof": algorithmic_bytes_per_cell(p, es, variant) / (3 * p**3),
                "kernel_ms": kernel_ms, "kernel_launches_timed": int(kernel_launches), "kernel_share_of_step": kernel_ms / loop_ms_rank,
                "frac_whole_step": frac_step,
                "timed": "cell-loop kernel alone, CUDA events on its stream inside the timed region "
                         f"({kernel_launches // max(args.steps, 1)} launch(es) per step)"}

    mf_p2p = bool(world > 1 and mf.exchange_is_p2p())
    # ---- run-to-run reproducibility at bench size: the scatter is an atomic add (red.global.add.f64), so the summation
    # order differs between launches; report the spread of two applications to the same vector
    dst2 = DeviceVector(mf)
    op.vmult(dst, src)
    op.vmult(dst2, src)
    a, b = dst.download()[:n_owned], dst2.download()[:n_owned]
    run_to_run = {"max_abs_diff_over_max_abs": float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-300)),
                  "rel_l2_diff": float(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-300)),
                  "note": "two vmults of the same vector; differences come from the order of the atomic scatter-adds"}
    del dst2, a, b

    # ---- N > 1: how much of the two ghost exchanges is exposed -----------
    exchange = None
    if world > 1:
        def timed(mode, k=10):
            mf.set_overlap(mode)
            for _ in range(3):
                op.vmult(dst, src)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(k):
                op.vmult(dst, src)
            e1.record(stream)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / k], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        t_serial, t_overlap, t_none = timed(0), timed(1), timed(-1)
        mf.set_overlap(1 if mf_p2p else 0)  # library default: overlapped with the peer-to-peer exchange, serial with NCCL
        exchange = {"ms_step_default": ms, "ms_overlapped_exchange": t_overlap, "ms_serial_exchange": t_serial,
                    "ms_no_exchange": t_none, "exposed_ms_overlapped": t_overlap - t_none, "exposed_ms_serial": t_serial - t_none,
                    "note": "serial = the reference's order update_ghost_values -> cells -> compress(add); overlapped = boundary cells "
                            "first, exchanges on a second stream behind the interior cells; the timed step uses the library default "
                            "(overlapped with the peer-to-peer exchange, serial with NCCL); "
                            "no_exchange = rank-local cells only (measurement mode); 10 steps each after 3 warm-ups"}

    # ---- e2e: host buffers through the C-ABI host entry point ----------------------------------
    e2e = None
    if not args.no_e2e and args.dtype == "f64":
        hs = torch.from_numpy(src_h).pin_memory()
        hd = torch.empty(n, dtype=torch.float64).pin_memory()
        import ctypes as C

        for _ in range(2):
            b200mf.lib.check(L.b200mf_nh_vmult_host(op.h, C.c_void_p(hd.data_ptr()), C.c_void_p(hs.data_ptr())))
        barrier()
        k = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k):
            b200mf.lib.check(L.b200mf_nh_vmult_host(op.h, C.c_void_p(hd.data_ptr()), C.c_void_p(hs.data_ptr())))
        barrier()
        el = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e = {"value": dofs_total * k / el, "unit": "DoFs/s", "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
               "steps": k, "ms_per_step": 1e3 * el / k}
        del hs, hd

    # ---- parity of this very configuration (rank grid, exchange, kernel) on a reduced mesh against the C oracle
    parity = None
    if not args.no_extras and args.dtype == "f64":
        parity = parity_block(torch, dist, comm, world, rank, p)

    # ---- the other points of BASELINE.json's metric: Q2..Q4 (Q1..Q6 with --sweep), numbering / geometry flavours,
    # Newton-step time.  One GPU (rank 0 keeps its device; the other ranks wait at the final barrier).
    sweep = numbering = newton = heat = tube = None
    if not args.no_extras and world == 1:
        del src, dst, op, mf, m
        sp = stream.cuda_stream
        sweep = {}
        for q in (range(1, 7) if args.sweep else (2, 3, 4)):
            if q == p and args.cells is None:
                sweep[f"Q{q}"] = {"dofs": n, "cells_per_dir": nc, "ms_per_step": ms, "kernel_ms": kernel_ms, "dofs_per_s": value,
                                  "GB/s_kernel": roofline["achieved"], "frac_kernel": frac_kernel, "frac_step": frac_step,
                                  "bytes_per_dof": roofline["algorithmic_bytes_per_dof"]}
            else:
                sweep[f"Q{q}"] = degree_line(torch, q, CELLS_PER_DIR[q], args.dtype, args.caching, args.formulation, variant, peak, sp)
        numbering = {"lexicographic_box": sweep.get(f"Q{p}"),
                     "cellwise_morton_box": degree_line(torch, p, nc, args.dtype, args.caching, args.formulation, variant, peak, sp,
                                                        numbering="cellwise_morton"),
                     "curved_cells_lexicographic": degree_line(torch, p, nc, args.dtype, args.caching, args.formulation, variant, peak, sp,
                                                               curved=True),
                     "note": "same Q%d workload; cellwise_morton = cells in z-curve order, nodes numbered cell by cell in first-touch "
                             "order (deal.II-like, mf_elasticity.h:416-425, 667-669); curved = smooth non-affine vertex mapping" % p}
        if args.dtype == "f64" and args.caching == "tensor2":
            # the multigrid level operator (LevelNumber = float, mf_elasticity.h:363-367): 10 of its applications per V-cycle are
            # about half of the Newton step below
            sweep["Q3_fp32_level_operator"] = degree_line(torch, 3, CELLS_PER_DIR[3], "f32", args.caching, args.formulation, variant, peak, sp)
            newton = newton_block(sp)
            if not args.no_heat:
                heat = heat_block(torch, sp)
                tube = tube_block(sp)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, ncpu, reps, threads, sec = cpu_oracle_rate(p, args.cpu_cells)
        cpu = {"value": v, "unit": "DoFs/s", "cores": threads, "kind": "port",
               "sample": f"Q{p} box {args.cpu_cells}^3 cells = {ncpu} DoFs x {reps} vmults ({sec:.2f} s each); {cpu_arm_description()}",
               "algorithmic_GB/s": v * algorithmic_bytes_per_cell(p) / (3 * p**3) / 1e9,
               "host_stream_triad_GB/s": host_stream_gbs()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "DoFs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "config": workload_config(p, nc, world, args.caching, args.formulation), "roofline": roofline, "cpu_baseline": cpu,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(),
            "step_ms_rank0": {"min": step_ms[0], "median": step_ms[len(step_ms) // 2], "max": step_ms[-1]},
            "run_to_run": run_to_run,
        }
        if world > 1:
            line["timed_region"] = ("CUDA-graph replay of the K vmults (captured once, untimed)" if use_graph
                                    else "plain loop of K library calls") + "; step_ms_rank0 / kernel_ms are from the plain loop"
            line["ms_per_step_plain_loop_rank0"] = loop_ms_rank
            line["ms_per_step_graph_replay_rank0"] = graph_ms
            line["exchange_path"] = "p2p (CUDA IPC push/pull kernels over NVLink)" if mf_p2p else "nccl send/recv"
        if parity:
            line["parity"] = parity
        if exchange:
            line["ghost_exchange"] = exchange
        if sweep:
            line["sweep"] = sweep
        if numbering:
            line["numbering"] = numbering
        if newton:
            line["newton_step"] = newton
        if heat:
            line["heat_step"] = heat
        if tube:
            line["tube3d_step"] = tube
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
