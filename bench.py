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


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the cell-loop kernel from the committed
# `ncu --set full` captures (profiles/r01_v3_ncu_summary.md), keyed by (degree, variant, dtype) at the default mesh
NCU_TRAFFIC = {(2, "tensor2", "f64"): 9.95e9, (3, "tensor2", "f64"): 7.46e9, (4, "tensor2", "f64"): 6.81e9,
               (3, "tensor2", "f32"): 3.63e9}


def algorithmic_bytes_per_cell(p, esize=8, variant="tensor2"):
    """src read + dst write per unique DoF, one uint32 per node, and the q-point
    table the kernel of this variant actually streams."""
    n = p + 1
    return 2 * esize * 3 * p**3 + 4 * n**3 + QPOINT_VALUES[variant] * esize * n**3


def build_problem(p, ncell_dir, dtype="f64", comm=None, caching="tensor2", formulation=1):
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
        m = StructuredMesh((ncell_dir,) * 3, p)
        m.clamp([(0, 0)], 3)
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
    op.vmult(dst, src)  # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        op.vmult(dst, src)
        reps += 1
        el = time.perf_counter() - t0
        if el > min_seconds or reps >= max_reps:
            break
    return n * reps / el, n, reps, CO.num_threads(), el / reps


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
    for _ in range(args.warmup):
        op.vmult(dst, src)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        op.vmult(dst, src)
    el = time.perf_counter() - t0
    val = n * args.steps / el
    sample = f"Q{p} box {nc}^3 cells = {n} DoFs per step (bounded sample of the {CELLS_PER_DIR[p]}^3-cell workload)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "DoFs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(p, CELLS_PER_DIR[p] if args.cells is None else args.cells, args.gpus),
        "cpu_baseline": {"value": val, "unit": "DoFs/s", "cores": CO.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference needs deal.II>=9.6 + preCICE>=3 (absent): CPU arm = oracle/mf_oracle.c port, OpenMP over cells",
    }
    print(json.dumps(line), flush=True)


def workload_config(p, nc, ngpus, caching="tensor2", formulation=1):
    return {"workload": f"neo-Hookean tangent vmult, {caching} caching, formulation {formulation}, Q{p}, box {nc}^3 cells per GPU",
            "degree": p, "cells_per_gpu": nc**3, "dofs_per_gpu": 3 * (nc * p + 1) ** 3, "caching": caching,
            "parallelism": f"dd{ngpus}", "l2": "vectors+tables per step >> 126 MB L2 (no flush needed)"}


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
    ap.add_argument("--sweep", action="store_true", help="also report Q1..Q6 device-resident rates")
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
    stream = torch.cuda.current_stream()
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
    mf.time_kernels(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        # an untimed ~2 ms spin kernel in front of the first event lets the host run ahead, so that the timed region
        # measures K steps of device work and not the host latency of enqueueing the first one after the barrier
        torch.cuda._sleep(4_000_000)
        ev[0].record(stream)
        for i in range(args.steps):
            op.vmult(dst, src)
            ev[i + 1].record(stream)
        barrier()
    launches = L.b200mf_launch_count() - launches0
    kernel_ms_total, kernel_launches = mf.kernel_time()
    mf.time_kernels(False)
    kernel_ms = kernel_ms_total / args.steps  # all cell-loop launches of one step (1 on one rank; interior/boundary/interior when partitioned)
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))  # this rank's per-step times (diagnostic)
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
    alg_bytes = algorithmic_bytes_per_cell(p, es, variant) * m.n_cells  # per GPU (per launch of the cell kernel)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": f"cell_loop_kernel<{'double' if es == 8 else 'float'}, {p + 1}, NHOp<{variant}>>",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC.get((p, variant, args.dtype)) if args.cells is None else None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                "algorithmic_bytes_per_launch": alg_bytes,
                "algorithmic_bytes_per_dof": algorithmic_bytes_per_cell(p, es, variant) / (3 * p**3),
                "kernel_ms": kernel_ms, "kernel_launches_timed": int(kernel_launches), "kernel_share_of_step": kernel_ms / ms,
                "frac_whole_step": alg_bytes / (ms * 1e-3) / 1e9 / peak,
                "timed": "cell-loop kernel alone, CUDA events on its stream inside the timed region "
                         f"({kernel_launches // max(args.steps, 1)} launch(es) per step)"}

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
        mf.set_overlap(0)  # library default
        exchange = {"ms_step_default": ms, "ms_overlapped_exchange": t_overlap, "ms_serial_exchange": t_serial,
                    "ms_no_exchange": t_none, "exposed_ms_overlapped": t_overlap - t_none, "exposed_ms_serial": t_serial - t_none,
                    "note": "default (timed step) = serial = the reference's order update_ghost_values -> cells -> compress(add); "
                            "overlapped = boundary cells first, exchanges on a second stream behind the interior cells (opt-in); "
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

    sweep = None
    if args.sweep and rank == 0:
        sweep = {}
        del src, dst, op, mf
        for q in range(1, 7):
            mq, mfq, opq = build_problem(q, CELLS_PER_DIR[q], args.dtype, None, args.caching, args.formulation)
            nq = 3 * mq.n_nodes
            s, d = DeviceVector(mfq, np.random.default_rng(q).uniform(-1, 1, nq)), DeviceVector(mfq)
            for _ in range(3):
                opq.vmult(d, s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                opq.vmult(d, s)
            e1.record()
            torch.cuda.synchronize()
            msq = e0.elapsed_time(e1) / 10
            gb = algorithmic_bytes_per_cell(q, es, variant) * mq.n_cells / (msq * 1e-3) / 1e9
            sweep[f"Q{q}"] = {"dofs": nq, "ms": msq, "dofs_per_s": nq / (msq * 1e-3), "GB/s": gb, "frac": gb / peak}
            del s, d, opq, mfq

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, ncpu, reps, threads, sec = cpu_oracle_rate(p, args.cpu_cells)
        cpu = {"value": v, "unit": "DoFs/s", "cores": threads, "kind": "port",
               "sample": f"Q{p} box {args.cpu_cells}^3 cells = {ncpu} DoFs x {reps} vmults ({sec:.2f} s each), oracle/mf_oracle.c"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "DoFs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "config": workload_config(p, nc, world, args.caching, args.formulation), "roofline": roofline, "cpu_baseline": cpu,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(),
            "step_ms_rank0": {"min": step_ms[0], "median": step_ms[len(step_ms) // 2], "max": step_ms[-1]},
        }
        if exchange:
            line["ghost_exchange"] = exchange
        if sweep:
            line["sweep"] = sweep
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
