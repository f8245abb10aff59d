#!/usr/bin/env python
"""Solver-level measurements on one B200 (BASELINE.json metric part 2 "Newton-step time", configs[1] and a one-GPU
version of configs[3]); prints one JSON line per run.

  python tools/bench_solver.py heat   --degree 4 --coarse 7 --levels 5      # heat: GMG-CG on (M + dt K), cube, ~9e7 DoF
  python tools/bench_solver.py newton --degree 3 --coarse 5 --levels 5      # solid: update_matrix_free + GMG-CG solve

"newton" = what the reference does per Newton iteration after the residual: cache() + compute_diagonal() of the FP64
operator, interpolate_to_mg of the displacement, cache() + compute_diagonal() of every FP32 level operator, Chebyshev
eigenvalue re-estimation on every level (setup_gmg(false)), then SolverCG + PreconditionMG to tol_lin * ||rhs||
(mf_elasticity.h:1289-1353, 1981-2070).  All device-resident; wall clock with a stream synchronisation on both sides.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "matrix-free-dealii-precice_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

MAT = dict(mu=0.5e6, nu=0.4, rho=1000.0, alpha_1=1.0 / (0.25 * 0.01**2), formulation=1)
ALL_FACES = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)]


class Timer:
    def __init__(self, mf):
        self.mf, self.t = mf, {}

    def __call__(self, name):
        return _Section(self, name)


class _Section:
    def __init__(self, timer, name):
        self.timer, self.name = timer, name

    def __enter__(self):
        self.timer.mf.synchronize()
        self.t0 = time.perf_counter()

    def __exit__(self, *a):
        self.timer.mf.synchronize()
        self.timer.t[self.name] = self.timer.t.get(self.name, 0.0) + time.perf_counter() - self.t0


def run_heat(args):
    from b200mf import DeviceVector, LaplaceOperator, MatrixFree, MGTransferMatrixFree, PreconditionChebyshev, PreconditionMG, SolverCG
    from b200mf import lib as L
    from b200mf.mesh import StructuredMesh, children_of

    p = args.degree
    meshes = []
    for l in range(args.levels):
        m = StructuredMesh((args.coarse * 2**l,) * 3, p)
        m.clamp(ALL_FACES, 1)
        meshes.append(m)
    m = meshes[-1]
    sys_mf = MatrixFree.from_mesh(m, 1)
    tm = Timer(sys_mf)
    sys_op = LaplaceOperator(sys_mf, args.dt)
    mfs, ops, sms, trs = [], [], [], []
    launches0 = L.load().b200mf_launch_count()
    with tm("level MatrixFree objects (host, once: setup_system)"):
        for l, mm in enumerate(meshes):
            mfs.append(MatrixFree.from_mesh(mm, 1, dtype="f32"))
    with tm("per time step: diagonals, transfers (solve_gmg_preconditioner :761-788)"):
        for l, mm in enumerate(meshes):
            mf = mfs[l]
            op = LaplaceOperator(mf, args.dt)
            op.compute_diagonal()  # heat_transfer.h:784
            ops.append(op)
            if l == 0:  # :780-782
                sms.append(PreconditionChebyshev(op, degree=-1, smoothing_range=1e-3, eig_cg_n_iterations=mm.n_nodes))
            else:       # :774-776
                sms.append(PreconditionChebyshev(op, degree=5, smoothing_range=15.0, eig_cg_n_iterations=10))
                trs.append(MGTransferMatrixFree(mfs[l - 1], mf, children_of(meshes[l - 1])))
    with tm("chebyshev eigenvalue estimates"):
        info = [s.estimate_eigenvalues() for s in sms]
    mg = PreconditionMG(sms, trs, sys_mf)
    b = np.random.default_rng(0).uniform(-1, 1, m.n_nodes)
    b[m.constrained_dofs(1)] = 0.0
    rhs, x = DeviceVector(sys_mf, b), DeviceVector(sys_mf)
    tol = 1e-12 * np.linalg.norm(b)  # heat_transfer.h:719-720
    cg = SolverCG(tol, 200)
    cg.solve(sys_op, x, rhs, mg, x_is_zero=True)  # warm-up
    x.set_zero()
    with tm("cg solve"):
        res = cg.solve(sys_op, x, rhs, mg, x_is_zero=True)
    # one V-cycle and one system vmult in isolation
    y = DeviceVector(sys_mf)
    sys_op.vmult(y, rhs)  # first launch of the plain-vmult instantiation (CG uses the fused-dot one): module load, not timed
    with tm("10 V-cycles"):
        for _ in range(10):
            mg.vmult(y, rhs)
    with tm("10 system vmults"):
        for _ in range(10):
            sys_op.vmult(y, rhs)
    n = m.n_nodes
    print(json.dumps({"workload": f"heat GMG-CG, Q{p}, cube {m.ncells[0]}^3 cells, {n} DoFs, {args.levels} levels, dt={args.dt}",
                      "n_dofs": n, "cg_iterations": res["iterations"], "residual": res["residual"], "tolerance": tol,
                      "solve_s": tm.t["cg solve"], "dofs_per_s_solve": n / tm.t["cg solve"],
                      "vcycle_ms": 100 * tm.t["10 V-cycles"], "system_vmult_ms": 100 * tm.t["10 system vmults"],
                      "times_s": tm.t, "level0": info[0], "level_top": info[-1],
                      "gpu_launches": L.load().b200mf_launch_count() - launches0}), flush=True)


def run_newton(args):
    from b200mf import DeviceVector, MatrixFree, MGTransferMatrixFree, NeoHookOperator, PreconditionChebyshev, PreconditionMG, SolverCG
    from b200mf import lib as L
    from b200mf.mesh import StructuredMesh, children_of, synthetic_displacement

    p = args.degree
    meshes = []
    for l in range(args.levels):
        m = StructuredMesh((args.coarse * 2**l,) * 3, p)
        m.clamp([(0, 0)], 3)
        meshes.append(m)
    m = meshes[-1]
    u = synthetic_displacement(m.node_coordinates(), args.amplitude)
    sys_mf = MatrixFree.from_mesh(m, 3)
    tm = Timer(sys_mf)
    sys_op = NeoHookOperator()
    sys_op.set_material(sys_mf, MAT)
    disp = DeviceVector(sys_mf, u)
    sys_op.initialize(disp, args.caching)
    mfs, ops, sms, trs, disps = [], [], [], [], []
    for l, mm in enumerate(meshes):
        mf = MatrixFree.from_mesh(mm, 3, dtype="f32")
        op = NeoHookOperator()
        op.set_material(mf, MAT)
        d = DeviceVector(mf)
        op.initialize(d, args.caching)
        mfs.append(mf)
        ops.append(op)
        disps.append(d)
        if l > 0:
            trs.append(MGTransferMatrixFree(mfs[l - 1], mf, children_of(meshes[l - 1])))
    launches0 = L.load().b200mf_launch_count()

    def update_matrix_free(first):
        with tm("cache + diagonal (FP64 system)"):
            sys_op.cache()
            sys_op.compute_diagonal()
        with tm("interpolate_to_mg"):
            disps[-1].copy_from(disp)
            for l in range(len(meshes) - 1, 0, -1):
                trs[l - 1].interpolate(disps[l - 1], disps[l])
        with tm("cache + diagonal (FP32 levels)"):
            for op in ops:
                op.cache()
                op.compute_diagonal()
        with tm("chebyshev eigenvalue estimates"):
            if first:
                for l, op in enumerate(ops):
                    if l == 0:
                        sms.append(PreconditionChebyshev(op, degree=-1, smoothing_range=1e-3, eig_cg_n_iterations=min(3 * meshes[0].n_nodes, 500)))
                    else:
                        sms.append(PreconditionChebyshev(op, degree=5, smoothing_range=20.0, eig_cg_n_iterations=30))
            return [s.estimate_eigenvalues() for s in sms]

    update_matrix_free(True)  # warm-up (allocations)
    mg = PreconditionMG(sms, trs, sys_mf)
    b = np.random.default_rng(0).uniform(-1, 1, 3 * m.n_nodes)
    b[m.constrained_dofs(3)] = 0.0
    rhs, x = DeviceVector(sys_mf, b), DeviceVector(sys_mf)
    tol = 1e-6 * np.linalg.norm(b)
    cg = SolverCG(tol, 500)
    cg.solve(sys_op, x, rhs, mg, x_is_zero=True)
    tm.t.clear()
    t0 = time.perf_counter()
    info = update_matrix_free(False)
    x.set_zero()
    with tm("cg solve"):
        res = cg.solve(sys_op, x, rhs, mg, x_is_zero=True)
    total = time.perf_counter() - t0
    n = 3 * m.n_nodes
    print(json.dumps({"workload": f"neo-Hookean Newton step (update_matrix_free + GMG-CG), {args.caching}, Q{p}, box {m.ncells[0]}^3 cells, "
                                  f"{n} DoFs, {args.levels} levels", "n_dofs": n, "newton_step_s": total,
                      "cg_iterations": res["iterations"], "residual": res["residual"], "tolerance": tol, "times_s": tm.t,
                      "level0": info[0], "level_top": info[-1], "gpu_launches": L.load().b200mf_launch_count() - launches0}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["heat", "newton"])
    ap.add_argument("--degree", type=int, default=None)
    ap.add_argument("--coarse", type=int, default=None)
    ap.add_argument("--levels", type=int, default=None)
    ap.add_argument("--dt", type=float, default=0.1)
    ap.add_argument("--caching", default="tensor2")
    ap.add_argument("--amplitude", type=float, default=0.01)
    a = ap.parse_args()
    if a.what == "heat":
        a.degree, a.coarse, a.levels = a.degree or 4, a.coarse or 7, a.levels or 5
        run_heat(a)
    else:
        a.degree, a.coarse, a.levels = a.degree or 3, a.coarse or 5, a.levels or 5
        run_newton(a)
