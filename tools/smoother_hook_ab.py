#!/usr/bin/env python
"""A/B of one Chebyshev smoother application on the finest FP32 level (neo-Hookean Q3 box): event time of the whole application
and of its cell-loop launches alone.  Run twice, with B200MF_SMOOTHER_HOOK=1 / 0."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "matrix-free-dealii-precice_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from b200mf import DeviceVector, MatrixFree, NeoHookOperator, PreconditionChebyshev
from b200mf.mesh import StructuredMesh, synthetic_displacement

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = StructuredMesh((nc, nc, nc), 3)
m.clamp([(0, 0)], 3)
mf = MatrixFree.from_mesh(m, 3, dtype="f32")
op = NeoHookOperator()
op.set_material(mf, dict(mu=0.5e6, nu=0.4, rho=1000.0, alpha_1=4e4, formulation=1))
d = DeviceVector(mf, synthetic_displacement(m.node_coordinates(), 0.02))
op.initialize(d, "tensor2")
op.cache()
op.compute_diagonal()
sm = PreconditionChebyshev(op, degree=5, smoothing_range=20.0, eig_cg_n_iterations=10, n_global_dofs=3 * m.n_nodes)
sm.estimate_eigenvalues()
src = DeviceVector(mf, np.random.default_rng(0).uniform(-1, 1, 3 * m.n_nodes))
dst = DeviceVector(mf)
for _ in range(3):
    sm.step(dst, src)
mf.synchronize()
mf.time_kernels(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    sm.step(dst, src)
e1.record()
torch.cuda.synchronize()
ms, n = mf.kernel_time()
print(json.dumps({"hook": os.environ.get("B200MF_SMOOTHER_HOOK", "1"), "dofs": 3 * m.n_nodes, "ms_per_step_application": e0.elapsed_time(e1) / reps,
                  "cell_kernel_ms_each": ms / n, "cell_launches_per_application": n / reps,
                  "other_ms_per_application": e0.elapsed_time(e1) / reps - ms / reps, "checksum": float(np.abs(dst.download()).sum())}))
