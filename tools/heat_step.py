#!/usr/bin/env python
"""One backward-Euler heat step (BASELINE configs[1]) for launch-list captures: python tools/heat_step.py [--degree 4 --coarse 7 --levels 5]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "matrix-free-dealii-precice_b200")):
    sys.path.insert(0, p)
import numpy as np
from b200mf.heat import HeatProblem
from b200mf.mesh import StructuredMesh

ap = argparse.ArgumentParser()
ap.add_argument("--degree", type=int, default=4)
ap.add_argument("--coarse", type=int, default=7)
ap.add_argument("--levels", type=int, default=5)
a = ap.parse_args()
faces = [(d, s) for d in range(3) for s in range(2)]
meshes = []
for l in range(a.levels):
    m = StructuredMesh((a.coarse * 2**l,) * 3, a.degree)
    m.clamp(faces, 1)
    meshes.append(m)
hp = HeatProblem(meshes, 0.1, source_const=1.0)
m = meshes[-1]
u0 = np.prod(np.sin(np.pi * m.node_coordinates()), axis=1)
u0[m.constrained_dofs(1)] = 0.0
hp.solution.upload(u0)
hp.time_step()
hp.mf.synchronize()
t0 = time.perf_counter()
res = hp.time_step()
hp.mf.synchronize()
t1 = time.perf_counter()
y = hp.mf.initialize_dof_vector()
hp.mg.vmult(y, hp.system_rhs)
hp.mf.synchronize()
t2 = time.perf_counter()
for _ in range(5):
    hp.mg.vmult(y, hp.system_rhs)
hp.mf.synchronize()
t3 = time.perf_counter()
print(json.dumps({"dofs": m.n_nodes, "time_step_s": t1 - t0, "cg_iterations": res["iterations"], "vcycle_ms": (t3 - t2) / 5 * 1e3, "eig": hp.eig}))
