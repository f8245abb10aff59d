#!/usr/bin/env python
"""Generates tests/golden/tube_newton_q2_r1.npz: two Newmark steps of the solid solver on the tube geometry of
include/cases/elasticity/tube_3d.h:23-33, 74 (cylinder shell, curved = non-affine Q1 cells, both ends clamped, pressure-like
traction on the inner wall = the coupling interface), computed by the CPU restatement oracle/solid_driver.SolidOracle
(GMG-preconditioned CG, FP32 levels).  The live oracle run takes ~90 s, too long for the GPU test suite, so
tests/test_gpu_newton.py::test_tube_newton_steps_against_golden compares the device run with this fixture.

    python tests/golden/make_tube_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "matrix-free-dealii-precice_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

from b200mf.mesh import tube_hierarchy, tube_pressure_traction  # noqa: E402
from oracle.solid_driver import SolidOracle  # noqa: E402

P, LEVELS, PRESSURE = 2, 2, 5.0e4
PARAMS = dict(mu=115384.0, nu=0.3, rho=1200.0, delta_t=1e-4, n_cheb_cg=30)  # examples/tube3d/solid/tube3d.prm

if __name__ == "__main__":
    meshes = tube_hierarchy(P, LEVELS)
    so = SolidOracle(meshes, traction=tube_pressure_traction(PRESSURE), traction_faces=((2, 0),), **PARAMS)
    so.run(2)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tube_newton_q2_r1.npz")
    np.savez_compressed(out, displacement=so.total, newton_iterations=np.array(so.newton_iterations),
                        cg_history=np.array(so.cg_history), degree=P, levels=LEVELS, pressure=PRESSURE)
    print(out, so.newton_iterations, so.cg_history, np.abs(so.total).max())
