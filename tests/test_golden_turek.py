"""Pins the oracle against the reference's OWN golden outputs (tests/golden/turek_solid.json, extracted from
/root/reference/tests/solid/*/*_serial.output by tests/golden/make_turek_golden.py): the 2-D Q1 Turek-Hron flap,
320 cells / 738 DoFs, two Newmark steps of 0.1 s under the constant interface traction (0,-4) the reference's
fake coupling partner writes (tests/solid/dummy_tester.cpp:193-197) and body force (0,-0.5).

Reproduced with oracle/solid_driver.py on top of the same oracle operators and solver restatement the GPU parity
tests compare against: reference volume, Newton iterations per step (4, 4), watch-point displacement to the 6
significant digits the golden holds, and the average CG iterations per Newton step (29 for all four cache levels with
GMG, 24 for formulation 0, 701 with Jacobi) within +-1 (BASELINE.json north_star)."""
import json
import os

import numpy as np
import pytest

from b200mf.mesh import StructuredMesh
from oracle.solid_driver import SolidOracle

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "turek_solid.json")))


def turek_meshes(p, refinement):
    out = []
    for l in range(refinement + 1):  # include/cases/elasticity/turek_hron.h:23-41: 5 x 1 cells on [0.24899,0.6]x[0.19,0.21]
        m = StructuredMesh((5 * 2**l, 2**l), p, lower=(0.24899, 0.19), upper=(0.6, 0.21))
        m.clamp([(0, 0)], 2)  # boundary id 0 (x min) -> clamped_mesh_id, turek_hron.h:81-82
        out.append(m)
    return out


def sig(x, n=6):
    return float(f"{x:.{n - 1}e}")


@pytest.mark.parametrize("case", sorted(GOLD))
def test_reference_golden_output(case):
    g = GOLD[case]
    prm = g["prm"]
    meshes = turek_meshes(prm["degree"], prm["refinement"])
    m = meshes[-1]
    assert m.n_cells == g["n_cells"] and 2 * m.n_nodes == g["n_dofs"]
    s = SolidOracle(meshes, prm["mu"], prm["nu"], prm["rho"], prm["delta_t"], formulation=prm["formulation"],
                    caching=prm["caching"], body_force=(0.0, -0.5), traction=(0.0, -4.0),
                    traction_faces=((0, 1), (1, 0), (1, 1)),  # ids 1, 2, 3 -> interface_id, turek_hron.h:75-79
                    preconditioner=prm["preconditioner"], relaxation=prm["relaxation"], tol_lin=prm["tol_lin"],
                    n_cheb_cg=prm["n_cheb_cg"], max_cg_multiplier=2)
    assert abs(float(np.sum(s.cells[-1].JxW)) - g["reference_volume"]) < 5e-8
    xy = m.node_coordinates()
    node = int(np.argmin(np.abs(xy[:, 0] - 0.6) + np.abs(xy[:, 1] - 0.2)))  # Output points = 0.6, 0.2
    n_steps = int(round(prm["end_time"] / prm["delta_t"]))
    for step in range(n_steps):
        its = s.solve_nonlinear_timestep()
        s.advance_time_window()
        assert its == g["newton_iterations"][step]
        d = s.displacement_at_node(node)
        for c in range(2):
            assert sig(d[c]) == pytest.approx(g["displacement"][step][c], rel=2e-6), (step, c, d)
    avg = s.cg_iterations // s.cg_solves  # integer division as in mf_elasticity.h:619-627
    tol = 1 if prm["preconditioner"] == "gmg" else max(1, g["average_cg_iterations"] // 100)
    assert abs(avg - g["average_cg_iterations"]) <= tol, (avg, g["average_cg_iterations"])
