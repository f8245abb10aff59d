"""GPU parity of whole Newmark/Newton steps (BASELINE.json north_star: "full GMG-preconditioned Newton steps matching the
reference's results").

b200mf.solid.Solid drives residual -> cache -> diagonals -> interpolate_to_mg -> Chebyshev eigenvalue estimates ->
GMG-CG -> update entirely through the C ABI on device vectors; oracle/solid_driver.SolidOracle is the CPU restatement of
the same lines of mf_elasticity.h (the one that reproduces the reference's own 2-D goldens, tests/test_golden_turek.py),
run here in 3-D on a clamped bar.  Bars of the north star: Newton iterations equal, CG iterations within +-1 per Newton
step, displacement to 1e-8 of its maximum."""
import numpy as np
import pytest

from b200mf.mesh import StructuredMesh
from b200mf.solid import Solid
from oracle.solid_driver import SolidOracle

pytestmark = pytest.mark.gpu

MATERIAL = dict(mu=0.5e6, nu=0.4, rho=1000.0)  # elasticity.prm / turek_hron.h:46-47


def bar_hierarchy(p, levels, coarse=(2, 1, 1), mapped=False):
    from b200mf.mesh import smooth_mapping

    meshes = []
    for l in range(levels):
        m = StructuredMesh(tuple(c * 2**l for c in coarse), p, upper=(2.0, 1.0, 1.0),
                           mapping=smooth_mapping(0.03) if mapped else None)
        m.clamp([(0, 0)], 3)
        meshes.append(m)
    return meshes


def run_both(p, levels, n_steps=2, mapped=False, **kw):
    meshes = bar_hierarchy(p, levels, mapped=mapped)
    args = dict(MATERIAL, delta_t=0.01, body_force=(0.0, -800.0, 200.0), n_cheb_cg=30)
    args.update(kw)
    ref = SolidOracle(meshes, **args)
    dev = Solid(meshes, **args)
    for _ in range(n_steps):
        ref.solve_nonlinear_timestep()
        ref.advance_time_window()
        dev.solve_nonlinear_timestep()
        dev.advance_time_window()
    return ref, dev


@pytest.mark.parametrize("p,levels,caching,formulation", [(2, 2, "tensor2", 1), (1, 3, "tensor2", 1), (3, 2, "tensor2", 1),
                                                           (2, 2, "scalar_referential", 1), (2, 2, "tensor4", 1),
                                                           (2, 2, "tensor2", 0)])
def test_newmark_newton_gmg_steps(p, levels, caching, formulation):
    ref, dev = run_both(p, levels, caching=caching, formulation=formulation)
    assert dev.newton_iterations == ref.newton_iterations  # Newton iterations per time step: equal
    assert dev.cg_solves == ref.cg_solves
    # CG iterations: +-1 in every linear solve
    assert max(abs(a - b) for a, b in zip(dev.cg_history, ref.cg_history)) <= 1, (dev.cg_history, ref.cg_history)
    u_ref, u = ref.total, dev.total_displacement()
    scale = np.abs(u_ref).max()
    assert scale > 0.05  # a genuinely nonlinear step (10 % of the bar's thickness)
    assert np.abs(u - u_ref).max() <= 1e-8 * scale
    # Newmark state after the last window
    assert np.abs(dev.velocity_old.download() - ref.velocity_old).max() <= 1e-6 * np.abs(ref.velocity_old).max()
    assert np.abs(dev.acceleration_old.download() - ref.acceleration_old).max() <= 1e-5 * np.abs(ref.acceleration_old).max()


def test_newmark_newton_jacobi_mapped_cells():
    """Jacobi-preconditioned CG (mf_elasticity.h:2028-2030, relaxation 0.65) on non-affine cells."""
    ref, dev = run_both(2, 1, mapped=True, preconditioner="jacobi", max_cg_multiplier=2)
    assert dev.newton_iterations == ref.newton_iterations
    assert max(abs(a - b) for a, b in zip(dev.cg_history, ref.cg_history)) <= 1, (dev.cg_history, ref.cg_history)
    scale = np.abs(ref.total).max()
    assert np.abs(dev.total_displacement() - ref.total).max() <= 1e-8 * scale


def test_newmark_newton_with_interface_traction():
    """The coupled configuration of the reference's tests: a constant traction on the interface boundary (the fake
    partner of tests/solid/dummy_tester.cpp:193-197) pulled back in every Newton iteration (mf_elasticity.h:1849-1897)."""
    ref, dev = run_both(2, 2, traction=(0.0, -2.0e4, 5.0e3), traction_faces=((1, 1), (0, 1)))
    assert dev.newton_iterations == ref.newton_iterations
    assert max(abs(a - b) for a, b in zip(dev.cg_history, ref.cg_history)) <= 1, (dev.cg_history, ref.cg_history)
    scale = np.abs(ref.total).max()
    assert scale > 0.05
    assert np.abs(dev.total_displacement() - ref.total).max() <= 1e-8 * scale


def test_newmark_newton_cg_coarse_solver():
    """'MF Chebyshev coarse = false': SolverCG coarse grid solver inside the V-cycle (mf_elasticity.h:1183-1198, 1217-1224)."""
    ref, dev = run_both(2, 2, coarse_solver="cg")
    assert dev.newton_iterations == ref.newton_iterations
    assert max(abs(a - b) for a, b in zip(dev.cg_history, ref.cg_history)) <= 1, (dev.cg_history, ref.cg_history)
    scale = np.abs(ref.total).max()
    assert np.abs(dev.total_displacement() - ref.total).max() <= 1e-8 * scale


def test_tube_newton_steps_against_golden():
    """BASELINE configs[3] in miniature: the tube of include/cases/elasticity/tube_3d.h (cylinder shell: curved, non-affine
    Q1 cells, periodic around the axis, ends clamped, pressure-like traction on the inner wall = the coupling interface),
    two Newmark steps with the GMG-preconditioned Newton solver on the device against the fixture the CPU restatement
    produced (tests/golden/make_tube_golden.py; the live oracle run takes ~90 s)."""
    import os

    from b200mf.mesh import tube_hierarchy, tube_pressure_traction

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tube_newton_q2_r1.npz"))
    meshes = tube_hierarchy(int(g["degree"]), int(g["levels"]))
    dev = Solid(meshes, mu=115384.0, nu=0.3, rho=1200.0, delta_t=1e-4, n_cheb_cg=30,
                traction=tube_pressure_traction(float(g["pressure"])), traction_faces=((2, 0),))
    assert not dev.mf.all_cells_affine()
    dev.run(2)
    assert dev.newton_iterations == list(g["newton_iterations"])
    assert max(abs(a - b) for a, b in zip(dev.cg_history, list(g["cg_history"]))) <= 1, (dev.cg_history, list(g["cg_history"]))
    u_ref = g["displacement"]
    assert np.abs(dev.total_displacement() - u_ref).max() <= 1e-8 * np.abs(u_ref).max()
