"""GPU parity of whole backward-Euler steps of the heat solver (BASELINE.json configs[1]): b200mf.heat.
HeatProblem (assemble_rhs cell loop, level diagonals, Chebyshev estimates, GMG-CG - all through the C ABI; the scalar
operator on affine cells runs the stacked warp kernel for p <= 4) against the CPU restatement of the same steps
(oracle/np_oracle.heat_rhs_cells + oracle/solver_oracle.py, heat_transfer.h:577-812)."""
import numpy as np
import pytest

from b200mf.heat import HeatProblem
from common import np_cells, rel
from oracle import np_oracle as NO
from oracle import solver_oracle as SO
from test_solver_oracle import build_gmg, hierarchy, laplace_level

pytestmark = pytest.mark.gpu
ALL_FACES = ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1))


@pytest.mark.parametrize("p,levels,mapped", [(1, 3, False), (2, 2, False), (4, 2, False), (3, 2, True), (5, 1, False)])
def test_heat_time_steps(p, levels, mapped):
    meshes = hierarchy(p, (2, 2, 2), levels, 1, mapped=mapped, clamp=ALL_FACES)
    m = meshes[-1]
    dt, f = 0.1, 2.5
    hp = HeatProblem(meshes, dt, source_const=f)
    assert hp.mf.all_cells_affine() == (not mapped)
    x = m.node_coordinates()
    u0 = np.prod(np.sin(np.pi * x), axis=1)
    con = m.constrained_dofs(1)
    u0[con] = 0.0
    hp.solution.upload(u0)
    ops_ref, As, mg_ref = build_gmg(meshes, 1, lambda mm: laplace_level(mm, dt))
    cells = np_cells(m, 1)
    u_ref = u0.copy()
    for step in range(2):
        res = hp.time_step()
        rhs = NO.heat_rhs_cells(cells, dt, u_ref, None, f)  # (u_old + dt f, phi), constrained rows skipped (:599-615)
        rhs[con] = 0.0
        assert rel(hp.system_rhs.download(), rhs) < 1e-12
        xr = np.zeros_like(rhs)
        r = SO.solver_cg(As[-1], xr, rhs, mg_ref.vmult if levels > 1 else mg_ref.vmult, 1e-12 * np.linalg.norm(rhs), rhs.size)
        xr[con] = 0.0
        u_ref = xr
        assert res["converged"]
        assert abs(res["iterations"] - r.iterations) <= 1, (res["iterations"], r.iterations)
        assert rel(hp.solution.download(), u_ref) < 1e-8


# ---------------------------------------------------------------------------
# coupling faces of the heat solver: BoundaryMassOperator / BoundaryProjector / flux faces / evaluate_boundary_flux
def _interface_faces(m, faces):
    cells, fno = [], []
    for d, side in faces:
        on = np.flatnonzero(m.cell_ijk[:, d] == (0 if side == 0 else m.ncells[d] - 1))
        cells.append(on)
        fno.append(np.full(on.size, 2 * d + side))
    return np.concatenate(cells), np.concatenate(fno)


@pytest.mark.parametrize("p,mapped,dtype", [(1, False, "f64"), (2, True, "f64"), (3, False, "f64"), (4, True, "f64"), (2, False, "f32")])
def test_boundary_mass_operator(p, mapped, dtype):
    """include/adapter/boundary_mass_operator.h through the C ABI: vmult, vmult_add, the row-sum 'diagonal' and its guarded
    inverse, restrict (rhs[i] = residual[i] on the interface DoFs) and the flux faces of assemble_rhs, against
    oracle/boundary_oracle.py on affine and mapped cells, with and without constrained DoFs touching the interface."""
    from b200mf import BoundaryMassOperator, DeviceVector, InterfaceTraction, MatrixFree
    from b200mf.mesh import StructuredMesh, smooth_mapping
    from oracle import boundary_oracle as BO

    m = StructuredMesh((3, 2, 2), p, mapping=smooth_mapping(0.05) if mapped else None)
    m.clamp([(1, 0)], 1)  # Dirichlet on y = 0: shares an edge with both interface faces
    faces = ((0, 1), (2, 0))
    cells, fno = _interface_faces(m, faces)
    ref = BO.BoundaryMassOperator(m, cells, fno, 1, m.constrained_dofs(1))
    mf = MatrixFree.from_mesh(m, 1, dtype=dtype)
    fs = InterfaceTraction(mf, m, faces)
    op = BoundaryMassOperator(fs)
    tol = 1e-13 if dtype == "f64" else 2e-6
    rng = np.random.default_rng(p)
    u = rng.uniform(-1, 1, m.n_nodes)
    out_ref = np.zeros(m.n_nodes)
    ref.vmult(out_ref, u)
    s, d = DeviceVector(mf, u), DeviceVector(mf)
    op.vmult(d, s)
    assert rel(d.download(), out_ref) < tol
    op.vmult_add(d, s)
    assert rel(d.download(), 2 * out_ref) < tol
    op.compute_diagonal()
    ref.compute_diagonal()
    assert rel(op.diagonal(), ref.diagonal) < tol
    assert rel(op.inverse_diagonal(), ref.inverse_diagonal) < 10 * tol
    assert np.count_nonzero(op.inverse_diagonal()) == np.count_nonzero(ref.inverse_diagonal)
    op.restrict(d, s)
    want = np.zeros(m.n_nodes)
    want[ref.boundary_indices] = u[ref.boundary_indices]
    assert np.array_equal(d.download(), want if dtype == "f64" else want.astype(np.float32).astype(np.float64))
    assert np.abs(fs.quadrature_points() - ref.points.reshape(-1, 3)).max() < 1e-13
    flux = rng.uniform(-1, 1, (len(cells), (p + 1) ** 2, 1))
    fs.set_traction(flux)
    d.set_zero()
    fs.assemble_flux(d, -0.1)
    assert rel(d.download(), BO.flux_rhs(ref, flux, -0.1)) < tol


@pytest.mark.parametrize("p,mapped", [(1, False), (2, True), (4, False)])
def test_boundary_projector(p, mapped):
    """BoundaryProjector::project (include/adapter/boundary_projector.h:86-110): same CG (iterations +-1), same projected
    values on the interface, entries off the interface untouched; exact for polynomial data of degree <= p on affine cells."""
    from b200mf import BoundaryMassOperator, DeviceVector, InterfaceTraction, MatrixFree
    from b200mf.mesh import StructuredMesh, smooth_mapping
    from oracle import boundary_oracle as BO

    m = StructuredMesh((3, 2, 2), p, mapping=smooth_mapping(0.05) if mapped else None)
    faces = ((0, 1), (2, 0))
    cells, fno = _interface_faces(m, faces)
    ref = BO.BoundaryMassOperator(m, cells, fno)
    mf = MatrixFree.from_mesh(m, 1)
    op = BoundaryMassOperator(InterfaceTraction(mf, m, faces))
    x = m.node_coordinates()
    g = 1.0 + x[:, 0] ** p - 0.5 * x[:, 1] ** p + x[:, 1] * x[:, 2]
    residual = np.zeros(m.n_nodes)
    ref.vmult(residual, g)
    on = np.zeros(m.n_nodes, bool)
    on[ref.boundary_indices] = True
    residual[~on] = np.random.default_rng(1).uniform(-1, 1, m.n_nodes)[~on]
    dv = DeviceVector(mf, residual)
    info = op.project(dv)
    want = residual.copy()
    r = BO.project(ref, want)
    got = dv.download()
    assert abs(info["iterations"] - r.iterations) <= 1, (info, r.iterations)
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()
    assert np.array_equal(got[~on], residual[~on])
    if not mapped:
        assert np.abs(got[on] - g[on]).max() < 1e-9


@pytest.mark.parametrize("p", [2, 3])
def test_evaluate_boundary_flux(p):
    """LaplaceProblem::evaluate_boundary_flux (heat_transfer.h:899-940) on the device: residual = (A_inhom u - M(u_old + dt f))/dt
    projected onto the interface, after one backward-Euler step; checked against the same composition of the CPU oracles."""
    from oracle import boundary_oracle as BO

    faces_dirichlet = ((0, 0), (1, 0), (1, 1), (2, 1))
    interface = ((0, 1), (2, 0))
    meshes = hierarchy(p, (2, 2, 2), 2, 1, clamp=faces_dirichlet)
    m = meshes[-1]
    dt, f = 0.1, 1.5
    hp = HeatProblem(meshes, dt, source_const=f)
    hp.setup_interface(interface)
    x = m.node_coordinates()
    u0 = np.prod(np.sin(0.5 * np.pi * (x + 0.3)), axis=1)
    u0[m.constrained_dofs(1)] = 0.0
    hp.solution.upload(u0)
    hp.time_step()
    flux, info = hp.evaluate_boundary_flux()
    u, uo = hp.solution.download(), hp.solution_old.download()
    import copy
    free = copy.copy(m)
    free.constrained_nodes = None
    cells_free = np_cells(free, 1)
    A = NO.LaplaceOperator(cells_free, dt)
    res = np.zeros(m.n_nodes)
    A.vmult(res, u)
    res -= NO.heat_rhs_cells(cells_free, dt, uo, None, f)
    res /= dt
    fc, fn = _interface_faces(m, interface)
    r = BO.project(BO.BoundaryMassOperator(m, fc, fn), res)
    assert abs(info["iterations"] - r.iterations) <= 1
    assert np.abs(flux.download() - res).max() < 1e-8 * np.abs(res).max()
