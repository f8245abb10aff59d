"""CPU tests that pin the oracle: the reference has no
operator-level golden vectors, so the restatement is pinned by
self-consistency — dense numpy twin vs sum-factorised C, all cache levels are
the same operator (the reference's own goldens say so, tests/solid/*/*.output),
symmetry, finite differences of the residual (mf_elasticity.h:1801-1834),
scipy-assembled Laplace."""
import numpy as np
import pytest

from common import MAT, c_cells, displacement, make_mesh, np_cells, rel
from oracle import c_oracle as CO
from oracle import np_oracle as NO

LEVELS = NO.NeoHookOperator.LEVELS


def _np_op(m, level, formulation=1):
    cd = np_cells(m)
    op = NO.NeoHookOperator(cd, NO.Material(formulation=formulation, **MAT))
    op.initialize(displacement(m), level)
    op.cache()
    return op


def _c_op(m, level, formulation=1):
    op = CO.NeoHookOperator(c_cells(m), formulation=formulation, **MAT)
    op.initialize(displacement(m), level)
    op.cache()
    return op


def test_shape_1d_matches():
    for p in range(1, 7):
        S, D, w = NO.shape_1d(p, p + 1)
        S2, D2, C2, w2, nodes, qp = CO.shape_1d(p, p + 1)
        assert np.abs(S - S2).max() < 1e-13 and np.abs(D - D2).max() < 2e-12 and np.abs(w - w2).max() < 1e-14
        assert np.abs(C2 @ S2 - D2).max() < 1e-11  # collocation derivative identity
        assert abs(w.sum() - 1) < 1e-14 and np.allclose(S.sum(1), 1)


@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("level", LEVELS)
def test_c_vs_numpy_vmult(p, level):
    m = make_mesh(p)
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)
    a, b = np.zeros_like(src), np.zeros_like(src)
    _np_op(m, level).vmult(a, src)
    _c_op(m, level).vmult(b, src)
    assert rel(b, a) < 1e-13


def test_formulation0_c_vs_numpy():
    m = make_mesh(2)
    src = np.random.default_rng(0).uniform(-1, 1, 3 * m.n_nodes)
    a, b = np.zeros_like(src), np.zeros_like(src)
    _np_op(m, "tensor2", 0).vmult(a, src)
    _c_op(m, "tensor2", 0).vmult(b, src)
    assert rel(b, a) < 1e-13


def test_cache_levels_are_the_same_operator():
    m = make_mesh(2)
    src = np.random.default_rng(1).uniform(-1, 1, 3 * m.n_nodes)
    outs = []
    for level in LEVELS:
        d = np.zeros_like(src)
        _c_op(m, level).vmult(d, src)
        outs.append(d)
    for d in outs[1:]:
        assert rel(d, outs[0]) < 1e-12


def test_symmetry_and_constrained_identity():
    m = make_mesh(2)
    op = _c_op(m, "tensor2")
    rng = np.random.default_rng(2)
    u, v = rng.uniform(-1, 1, (2, 3 * m.n_nodes))
    Au, Av = np.zeros_like(u), np.zeros_like(u)
    op.vmult(Au, u)
    op.vmult(Av, v)
    assert abs(v @ Au - u @ Av) < 1e-12 * abs(v @ Au)
    c = m.constrained_dofs(3)
    assert np.array_equal(Au[c], u[c])  # mf_nh_operator.h:623-624


@pytest.mark.parametrize("formulation", [0, 1])
def test_tangent_is_derivative_of_residual(formulation):
    """A(u) d = -dR/du d  (R = -internal - inertia, mf_elasticity.h:1830-1834) with
    a = alpha_1 * u so that the inertia term linearises to rho*alpha_1*M."""
    m = make_mesh(2, clamp=())
    cd = np_cells(m)
    mat = NO.Material(formulation=formulation, **MAT)
    u = displacement(m)
    op = NO.NeoHookOperator(cd, mat)
    op.initialize(u, "tensor2")
    op.cache()
    d = np.random.default_rng(3).uniform(-1, 1, u.shape)
    Ad = np.zeros_like(u)
    op.vmult(Ad, d)
    eps = 1e-6
    R = lambda w: NO.nh_residual(cd, mat, w, MAT["alpha_1"] * w, np.zeros(3))
    fd = -(R(u + eps * d) - R(u - eps * d)) / (2 * eps)
    assert rel(fd, Ad) < 1e-7


def test_diagonal_matches_explicit():
    m = make_mesh(1, ncells=(2, 2, 2))
    op = _c_op(m, "tensor2")
    diag = op.compute_diagonal().copy()
    n = 3 * m.n_nodes
    for i in range(0, n, 7):
        e = np.zeros(n)
        e[i] = 1
        out = np.zeros(n)
        op.vmult(out, e)
        assert abs(out[i] - diag[i]) < 1e-12 * abs(diag[i])
    npop = _np_op(m, "tensor2")
    assert rel(npop.compute_diagonal(), diag) < 1e-13


def test_laplace_against_assembled_matrix():
    import scipy.sparse as sp

    p = 2
    m = make_mesh(p, ncomp=1, clamp=((0, 0), (1, 1)))
    cd = np_cells(m, 1)
    dt = 0.1
    op = NO.LaplaceOperator(cd, dt)
    # assemble M + dt K densely per cell from the same tables
    N, dN = cd.N, cd.dN
    rows, cols, vals = [], [], []
    for c in range(cd.n_cells):
        G = np.einsum("qef,fqa->qea", cd.invjac[c], dN)  # physical gradients of basis
        Ke = np.einsum("q,qa,qb->ab", cd.JxW[c], N, N) + dt * np.einsum("q,qea,qeb->ab", cd.JxW[c], G, G)
        idx = cd.dof_idx[c]
        rows.append(np.repeat(idx, len(idx)))
        cols.append(np.tile(idx, len(idx)))
        vals.append(Ke.ravel())
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(cd.n_dofs, cd.n_dofs)).tocsr().toarray()
    con = cd.constrained
    A[con, :] = 0
    A[:, con] = 0
    A[con, con] = 1
    src = np.random.default_rng(4).uniform(-1, 1, cd.n_dofs)
    d = np.zeros_like(src)
    op.vmult(d, src)
    assert rel(d, A @ src) < 1e-13
    cop = CO.LaplaceOperator(c_cells(m, 1), dt)
    d2 = np.zeros_like(src)
    cop.vmult(d2, src)
    assert rel(d2, d) < 1e-13
    assert rel(cop.compute_diagonal(), np.diag(A)) < 1e-13


def test_heat_rhs_cells_is_mass_action_plus_source():
    """oracle pin for SURVEY §8(f) row 4 (heat_transfer.h:599-615): int (u_old + dt f) phi == M u_old + dt f M 1 for constant f,
    with M taken from the (independently pinned) LaplaceOperator at dt = 0."""
    from common import make_mesh, np_cells

    m = make_mesh(2, (3, 2, 2), ncomp=1, clamp=None)
    cd = np_cells(m, 1)
    u = np.random.default_rng(0).uniform(-1, 1, m.n_nodes)
    r = NO.heat_rhs_cells(cd, 0.1, u, None, 2.0)
    mass = NO.LaplaceOperator(cd, 0.0)
    mu, m1 = np.zeros_like(u), np.zeros_like(u)
    mass.vmult(mu, u)
    mass.vmult(m1, np.ones_like(u))
    assert np.linalg.norm(r - (mu + 0.1 * 2.0 * m1)) <= 1e-13 * np.linalg.norm(r)
    # a q-point table equal to the constant gives the same vector
    r2 = NO.heat_rhs_cells(cd, 0.1, u, np.full((cd.n_cells, cd.n_q), 2.0), 0.0)
    assert np.array_equal(r, r2)



def test_interface_traction_general_geometry_matches_box_version():
    """The Q1-geometry restatement of the interface-face loop (mf_elasticity.h:1849-1897) equals the axis-aligned one
    that reproduces the reference's Turek-Hron goldens (tests/test_golden_turek.py runs it in 2-D), in 3-D on a box."""
    from b200mf.mesh import StructuredMesh, synthetic_displacement
    from oracle.solid_driver import interface_traction_rhs, interface_traction_rhs_q1

    m = StructuredMesh((2, 2, 1), 2, upper=(1.0, 0.5, 0.25))
    u = synthetic_displacement(m.node_coordinates(), 0.05).ravel()
    faces = [(1, 1), (0, 0), (2, 1)]
    t = np.array([0.3, -4.0, 1.5])
    ref = interface_traction_rhs(m, u, t, faces)
    cells, fno = [], []
    for d, side in faces:
        on = np.flatnonzero(m.cell_ijk[:, d] == (0 if side == 0 else m.ncells[d] - 1))
        cells += list(on)
        fno += [2 * d + side] * on.size
    got = interface_traction_rhs_q1(m, u, t, cells, fno)
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_simd_over_cells_vmult_equals_scalar_loop(p):
    """The CPU arm of bench.py (8 cells per batch, one SIMD lane per cell, oracle/mf_oracle.c orc_nh_vmult_simd) is the
    same operator as the scalar restatement; other cache levels report that they have no batched form."""
    from common import MAT, c_cells, displacement, make_mesh
    from oracle import c_oracle as CO

    m = make_mesh(p, (3, 3, 2) if p < 4 else (2, 2, 1))
    m.material_id[::3] = 1
    op = CO.NeoHookOperator(c_cells(m), formulation=1, mu_inclusion=2.0e6, **MAT)
    op.initialize(displacement(m), "tensor2")
    op.cache()
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)
    a, b = np.zeros_like(src), np.zeros_like(src)
    op.vmult(a, src)
    assert op.vmult_simd(b, src)
    assert np.linalg.norm(a - b) <= 1e-14 * np.linalg.norm(a)
    op4 = CO.NeoHookOperator(c_cells(m), formulation=1, **MAT)
    op4.initialize(displacement(m), "tensor4")
    op4.cache()
    assert not op4.vmult_simd(b, src)


# ---------------------------------------------------------------------------
# heat coupling faces: BoundaryMassOperator / BoundaryProjector / flux faces (oracle/boundary_oracle.py)
def _interface_faces(m, faces):
    cells, fno = [], []
    for d, side in faces:
        on = np.flatnonzero(m.cell_ijk[:, d] == (0 if side == 0 else m.ncells[d] - 1))
        cells.append(on)
        fno.append(np.full(on.size, 2 * d + side))
    return np.concatenate(cells), np.concatenate(fno)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_boundary_mass_oracle_closed_forms(p):
    """Pins of the face restatement that need no reference run: 1^T M 1 = area of the interface; M is symmetric; the row-sum
    'diagonal' is positive exactly on the interface DoFs; the projector reproduces polynomial data of degree <= p exactly on the
    interface and leaves every other entry of the vector untouched (boundary_projector.h:86-110)."""
    from b200mf.mesh import StructuredMesh
    from oracle import boundary_oracle as BO

    m = StructuredMesh((3, 2, 2), p, lower=(0.0, 0.0, 0.0), upper=(1.5, 1.0, 0.5))
    faces = ((0, 1), (2, 0))  # x = 1.5 (area 0.5) and z = 0 (area 1.5)
    cells, fno = _interface_faces(m, faces)
    op = BO.BoundaryMassOperator(m, cells, fno)
    one = np.ones(m.n_nodes)
    out = np.zeros(m.n_nodes)
    op.vmult(out, one)
    assert abs(out.sum() - 2.0) < 1e-13
    rng = np.random.default_rng(5)
    a, b = rng.uniform(-1, 1, m.n_nodes), rng.uniform(-1, 1, m.n_nodes)
    Ma, Mb = np.zeros_like(a), np.zeros_like(b)
    op.vmult(Ma, a)
    op.vmult(Mb, b)
    assert abs(b @ Ma - a @ Mb) < 1e-13
    diag = op.compute_diagonal()
    on = np.zeros(m.n_nodes, bool)
    on[op.boundary_indices] = True
    assert np.all(diag[on] > 0) and np.all(diag[~on] == 0) and np.all(op.inverse_diagonal[~on] == 0)
    x = m.node_coordinates()
    assert np.array_equal(on, np.isclose(x[:, 0], 1.5) | np.isclose(x[:, 2], 0.0))
    g = 1.0 + x[:, 0] ** p - 0.5 * x[:, 1] ** p + x[:, 1] * x[:, 2]
    residual = np.zeros(m.n_nodes)
    op.vmult(residual, g)            # nodal 'residual' whose interface part is M g
    off = rng.uniform(-1, 1, m.n_nodes)
    residual[~on] = off[~on]         # entries off the interface: zero row, zero rhs, zero preconditioner -> kept
    r = BO.project(op, residual)
    assert r.converged and r.iterations <= 500
    assert np.abs(residual[on] - g[on]).max() < 1e-9
    assert np.array_equal(residual[~on], off[~on])
    # flux faces: int_Gamma (-q dt) phi dA with constant q sums to -q dt * area
    rhs = BO.flux_rhs(op, np.full((len(cells), (p + 1) ** 2, 1), 3.0), -0.1)
    assert abs(rhs.sum() + 3.0 * 0.1 * 2.0) < 1e-13


def test_boundary_mass_oracle_curved_area():
    """on the cylinder shell the face Jacobian is not constant: the Q1 facets of the inner wall lie just inside r_in and their
    area tends to 2 pi r_in L from below (cases/elasticity/tube_3d.h:23-33: r_in 0.5, length 5, scaled 1e-2)"""
    from b200mf.mesh import tube_hierarchy
    from oracle import boundary_oracle as BO

    m = tube_hierarchy(2, 1, first_refinement=1)[-1]
    d_rad = int(np.argmin(m.ncells))  # one cell layer through the wall
    cells, fno = _interface_faces(m, ((d_rad, 0),))
    op = BO.BoundaryMassOperator(m, cells, fno)
    x = op.points.reshape(-1, 3)
    axis = int(np.argmax(x.max(0) - x.min(0)))  # the tube axis: the longest extent of the wall
    perp = [e for e in range(3) if e != axis]
    rad = np.sqrt(x[:, perp[0]] ** 2 + x[:, perp[1]] ** 2)
    assert rad.max() <= 0.5e-2 + 1e-12 and rad.min() > 0.45e-2
    exact = 2 * np.pi * 0.5e-2 * 5e-2
    assert 0.9 * exact < op.jxw.sum() < exact


def test_host_stream_triad_is_reported():
    """bench.py puts the host's STREAM-triad bandwidth next to the CPU arm's algorithmic GB/s (oracle/mf_oracle.c)"""
    assert CO.host_stream_triad_gbs(2_000_000, 2) > 0.1
