"""GPU parity of the interface-face part of the residual (mf_elasticity.h:1849-1897) through the
C ABI (b200mf_face_set_*) against oracle/solid_driver.interface_traction_rhs_q1."""
import numpy as np
import pytest

from b200mf import DeviceVector, InterfaceTraction, MatrixFree
from common import displacement, make_mesh, rel
from oracle import np_oracle as NO
from oracle.solid_driver import interface_traction_rhs_q1

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("mapped", [True, False], ids=["general", "affine"])
def test_interface_traction_faces(p, mapped):
    m = make_mesh(p, (3, 2, 2), mapped=mapped, permute=7)
    mf = MatrixFree.from_mesh(m, 3)
    u = displacement(m, 0.08).ravel()
    u[m.constrained_dofs(3)] = 0.0
    faces = [(1, 1), (0, 1), (2, 0)]
    ft = InterfaceTraction(mf, m, faces)
    n = p + 1
    assert ft.n_q_points() == ft.cells.size * n * n
    rng = np.random.default_rng(p)
    tq = rng.uniform(-4, 4, (ft.cells.size, n * n, 3))
    ref = interface_traction_rhs_q1(m, u, tq, ft.cells, ft.face_no)
    ref[m.constrained_dofs(3)] = 0.0  # distribute_local_to_global does not write constrained rows
    rhs, du = DeviceVector(mf), DeviceVector(mf, u)
    ft.set_traction(tq)
    ft.assemble(rhs, du)
    assert rel(rhs.download(), ref) < 1e-12
    # constant traction entry point + accumulation into an existing rhs
    t = np.array([0.0, -4.0, 0.5])
    ref2 = interface_traction_rhs_q1(m, u, t, ft.cells, ft.face_no)
    ref2[m.constrained_dofs(3)] = 0.0
    ft.assemble(rhs, du, t)
    assert rel(rhs.download(), ref + ref2) < 1e-12


def test_face_quadrature_points():
    p = 2
    m = make_mesh(p, (2, 2, 1), mapped=True)
    mf = MatrixFree.from_mesh(m, 3)
    ft = InterfaceTraction(mf, m, [(2, 1)])
    pts = ft.quadrature_points().reshape(ft.cells.size, (p + 1) ** 2, 3)
    xq, _ = NO.gauss(p + 1)
    for f, c in enumerate(ft.cells):
        v = m.vertices[c]
        for q2 in range(p + 1):
            for q1 in range(p + 1):
                xi = (xq[q1], xq[q2], 1.0)
                x = sum(np.prod([xi[e] if (k >> e) & 1 else 1 - xi[e] for e in range(3)]) * v[k] for k in range(8))
                assert np.abs(pts[f, q1 + (p + 1) * q2] - x).max() < 1e-14
