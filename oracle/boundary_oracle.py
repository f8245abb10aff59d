"""CPU restatement of the heat solver's interface-face operators.  TEST INFRASTRUCTURE ONLY (imported by tests/ only).

Follows
  include/adapter/boundary_mass_operator.h:137-161   BoundaryMassOperator::apply_add: dst += int_Gamma u phi dA over the
                                                     boundary faces with id == interface_id (FEFaceIntegrator, values only)
  include/adapter/boundary_mass_operator.h:79-118    compute_diagonal: diagonal := apply_add(1) (row sums), constrained entries 1,
                                                     inverse only where the entry is > 0
  include/adapter/boundary_projector.h:86-110        project: rhs = residual on the owned interface DoFs, SolverCG under
                                                     ReductionControl(500, 0, 1e-12), PreconditionJacobi(0.8), x0 = residual
  include/heat_transfer/heat_transfer.h:617-652      assemble_rhs, interface faces: rhs += int_Gamma (-flux(q) dt) phi dA
  include/heat_transfer/heat_transfer.h:899-940      evaluate_boundary_flux: residual = A_inhom u - M(u_old + dt f), / dt, project
parity unpinned at the operator level: the reference holds no golden vector for these face operators (its heat tests need
the coupling partner); pinned here by closed-form identities (tests/test_oracle.py: area of the face set, exact projection
of polynomial data) and used as the checker of the CUDA path (tests/test_gpu_heat.py).

Geometry: Q1 cells from their 8 vertices (vertex m at (m&1, m>>1&1, m>>2&1)); face f = (cells[f], face_no[f] = 2*direction +
side); face quadrature points ordered q = q_t1 + nq*q_t2 with the tangential directions t1 < t2; nq = p + 1 Gauss points.
Dense per-face loops - sized for the test meshes."""
from __future__ import annotations

import numpy as np

from . import np_oracle as NO
from . import solver_oracle as SO


def face_tables(mesh, cells, face_no):
    """val [n_faces, nq^2, n^3]: phi_a at the face quadrature points (zero unless node a lies on the face);
    jxw [n_faces, nq^2]: |det J J^-T N_ref| w_q1 w_q2;  xq [n_faces, nq^2, 3]: positions."""
    dim, p, n = mesh.dim, mesh.p, mesh.p + 1
    assert dim == 3
    xg, wg = NO.gauss(n)
    nodes = NO.gll_nodes(p)
    Vq, _ = NO.lagrange_eval(nodes, xg)
    aidx = mesh.loc_ijk
    nf = len(cells)
    val = np.zeros((nf, n * n, n**3))
    jxw = np.zeros((nf, n * n))
    pos = np.zeros((nf, n * n, 3))
    for f, (c, fn) in enumerate(zip(cells, face_no)):
        d, side = int(fn) // 2, int(fn) % 2
        t1, t2 = [e for e in range(3) if e != d]
        Vs, _ = NO.lagrange_eval(nodes, np.array([float(side)]))
        verts = mesh.vertices[c]
        for q2 in range(n):
            for q1 in range(n):
                q = q1 + n * q2
                tabs = {d: Vs[0], t1: Vq[q1], t2: Vq[q2]}
                v = np.ones(n**3)
                for e in range(3):
                    v *= tabs[e][aidx[:, e]]
                val[f, q] = v
                xi = np.zeros(3)
                xi[d], xi[t1], xi[t2] = side, xg[q1], xg[q2]
                J = np.zeros((3, 3))
                for m in range(8):
                    bits = [(m >> e) & 1 for e in range(3)]
                    l = [xi[e] if bits[e] else 1 - xi[e] for e in range(3)]
                    dl = [1.0 if bits[e] else -1.0 for e in range(3)]
                    g = np.array([dl[0] * l[1] * l[2], l[0] * dl[1] * l[2], l[0] * l[1] * dl[2]])
                    J += np.outer(verts[m], g)
                    pos[f, q] += l[0] * l[1] * l[2] * verts[m]
                nref = np.zeros(3)
                nref[d] = 1.0
                jxw[f, q] = np.linalg.norm(np.linalg.det(J) * np.linalg.inv(J).T @ nref) * wg[q1] * wg[q2]
    return val, jxw, pos


class BoundaryMassOperator:
    def __init__(self, mesh, cells, face_no, ncomp=1, constrained=None):
        self.mesh, self.ncomp = mesh, ncomp
        self.cells, self.face_no = np.asarray(cells, dtype=np.int64), np.asarray(face_no, dtype=np.int64)
        self.val, self.jxw, self.points = face_tables(mesh, self.cells, self.face_no)
        self.n_dofs = ncomp * mesh.n_nodes
        self.constrained = np.zeros(self.n_dofs, dtype=bool)
        if constrained is not None and len(constrained):
            self.constrained[np.asarray(constrained)] = True
        # DoFTools::extract_boundary_dofs(interface id): the DoFs whose node lies on a face of the set
        on = np.zeros(mesh.n_nodes, dtype=bool)
        for c, fn in zip(self.cells, self.face_no):  # local nodes of the cell with index 0 / p in the face's normal direction
            on[mesh.cell_nodes[c][mesh.loc_ijk[:, fn // 2] == (mesh.p if fn % 2 else 0)]] = True
        self.boundary_indices = np.flatnonzero(np.repeat(on, ncomp))

    def _dofs(self, c):
        return (self.ncomp * self.mesh.cell_nodes[c][:, None] + np.arange(self.ncomp)[None, :])  # [n^3, ncomp]

    def apply_add(self, dst, src):  # :137-161
        for f, c in enumerate(self.cells):
            dofs = self._dofs(c)
            loc = np.where(self.constrained[dofs], 0.0, src[dofs])     # gather (constrained entries read as 0)
            uq = self.val[f] @ loc                                       # [nq^2, ncomp]  phi.get_value(q)
            out = self.val[f].T @ (uq * self.jxw[f][:, None])            # integrate(values)
            keep = ~self.constrained[dofs]
            np.add.at(dst, dofs[keep], out[keep])

    def vmult(self, dst, src):  # MatrixFreeOperators::Base::vmult
        dst[:] = 0.0
        self.apply_add(dst, src)
        dst[self.constrained] += src[self.constrained]

    def compute_diagonal(self):  # :79-118
        ones = np.ones(self.n_dofs)
        diag = np.zeros(self.n_dofs)
        self.apply_add(diag, ones)
        diag[self.constrained] = 1.0
        self.diagonal = diag
        self.inverse_diagonal = np.where(diag > 0, 1.0 / np.where(diag > 0, diag, 1.0), diag)
        return diag


def flux_rhs(op: BoundaryMassOperator, flux_q, scale):
    """rhs = scale * int_Gamma flux(q) phi dA (heat_transfer.h:636-647 with scale = -dt); flux_q [n_faces, nq^2, ncomp]"""
    rhs = np.zeros(op.n_dofs)
    flux_q = np.asarray(flux_q, dtype=float).reshape(len(op.cells), -1, op.ncomp)
    for f, c in enumerate(op.cells):
        dofs = op._dofs(c)
        out = op.val[f].T @ (scale * flux_q[f] * op.jxw[f][:, None])
        keep = ~op.constrained[dofs]
        np.add.at(rhs, dofs[keep], out[keep])
    return rhs


def project(op: BoundaryMassOperator, residual):
    """BoundaryProjector::project (boundary_projector.h:86-110); residual is overwritten; returns the CG result."""
    op.compute_diagonal()
    rhs = np.zeros_like(residual)
    rhs[op.boundary_indices] = residual[op.boundary_indices]
    r0 = np.empty_like(rhs)
    op.vmult(r0, residual)
    tol = 1e-12 * float(np.linalg.norm(rhs - r0))  # ReductionControl(500, 0., 1e-12): relative to the start residual

    def jacobi(dst, src):
        dst[:] = 0.8 * op.inverse_diagonal * src

    return SO.solver_cg(op.vmult, residual, rhs, jacobi, tol, 500)
