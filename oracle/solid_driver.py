"""CPU restatement of the reference's solid time/Newton loop for ONE purpose: reproducing the reference's own golden
outputs (tests/solid/turek_*/*.output) with the oracle operators, which pins the oracle's quadrature-point arithmetic
(residual + tangent + caches) and the solver restatement (GMG/Chebyshev/CG) against numbers produced by the real
reference.  TEST INFRASTRUCTURE ONLY.

Follows include/solid_mechanics/mf_elasticity.h:
  run                       :553-628     solve_nonlinear_timestep :1400-1491   check_convergence :1357-1395
  update_acceleration       :1612-1620   update_velocity          :1600-1608   Newmark constants :237-245
  assemble_residual (fast)  :1764-1846   interface faces          :1849-1897   constraints.set_zero + norm :1899-1907
  solve_linear_system       :1981-2070   setup_gmg                :1140-1279   setup_mg_interpolation :881-929
The coupling partner of the reference's tests (tests/solid/dummy_tester.cpp:193-197, 224) writes the constant stress
(0, -4) on the interface mesh every window and the consistent RBF mapping reproduces constants, so the preCICE read in
:1876-1879 is replaced by that constant.  Geometry/BCs: include/cases/elasticity/turek_hron.h:19-92.
"""
from __future__ import annotations

import numpy as np

from . import np_oracle as NO
from . import solver_oracle as SO


def lagrange_table(p, x):
    V, D = NO.lagrange_eval(NO.gll_nodes(p), np.atleast_1d(np.asarray(x, dtype=float)))
    return V, D


def interface_traction_rhs(mesh, total_displacement, traction, faces):
    """FEFaceEvaluation loop of assemble_residual (:1849-1897) on axis-aligned box cells:
    rhs_a += int_face phi_a * t * |det F F^-T N| dA over the listed (direction, side) faces of the box."""
    dim, p, n = mesh.dim, mesh.p, mesh.p + 1
    xq, wq = NO.gauss(n)  # Quadrature order = p+1 on faces as in the cells
    Vq, Dq = lagrange_table(p, xq)
    rhs = np.zeros(mesh.n_nodes * dim)
    h = (mesh.upper - mesh.lower) / np.asarray(mesh.ncells)
    u = total_displacement.reshape(mesh.n_nodes, dim)
    for d, side in faces:
        Vs, Ds = lagrange_table(p, [float(side)])
        cells = np.flatnonzero(mesh.cell_ijk[:, d] == (0 if side == 0 else mesh.ncells[d] - 1))
        tang = [e for e in range(dim) if e != d]
        # tensor tables at the face quadrature points: index order x fastest
        nfq = n ** (dim - 1)
        val = np.ones((nfq, n**dim))
        grads = np.ones((dim, nfq, n**dim))
        qidx = np.stack(np.meshgrid(*[np.arange(n)] * (dim - 1), indexing="ij")[::-1], axis=-1).reshape(nfq, dim - 1)
        aidx = mesh.loc_ijk  # [n^dim, dim] local node index per direction
        w = np.ones(nfq)
        for qi in range(nfq):
            for e in range(dim):
                if e == d:
                    v1, d1 = Vs[0], Ds[0]
                else:
                    t = tang.index(e)
                    v1, d1 = Vq[qidx[qi, t]], Dq[qidx[qi, t]]
                for g in range(dim):
                    grads[g, qi] *= (d1 if g == e else v1)[aidx[:, e]]
                val[qi] *= v1[aidx[:, e]]
            for t in range(dim - 1):
                w[qi] *= wq[qidx[qi, t]]
        area = np.prod([h[e] for e in tang])
        N = np.zeros(dim)
        N[d] = 1.0 if side == 1 else -1.0
        for c in cells:
            uloc = u[mesh.cell_nodes[c]]  # [n^dim, dim]
            for qi in range(nfq):
                gradu = np.einsum("ga,ac->cg", grads[:, qi], uloc) / h[None, :]  # du_c/dx_g
                F = np.eye(dim) + gradu
                n_star = np.linalg.det(F) * np.linalg.inv(F).T @ N
                t0 = np.asarray(traction) * np.linalg.norm(n_star)
                contrib = np.outer(val[qi], t0) * (w[qi] * area)
                np.add.at(rhs, (dim * mesh.cell_nodes[c][:, None] + np.arange(dim)[None, :]).ravel(), contrib.ravel())
    return rhs


def interface_traction_rhs_q1(mesh, total_displacement, traction_q, cells, face_no):
    """The same face loop (:1849-1897) for general Q1 cells (mapped boxes, the tube) and a traction given per face
    quadrature point: traction_q [n_faces, nq^2, dim], one vector, or a callable x -> t(x) evaluated at the quadrature
    point's position in the reference configuration (what the coupling partner's mapped data is).  Face f = (cells[f],
    face_no[f] = 2*direction + side); quadrature points of a face are ordered q = q_t1 + nq*q_t2 with the tangential
    directions t1 < t2.  3-D only."""
    dim, p, n = mesh.dim, mesh.p, mesh.p + 1
    assert dim == 3
    xq, wq = NO.gauss(n)
    Vq, Dq = lagrange_table(p, xq)
    rhs = np.zeros(mesh.n_nodes * dim)
    u = total_displacement.reshape(mesh.n_nodes, dim)
    tfun = traction_q if callable(traction_q) else None
    if tfun is None:
        traction_q = np.asarray(traction_q, dtype=float)
    aidx = mesh.loc_ijk
    for f, (c, fn) in enumerate(zip(cells, face_no)):
        d, side = fn // 2, fn % 2
        t1, t2 = [e for e in range(dim) if e != d]
        Vs, Ds = lagrange_table(p, [float(side)])
        verts = mesh.vertices[c]  # [8, 3], vertex m at (m&1, m>>1&1, m>>2&1)
        uloc = u[mesh.cell_nodes[c]]
        for q2 in range(n):
            for q1 in range(n):
                tabs = {d: (Vs[0], Ds[0]), t1: (Vq[q1], Dq[q1]), t2: (Vq[q2], Dq[q2])}
                val = np.ones(n**dim)
                grads = np.ones((dim, n**dim))
                for e in range(dim):
                    v1, d1 = tabs[e]
                    val *= v1[aidx[:, e]]
                    for g in range(dim):
                        grads[g] *= (d1 if g == e else v1)[aidx[:, e]]
                xi = np.zeros(dim)
                xi[d], xi[t1], xi[t2] = side, xq[q1], xq[q2]
                J = np.zeros((dim, dim))  # dx_e / dxi_f
                xpos = np.zeros(dim)
                for m in range(8):
                    bits = [(m >> e) & 1 for e in range(dim)]
                    l = [xi[e] if bits[e] else 1 - xi[e] for e in range(dim)]
                    dl = [1.0 if bits[e] else -1.0 for e in range(dim)]
                    g = np.array([dl[0] * l[1] * l[2], l[0] * dl[1] * l[2], l[0] * l[1] * dl[2]])
                    J += np.outer(verts[m], g)
                    xpos += l[0] * l[1] * l[2] * verts[m]
                Jinv = np.linalg.inv(J)
                nref = np.zeros(dim)
                nref[d] = 1.0 if side else -1.0
                nA = np.linalg.det(J) * Jinv.T @ nref  # Nanson: N dA on the reference configuration
                dA = np.linalg.norm(nA)
                N = nA / dA
                gradu = (grads @ uloc).T @ Jinv  # du_c/dx_e = sum_f dU_c/dxi_f dxi_f/dx_e
                F = np.eye(dim) + gradu
                n_star = np.linalg.det(F) * np.linalg.inv(F).T @ N
                t = np.asarray(tfun(xpos), dtype=float) if tfun is not None else (traction_q if traction_q.ndim == 1 else traction_q[f, q1 + n * q2])
                t0 = t * np.linalg.norm(n_star)
                contrib = np.outer(val, t0) * (wq[q1] * wq[q2] * dA)
                np.add.at(rhs, (dim * mesh.cell_nodes[c][:, None] + np.arange(dim)[None, :]).ravel(), contrib.ravel())
    return rhs


class SolidOracle:
    def __init__(self, meshes, mu, nu, rho, delta_t, formulation=1, caching="tensor2", body_force=None, traction=None,
                 traction_faces=(), preconditioner="gmg", relaxation=0.65, tol_lin=1e-6, tol_f=1e-9, tol_u=1e-6,
                 tol_f_abs=2e-9, tol_u_abs=5e-15, max_newton=10, n_cheb_cg=30, level_dtype=np.float32, max_cg_multiplier=1,
                 coarse_solver="chebyshev"):
        """meshes: level hierarchy (coarsest first, finest = active mesh) of StructuredMesh with clamp() applied"""
        self.meshes, self.mesh = meshes, meshes[-1]
        m = self.mesh
        self.dim = m.dim
        self.dt = delta_t
        beta, gamma = 0.25, 0.5  # mf_elasticity.h:237-238
        self.a1 = 1.0 / (beta * delta_t**2)
        self.a2 = 1.0 / (beta * delta_t)
        self.a3 = (1 - 2 * beta) / (2 * beta)
        self.a4 = gamma / (beta * delta_t)
        self.a5 = 1 - gamma / beta
        self.a6 = (1 - gamma / (2 * beta)) * delta_t
        self.material = NO.Material(mu, nu, rho, self.a1, formulation)
        self.caching = caching
        self.body_force = np.zeros(self.dim) if body_force is None else np.asarray(body_force, float)
        self.traction, self.traction_faces = traction, tuple(traction_faces)
        self.precond, self.relax, self.tol_lin = preconditioner, relaxation, tol_lin
        self.tol_f, self.tol_u, self.tol_f_abs, self.tol_u_abs, self.max_newton = tol_f, tol_u, tol_f_abs, tol_u_abs, max_newton
        self.n_cheb_cg, self.level_dtype = n_cheb_cg, level_dtype
        self.max_cg_multiplier = max_cg_multiplier  # 'Max iteration multiplier' (parameter_handling.h:203-207)
        self.coarse_solver = coarse_solver  # 'MF Chebyshev coarse' (mf_elasticity.h:1147): "chebyshev" | "cg"
        self.cells = [self._cells(mm) for mm in meshes]
        n = m.n_nodes * self.dim
        self.con = m.constrained_dofs(self.dim)
        self.total = np.zeros(n)
        self.delta = np.zeros(n)
        self.velocity_old = np.zeros(n)
        self.acceleration = np.zeros(n)
        self.acceleration_old = np.zeros(n)
        self.op = NO.NeoHookOperator(self.cells[-1], self.material)
        self.level_ops = [NO.NeoHookOperator(c, self.material) for c in self.cells]
        if len(meshes) > 1:
            self.transfers = [SO.MGTransfer(meshes[l - 1], meshes[l], self.dim, meshes[l - 1].constrained_dofs(self.dim),
                                            meshes[l].constrained_dofs(self.dim)) for l in range(1, len(meshes))]
        self.cg_iterations, self.cg_solves, self.newton_iterations, self.cg_history = 0, 0, [], []

    def _cells(self, mm):
        nq = mm.p + 1
        invjac, jxw, _ = NO.q1_metric(mm.vertices, nq)
        return NO.CellData(mm.p, nq, mm.dim, mm.dim, mm.dof_index(mm.dim), mm.constrained_dofs(mm.dim), invjac, jxw, mm.material_id)

    # ---- assemble_residual -------------------------------------------------------------------
    def assemble_residual(self):
        rhs = NO.nh_residual(self.cells[-1], self.material, self.total, self.acceleration, self.body_force)
        if self.traction is not None and (callable(self.traction) or self.mesh.mapping is not None):
            cells, fno = [], []
            for d, side in self.traction_faces:
                on = np.flatnonzero(self.mesh.cell_ijk[:, d] == (0 if side == 0 else self.mesh.ncells[d] - 1))
                cells += list(on)
                fno += [2 * d + side] * on.size
            t = interface_traction_rhs_q1(self.mesh, self.total, self.traction, cells, fno)
            t[self.con] = 0.0
            rhs += t
        elif self.traction is not None:
            t = interface_traction_rhs(self.mesh, self.total, self.traction, self.traction_faces)
            t[self.con] = 0.0  # distribute_local_to_global skips constrained rows
            rhs += t
        rhs[self.con] = 0.0  # constraints.set_zero(system_rhs) :1905
        return rhs

    # ---- update_matrix_free (:1289-1353) + setup_gmg(false) -------------------------------------
    def update_matrix_free(self):
        self.op.initialize(self.total, self.caching)
        self.op.cache()
        self.op.compute_diagonal()
        if self.precond != "gmg":
            return
        # setup_mg_interpolation: interpolate_to_mg of the fine displacement to every level (float levels)
        T = self.level_dtype
        u = [None] * len(self.meshes)
        u[-1] = self.total.astype(T).astype(np.float64)
        for l in range(len(self.meshes) - 1, 0, -1):
            uc = np.zeros(self.meshes[l - 1].n_nodes * self.dim)
            self.transfers[l - 1].interpolate(uc, u[l])
            u[l - 1] = uc.astype(T).astype(np.float64)
        As, sms = [], []
        for l, lop in enumerate(self.level_ops):
            lop.initialize(u[l], self.caching)
            lop.cache()
            d = lop.compute_diagonal()

            def A(dst, src, lop=lop):
                tmp = np.zeros(src.size)
                lop.vmult(tmp, np.asarray(src, dtype=np.float64))
                dst[:] = tmp

            As.append(A)
            con = self.meshes[l].constrained_dofs(self.dim)
            if l == 0 and self.coarse_solver == "chebyshev":  # mf_elasticity.h:1154-1166: solver mode, accurate eigenvalues -> m() CG iterations
                sms.append(SO.PreconditionChebyshev(A, 1.0 / d, con, degree=SO.INVALID_UNSIGNED_INT, smoothing_range=1e-3,
                                                    eig_cg_n_iterations=d.size, dtype=T))
            else:       # :1168-1174
                sms.append(SO.PreconditionChebyshev(A, 1.0 / d, con, degree=5, smoothing_range=20.0,
                                                    eig_cg_n_iterations=self.n_cheb_cg, dtype=T))
        coarse_cg = None
        if self.coarse_solver == "cg":  # MGCoarseGridIterativeSolver + ReductionControl(max(m(), 100), 1e-10, 1e-3), :1183-1198
            coarse_cg = (max(self.meshes[0].n_nodes * self.dim, 100), 1e-10, 1e-3)
        self.mg = SO.PreconditionMG(As, sms, self.transfers, dtype=T, coarse_cg=coarse_cg)

    def _A(self, dst, src):
        self.op.vmult(dst, src)

    # ---- solve_linear_system (:1981-2070) --------------------------------------------------------
    def solve_linear_system(self, rhs):
        x = np.zeros_like(rhs)
        tol = self.tol_lin * np.linalg.norm(rhs)
        max_it = rhs.size * self.max_cg_multiplier
        if self.precond == "direct":  # not a reference option: used to pin the nonlinear solution alone
            n = rhs.size
            cols = np.zeros((n, n))
            e = np.zeros(n)
            for i in range(n):
                e[:] = 0
                e[i] = 1
                self.op.vmult(cols[:, i], e)
            x = np.linalg.solve(cols, rhs)
            its = 0
        else:
            if self.precond == "jacobi":
                inv = self.op.inverse_diagonal

                def P(dst, src):
                    dst[:] = self.relax * inv * src
            elif self.precond == "gmg":
                def P(dst, src):
                    out = np.zeros(src.size, dtype=self.level_dtype)
                    self.mg.vmult(out, src)
                    dst[:] = out
            else:
                P = None
            r = SO.solver_cg(self._A, x, rhs, P, tol, max_it)
            assert r.converged
            its = r.iterations
        x[self.con] = 0.0  # constraints.set_zero(newton_update)
        return x, its

    # ---- solve_nonlinear_timestep (:1400-1491) ---------------------------------------------------
    def solve_nonlinear_timestep(self):
        self.delta[:] = 0.0
        err_res0 = err_upd0 = None
        for it in range(self.max_newton):
            self.acceleration = self.a1 * self.delta - self.a2 * self.velocity_old - self.a3 * self.acceleration_old  # :1612-1620
            rhs = self.assemble_residual()
            err_res = np.linalg.norm(rhs)
            if it == 0:
                err_res0 = err_res
            if it > 2:  # at least 3 iterations, :1372
                conv = err_res <= self.tol_f_abs or err_upd <= self.tol_u_abs
                if err_res / err_res0 <= self.tol_f and err_upd / err_upd0 <= self.tol_u:
                    conv = True
                if conv:
                    self.newton_iterations.append(it)
                    return it
            self.update_matrix_free()
            dx, its = self.solve_linear_system(rhs)
            self.cg_iterations += its
            self.cg_solves += 1
            self.cg_history.append(its)
            err_upd = np.linalg.norm(dx)
            if it == 0:
                err_upd0 = err_upd
            self.delta += dx
            self.total += dx
        raise RuntimeError("No convergence in nonlinear solver!")

    def advance_time_window(self):
        """post-step part of run() :595-600: update_velocity"""
        velocity = self.a4 * self.delta + self.a5 * self.velocity_old + self.a6 * self.acceleration_old
        self.velocity_old = velocity
        self.acceleration_old = self.acceleration.copy()

    def run(self, n_steps):
        for _ in range(n_steps):
            self.solve_nonlinear_timestep()
            self.advance_time_window()

    def displacement_at_node(self, node):
        return self.total.reshape(-1, self.dim)[node]
