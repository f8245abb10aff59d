"""Device-resident Newmark/Newton loop of the solid solver on top of the C ABI.

This is the call sequence the reference's `Solid<dim,degree,Number>` issues on its operators, restated on
DeviceVectors so that no vector leaves the GPU inside a time step:

  run / time loop                 include/solid_mechanics/mf_elasticity.h:553-628
  solve_nonlinear_timestep        :1400-1491   (check_convergence :1357-1395)
  update_acceleration / velocity  :1612-1620 / :1600-1608   (Newmark constants :237-245)
  assemble_residual (fast path)   :1764-1846 cells, :1849-1897 interface faces, :1899-1907 set_zero + norm
  update_matrix_free + setup_gmg  :1289-1353, :1140-1279, interpolate_to_mg :881-929
  solve_linear_system             :1981-2070

It is the host-side driver used by the parity tests (against oracle/solid_driver.py, which follows the same lines) and by
bench.py's `newton_step` block; the reference's own driver (parameter files, preCICE coupling, output) stays where it is.
Every arithmetic step is a library call: there is no numpy arithmetic on the vectors here.
"""
from __future__ import annotations

import numpy as np

from .operators import DeviceVector, MatrixFree, NeoHookOperator
from .solvers import INVALID_UNSIGNED_INT, MGTransferMatrixFree, PreconditionChebyshev, PreconditionMG, SolverCG


class Solid:
    def __init__(self, meshes, mu, nu, rho, delta_t, formulation=1, caching="tensor2", body_force=None, traction=None,
                 traction_faces=(), preconditioner="gmg", relaxation=0.65, tol_lin=1e-6, tol_f=1e-9, tol_u=1e-6,
                 tol_f_abs=2e-9, tol_u_abs=5e-15, max_newton=10, n_cheb_cg=30, max_cg_multiplier=1, children=None,
                 matrix_frees=None, coarse_solver="chebyshev"):
        """meshes: level hierarchy (coarsest first, finest = active mesh), clamp() applied.  children[l]: [n_cells_l][8]
        children of level-l cells on level l+1 (default: mesh.children_of).  matrix_frees: optional prebuilt
        (system MatrixFree, [level MatrixFree]) e.g. from partitions."""
        from .mesh import children_of

        self.meshes, self.mesh = list(meshes), meshes[-1]
        m = self.mesh
        beta, gamma = 0.25, 0.5  # mf_elasticity.h:237-238
        self.a1 = 1.0 / (beta * delta_t**2)
        self.a2 = 1.0 / (beta * delta_t)
        self.a3 = (1 - 2 * beta) / (2 * beta)
        self.a4 = gamma / (beta * delta_t)
        self.a5 = 1 - gamma / beta
        self.a6 = (1 - gamma / (2 * beta)) * delta_t
        self.material = dict(mu=mu, nu=nu, rho=rho, alpha_1=self.a1, formulation=formulation)
        self.caching = caching
        self.body_force = (0.0, 0.0, 0.0) if body_force is None else tuple(float(x) for x in body_force)
        # a constant vector, or a callable x -> t(x) evaluated once at the face quadrature points (the role of the data the
        # coupling partner writes on the interface mesh, adapter.h)
        self.traction = traction if (traction is None or callable(traction)) else tuple(float(x) for x in traction)
        self.traction_faces = tuple(traction_faces)
        self.precond, self.relax, self.tol_lin = preconditioner, relaxation, tol_lin
        self.tol_f, self.tol_u, self.tol_f_abs, self.tol_u_abs, self.max_newton = tol_f, tol_u, tol_f_abs, tol_u_abs, max_newton
        self.n_cheb_cg, self.max_cg_multiplier, self.coarse_solver = n_cheb_cg, max_cg_multiplier, coarse_solver

        # ---- setup_matrix_free (:806-1050): FP64 system + FP32 levels
        if matrix_frees is None:
            self.mf = MatrixFree.from_mesh(m, 3)
            self.level_mf = [MatrixFree.from_mesh(mm, 3, dtype="f32") for mm in self.meshes] if preconditioner == "gmg" else []
        else:
            self.mf, self.level_mf = matrix_frees
        self.n_global = getattr(m, "n_global_dofs", 3 * m.n_nodes)
        V = lambda: DeviceVector(self.mf)  # noqa: E731
        self.total, self.delta, self.newton_update = V(), V(), V()
        self._cg = None
        self.velocity, self.velocity_old = V(), V()
        self.acceleration, self.acceleration_old = V(), V()
        self.rhs = V()
        self.op = NeoHookOperator()
        self.op.set_material(self.mf, self.material)
        self.op.initialize(self.total, caching)
        self.level_ops, self.level_disp, self.transfers, self.smoothers = [], [], [], []
        self.mg = None
        if preconditioner == "gmg":
            for l, mf in enumerate(self.level_mf):
                op = NeoHookOperator()
                op.set_material(mf, self.material)
                d = DeviceVector(mf)
                op.initialize(d, caching)
                self.level_ops.append(op)
                self.level_disp.append(d)
                if l > 0:
                    ch = children_of(self.meshes[l - 1]) if children is None else children[l - 1]
                    self.transfers.append(MGTransferMatrixFree(self.level_mf[l - 1], mf, ch))
        self.cg_iterations, self.cg_solves, self.newton_iterations, self.cg_history = 0, 0, [], []
        self.face_op = None
        if self.traction is not None:
            from .operators import InterfaceTraction

            self.face_op = InterfaceTraction(self.mf, m, self.traction_faces)
            if callable(self.traction):
                pts = self.face_op.quadrature_points()
                self.face_op.set_traction(np.array([self.traction(x) for x in pts], dtype=np.float64))
            else:
                self.face_op.set_traction(self.traction)

    # ---- update_acceleration (:1612-1620): a = a1*delta - a2*v_old - a3*a_old
    def update_acceleration(self):
        a = self.acceleration
        a.sadd(0.0, self.a1, self.delta)
        a.sadd(1.0, -self.a2, self.velocity_old)
        a.sadd(1.0, -self.a3, self.acceleration_old)

    # ---- assemble_residual (:1647-1907, fast path) + constraints.set_zero + l2_norm
    def assemble_residual(self):
        r = self.rhs
        r.set_zero()
        self.op.assemble_residual(r, self.total, self.acceleration, self.body_force)
        if self.face_op is not None:
            self.face_op.assemble(r, self.total)  # :1849-1897
        r.compress_add()          # :1899
        r.set_zero_constrained()  # :1905
        return r.l2_norm()        # :1907

    # ---- update_matrix_free (:1289-1353) + setup_gmg(false) (:1140-1279)
    def update_matrix_free(self):
        self.op.cache()
        self.op.compute_diagonal()
        if self.precond != "gmg":
            return
        self.level_disp[-1].copy_from(self.total)  # interpolate_to_mg (:881-929): FP64 -> FP32, then level by level
        for l in range(len(self.level_mf) - 1, 0, -1):
            self.transfers[l - 1].interpolate(self.level_disp[l - 1], self.level_disp[l])
        for op in self.level_ops:
            op.cache()
            op.compute_diagonal()
        if not self.smoothers:
            for l, op in enumerate(self.level_ops):
                n_level = 3 * self.meshes[l].n_nodes if not hasattr(self.meshes[l], "n_global_dofs") else self.meshes[l].n_global_dofs
                first = getattr(self.meshes[l], "first_global_dof", 0)  # partitioned levels: this rank's offset in the global enumeration
                if l == 0 and self.coarse_solver == "chebyshev":  # :1154-1166 (accurate eigenvalues: m() CG iterations)
                    self.smoothers.append(PreconditionChebyshev(op, degree=INVALID_UNSIGNED_INT, smoothing_range=1e-3,
                                                                eig_cg_n_iterations=n_level, n_global_dofs=n_level,
                                                                first_global_index=first))
                else:       # :1168-1174
                    self.smoothers.append(PreconditionChebyshev(op, degree=5, smoothing_range=20.0,
                                                                eig_cg_n_iterations=self.n_cheb_cg, n_global_dofs=n_level,
                                                                first_global_index=first))
            self.eig = [s.estimate_eigenvalues() for s in self.smoothers]
            self.mg = PreconditionMG(self.smoothers, self.transfers, self.mf)
            if self.coarse_solver == "cg":  # 'MF Chebyshev coarse = false': MGCoarseGridIterativeSolver (:1183-1198)
                n0 = getattr(self.meshes[0], "n_global_dofs", 3 * self.meshes[0].n_nodes)
                self.mg.set_coarse_solver_cg(max(n0, 100), 1e-10, 1e-3)
        else:
            self.eig = [s.estimate_eigenvalues() for s in self.smoothers]

    # ---- solve_linear_system (:1981-2070)
    def solve_linear_system(self, rhs_norm):
        x = self.newton_update
        x.set_zero()
        if self._cg is None:  # one solver object (its work vectors live on the device) for all Newton iterations
            self._cg = SolverCG(0.0, int(self.n_global * self.max_cg_multiplier))
        cg = self._cg
        cg.tolerance = self.tol_lin * rhs_norm
        if self.precond == "gmg":
            P = self.mg
        elif self.precond == "jacobi":
            P = ("jacobi", self.relax)
        else:
            P = None
        res = cg.solve(self.op, x, self.rhs, P, x_is_zero=True)
        x.set_zero_constrained()  # :2059
        return res["iterations"]

    # ---- solve_nonlinear_timestep (:1400-1491)
    def solve_nonlinear_timestep(self):
        self.delta.set_zero()
        err_res0 = err_upd0 = err_upd = None
        for it in range(self.max_newton):
            self.update_acceleration()
            err_res = self.assemble_residual()
            if it == 0:
                err_res0 = err_res
            if it > 2:  # at least 3 iterations (:1372)
                conv = err_res <= self.tol_f_abs or err_upd <= self.tol_u_abs
                if err_res / err_res0 <= self.tol_f and err_upd / err_upd0 <= self.tol_u:
                    conv = True
                if conv:
                    self.newton_iterations.append(it)
                    return it
            self.update_matrix_free()
            its = self.solve_linear_system(err_res)
            self.cg_iterations += its
            self.cg_solves += 1
            self.cg_history.append(its)
            err_upd = self.newton_update.l2_norm()
            if it == 0:
                err_upd0 = err_upd
            self.delta.sadd(1.0, 1.0, self.newton_update)  # :1470-1475
            self.total.sadd(1.0, 1.0, self.newton_update)
        raise RuntimeError("No convergence in nonlinear solver!")

    def advance_time_window(self):
        """update_velocity (:1600-1608) and the old-state swap of run() (:595-600)."""
        v = self.velocity
        v.sadd(0.0, self.a4, self.delta)
        v.sadd(1.0, self.a5, self.velocity_old)
        v.sadd(1.0, self.a6, self.acceleration_old)
        self.velocity_old.sadd(0.0, 1.0, v)
        self.acceleration_old.sadd(0.0, 1.0, self.acceleration)

    def run(self, n_steps):
        for _ in range(n_steps):
            self.solve_nonlinear_timestep()
            self.advance_time_window()

    def total_displacement(self):
        return self.total.download()
