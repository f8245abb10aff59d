"""Device-resident backward-Euler time step of the heat solver on top of the C ABI.

The call sequence of `LaplaceProblem<dim>` (include/heat_transfer/heat_transfer.h) restated on DeviceVectors:

  assemble_rhs                :577-655   lifting -A_inhom u_bc (:586-592), cell loop (u_old + dt f, phi) (:599-615), compress (:654)
  solve / solve_gmg_preconditioner :724-812  per step: level diagonals (:784), Chebyshev smoothers (degree 5, range 15,
                                             10 eigenvalue CG iterations; level 0: solver mode 1e-3, m() iterations, :769-783),
                                             V-cycle PreconditionMG, SolverCG to 1e-12 * ||rhs|| (:719-720)
  solution += solution_update :748-752

What the reference rebuilds in every time step - MGTransferMatrixFree::build (:759-760) and the smoother objects - is
built ONCE here: the mesh does not change between steps, the transfer depends on nothing else, and only the level
diagonals and eigenvalue estimates depend on (dt, coefficient).  That removes the 1.2 s of host set-up per step that
round 1 measured around a 0.18 s solve (profiles/r01_v3_solver_bench.jsonl).
"""
from __future__ import annotations

import numpy as np

from .operators import DeviceVector, LaplaceOperator, MatrixFree
from .solvers import INVALID_UNSIGNED_INT, MGTransferMatrixFree, PreconditionChebyshev, PreconditionMG, SolverCG


class HeatProblem:
    def __init__(self, meshes, delta_t, coefficient=None, source_const=0.0, tol=1e-12):
        """meshes: level hierarchy (coarsest first), Dirichlet faces clamp()ed with one component; homogeneous boundary values
        (the lifting of inhomogeneous ones is a vmult of an unconstrained operator, LaplaceOperator.vmult on such a context)."""
        from .mesh import children_of

        self.meshes, self.mesh = list(meshes), meshes[-1]
        self.dt, self.source_const, self.tol = float(delta_t), float(source_const), float(tol)
        self.mf = MatrixFree.from_mesh(self.mesh, 1)
        self.system_matrix = LaplaceOperator(self.mf, self.dt, coefficient)
        self.level_mf = [MatrixFree.from_mesh(m, 1, dtype="f32") for m in self.meshes]
        self.mg_matrices = [LaplaceOperator(mf, self.dt) for mf in self.level_mf]
        self.transfers = [MGTransferMatrixFree(self.level_mf[l - 1], self.level_mf[l], children_of(self.meshes[l - 1]))
                          for l in range(1, len(self.meshes))]
        self.smoothers, self.mg = [], None
        V = lambda: DeviceVector(self.mf)  # noqa: E731
        self.solution, self.solution_old, self.solution_update, self.system_rhs = V(), V(), V(), V()
        self.cg = SolverCG(0.0, self.mesh.n_nodes)
        self.total_n_cg_iterations, self.total_n_cg_solve, self.cg_history = 0, 0, []

    def set_delta_t(self, dt):
        self.dt = float(dt)
        self.system_matrix.set_delta_t(dt)
        for op in self.mg_matrices:
            op.set_delta_t(dt)

    # ---- assemble_rhs (:577-655), homogeneous boundary values
    def assemble_rhs(self):
        self.system_rhs.set_zero()
        self.system_matrix.assemble_rhs(self.system_rhs, self.solution_old, None, self.source_const)
        self.system_rhs.compress_add()
        self.system_rhs.set_zero_constrained()

    # ---- solve_gmg_preconditioner (:757-812) without the per-step rebuild of transfer and smoother objects
    def setup_preconditioner(self):
        for op in self.mg_matrices:
            op.compute_diagonal()  # :784
        if not self.smoothers:
            for l, op in enumerate(self.mg_matrices):
                n_level = self.meshes[l].n_nodes
                if l == 0:  # :780-782
                    self.smoothers.append(PreconditionChebyshev(op, degree=INVALID_UNSIGNED_INT, smoothing_range=1e-3,
                                                                eig_cg_n_iterations=n_level, n_global_dofs=n_level))
                else:       # :774-776
                    self.smoothers.append(PreconditionChebyshev(op, degree=5, smoothing_range=15.0, eig_cg_n_iterations=10,
                                                                n_global_dofs=n_level))
            self.mg = PreconditionMG(self.smoothers, self.transfers, self.mf)
        self.eig = [s.estimate_eigenvalues() for s in self.smoothers]

    def solve(self):
        self.setup_preconditioner()
        self.cg.tolerance = self.tol * self.system_rhs.l2_norm()  # :719-720
        self.solution_update.set_zero()
        res = self.cg.solve(self.system_matrix, self.solution_update, self.system_rhs, self.mg, x_is_zero=True)
        self.solution_update.set_zero_constrained()       # :746
        self.solution.sadd(0.0, 1.0, self.solution_update)  # homogeneous boundary values: solution = update (:748-750)
        self.total_n_cg_iterations += res["iterations"]
        self.total_n_cg_solve += 1
        self.cg_history.append(res["iterations"])
        return res

    # ---- coupling: interface faces of assemble_rhs (:617-652) and evaluate_boundary_flux (:899-940)
    def setup_interface(self, faces):
        """faces: (direction, side) pairs of the box that carry the coupling interface id.  Builds what the reference keeps
        next to the constrained system: the unconstrained MatrixFree + `inhomogeneous_operator` (:586-592, :907), the face
        set of the interface and the BoundaryMassOperator of the projector."""
        import copy

        from .operators import BoundaryMassOperator, InterfaceTraction

        free = copy.copy(self.mesh)
        free.constrained_nodes = None
        self.mesh_inhom = free
        self.mf_inhom = MatrixFree.from_mesh(free, 1)
        self.inhomogeneous_operator = LaplaceOperator(self.mf_inhom, self.dt)
        self.interface = InterfaceTraction(self.mf_inhom, free, faces)
        self.boundary_mass = BoundaryMassOperator(self.interface)
        W = lambda: DeviceVector(self.mf_inhom)  # noqa: E731
        self.flux, self._u, self._uo = W(), W(), W()

    def add_interface_flux(self, flux_q):
        """Neumann participant (:617-652): system_rhs += int_Gamma (-flux(q) dt) phi dA with the flux read from the coupling
        partner at the face quadrature points; call between the cell loop and compress (assemble_rhs(flux_q=...))."""
        self.interface.set_traction(np.asarray(flux_q, dtype=np.float64).reshape(-1, 1) if np.size(flux_q) > 1 else np.array([float(flux_q)]))
        self.interface.assemble_flux(self.system_rhs, -self.dt)

    def evaluate_boundary_flux(self):
        """Dirichlet participant (:899-940): residual = (A_inhom u - M (u_old + dt f)) / dt, L2-projected onto the interface
        (BoundaryProjector).  Returns the device vector holding the nodal flux and the projector's CG record."""
        r = self.flux
        self._u.copy_from(self.solution)
        self._uo.copy_from(self.solution_old)
        self.inhomogeneous_operator.set_delta_t(self.dt)
        self.inhomogeneous_operator.vmult(r, self._u)                                       # :907
        r.sadd(-1.0, 0.0, self._u)                                                          # signs switched w.r.t. assemble_rhs (:905-906)
        self.inhomogeneous_operator.assemble_rhs(r, self._uo, None, self.source_const)      # :916-931 (with + sign, see above)
        r.compress_add()                                                                    # :932
        r.sadd(-1.0 / self.dt, 0.0, self._u)                                                # :933
        info = self.boundary_mass.project(r)                                                # :935-939
        return r, info

    def time_step(self):
        """one pass of the time loop (:861-940): old <- solution, assemble, solve"""
        self.solution_old.sadd(0.0, 1.0, self.solution)
        self.assemble_rhs()
        return self.solve()
