"""TEST INFRASTRUCTURE — numpy restatement ("naive twin") of the reference's
matrix-free operator path.  Dense per-cell basis tables, no sum factorisation,
dimension-generic (2-D and 3-D).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; the
product (the CUDA library behind include/b200mf.h) never does.

PARITY PIN STATUS: the reference holds no operator-level golden vectors
(SURVEY.md §8c).  This restatement is pinned (a) by self-consistency
(all cache levels agree, symmetry, finite differences of the residual,
scipy-assembled Laplace matrices) and (b) end-to-end against the reference's
only goldens, the 2-D Q1 Turek-Hron runs (tests/solid/*/*_serial.output),
through oracle/solid_driver.py + oracle/solver_oracle.py (tests/test_golden_turek.py:
all six goldens reproduced — displacement to 6 digits, Newton 4/4, CG 29/24/701).  deal.II semantics (FE_Q on Gauss-Lobatto points,
Gauss quadrature, lexicographic component-major cell layout, Jacobian
conventions) are restated from its documentation: SURVEY.md Appendix A.

All "reference file:line" citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial import legendre as _leg

# ---------------------------------------------------------------------------
# 1-D basis: FE_Q(p) on Gauss-Lobatto points, QGauss(nq) on [0,1]
# (deal.II semantics; used by include/base/fe_integrator.h:10-22)
# ---------------------------------------------------------------------------


def gll_nodes(p: int) -> np.ndarray:
    """p+1 Gauss-Lobatto points on [0,1]."""
    if p == 1:
        return np.array([0.0, 1.0])
    c = np.zeros(p + 1)
    c[p] = 1.0
    inner = _leg.legroots(_leg.legder(c))
    # Newton polish on P'_p
    for _ in range(3):
        d1 = _leg.legval(inner, _leg.legder(c))
        d2 = _leg.legval(inner, _leg.legder(c, 2))
        inner = inner - d1 / d2
    x = np.concatenate([[-1.0], np.sort(inner), [1.0]])
    x = 0.5 * (x - x[::-1])  # enforce symmetry
    return 0.5 * (x + 1.0)


def gauss(nq: int):
    """QGauss<1>(nq) on [0,1]: points, weights."""
    x, w = _leg.leggauss(nq)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_eval(nodes: np.ndarray, x: np.ndarray):
    """values[q,i] = l_i(x_q), derivs[q,i] = l_i'(x_q)."""
    n = len(nodes)
    V = np.zeros((len(x), n))
    D = np.zeros((len(x), n))
    for i in range(n):
        others = [j for j in range(n) if j != i]
        denom = np.prod([nodes[i] - nodes[j] for j in others])
        V[:, i] = np.prod([x - nodes[j] for j in others], axis=0) / denom
        s = np.zeros_like(x)
        for m in others:
            s += np.prod([x - nodes[j] for j in others if j != m], axis=0) if n > 2 else 1.0
        D[:, i] = s / denom
    return V, D


def shape_1d(p: int, nq: int):
    """S[q,i] = l_i(xi_q), D[q,i] = l_i'(xi_q), w[q]."""
    xq, w = gauss(nq)
    S, D = lagrange_eval(gll_nodes(p), xq)
    return S, D, w


def cell_tables(p: int, nq: int, dim: int):
    """Dense tables N[q,a], dN[f,q,a], w[q]; a and q lexicographic, x fastest
    (SURVEY.md A.1)."""
    S, D, w = shape_1d(p, nq)

    def kron_list(ms):  # ms[0] acts on x (fastest)
        out = ms[-1]
        for m in ms[-2::-1]:
            out = np.kron(out, m)
        return out

    N = kron_list([S] * dim)
    dN = np.stack([kron_list([D if g == f else S for g in range(dim)]) for f in range(dim)])
    W = kron_list([w.reshape(-1, 1)] * dim).ravel()
    return N, dN, W


# ---------------------------------------------------------------------------
# Q1 mapping metric (deal.II MappingQ1 / MatrixFree MappingInfo semantics, A.2)
# ---------------------------------------------------------------------------


def q1_metric(vertices: np.ndarray, nq: int):
    """vertices[cell, 2^dim, dim] (lexicographic vertex order, x fastest).
    Returns invjac[cell,q,e,f] = d xi_f / d x_e (deal.II inverse_jacobian,
    mf_nh_operator.h:872-887), JxW[cell,q], qpoints[cell,q,dim]."""
    nc, nv, dim = vertices.shape
    xq, w = gauss(nq)
    lin = np.stack([1.0 - xq, xq], axis=1)  # [q, 2]
    dlin = np.stack([-np.ones_like(xq), np.ones_like(xq)], axis=1)

    def kron_list(ms):
        out = ms[-1]
        for m in ms[-2::-1]:
            out = np.kron(out, m)
        return out

    Nv = kron_list([lin] * dim)  # [nq^dim, 2^dim]
    dNv = np.stack([kron_list([dlin if g == f else lin for g in range(dim)]) for f in range(dim)])
    W = kron_list([w.reshape(-1, 1)] * dim).ravel()
    # J[cell,q,e,f] = d x_e / d xi_f
    J = np.einsum("fqm,cme->cqef", dNv, vertices)
    Jinv = np.linalg.inv(J)  # Jinv[f,e] = d xi_f / d x_e
    invjac = np.swapaxes(Jinv, -1, -2)  # [e,f]
    JxW = np.linalg.det(J) * W[None, :]
    qp = np.einsum("qm,cme->cqe", Nv, vertices)
    return invjac, JxW, qp


# ---------------------------------------------------------------------------
# Material (include/solid_mechanics/material.h:78-94, 252-258)
# ---------------------------------------------------------------------------


class Material:
    def __init__(self, mu, nu, rho, alpha_1, formulation):
        self.kappa = (2.0 * mu * (1.0 + nu)) / (3.0 * (1.0 - 2.0 * nu))
        self.c_1 = mu / 2.0
        self.mu = mu
        self.rho = rho
        self.rho_alpha = rho * alpha_1
        self.lam = (2.0 * mu * nu) / (1.0 - 2.0 * nu)
        self.formulation = formulation

    def get_tau(self, F):
        """material.h:116-161 (Kirchhoff stress), F[...,d,d]."""
        dim = F.shape[-1]
        I = np.eye(dim)
        J = np.linalg.det(F)
        b = F @ np.swapaxes(F, -1, -2)
        if self.formulation == 0:
            Fbar = F * (J ** (-1.0 / dim))[..., None, None]
            bbar = Fbar @ np.swapaxes(Fbar, -1, -2)
            tmp = (self.kappa / 2.0) * (J - 1.0 / J) * J
            tau_bar = bbar * (2.0 * self.c_1)
            tr = np.trace(tau_bar, axis1=-2, axis2=-1)
            return (tmp - tr / dim)[..., None, None] * I + tau_bar
        tmp = self.mu - 2.0 * self.lam * np.log(J)
        return self.mu * b - tmp[..., None, None] * I


# ---------------------------------------------------------------------------
# Cell-data container handed to the operators ("what MatrixFree holds")
# ---------------------------------------------------------------------------


class CellData:
    """dof_idx[cell, ncomp*n^dim] component-major, lexicographic inside a
    component; constrained[global dof] bool;
    invjac/JxW from the *reference* mapping (StaticMappingQ1,
    mf_elasticity.h:739-746)."""

    def __init__(self, p, nq, dim, ncomp, dof_idx, constrained, invjac, JxW, material_id=None):
        self.p, self.nq1, self.dim, self.ncomp = p, nq, dim, ncomp
        self.dof_idx = np.asarray(dof_idx, dtype=np.int64)
        self.constrained = np.asarray(constrained, dtype=bool)
        self.invjac = invjac
        self.JxW = JxW
        self.n_cells = self.dof_idx.shape[0]
        self.n_dofs = self.constrained.shape[0]
        self.N, self.dN, self.W = cell_tables(p, nq, dim)
        self.n_q = self.N.shape[0]
        self.n_loc = self.N.shape[1]
        self.material_id = (
            np.zeros(self.n_cells, dtype=np.int64) if material_id is None else np.asarray(material_id)
        )

    # phi.read_dof_values (constraints resolved homogeneously, A.3)
    def read(self, v):
        loc = v[self.dof_idx]
        loc = np.where(self.constrained[self.dof_idx], 0.0, loc)
        return loc.reshape(self.n_cells, self.ncomp, self.n_loc)

    # phi.read_dof_values_plain
    def read_plain(self, v):
        return v[self.dof_idx].reshape(self.n_cells, self.ncomp, self.n_loc)

    # phi.distribute_local_to_global (constrained rows skipped)
    def scatter_add(self, dst, loc):
        loc = loc.reshape(self.n_cells, -1)
        keep = ~self.constrained[self.dof_idx]
        np.add.at(dst, self.dof_idx[keep], loc[keep])

    # phi.evaluate(values|gradients) followed by get_value/get_gradient
    def evaluate(self, loc, invjac):
        val = np.einsum("qa,kca->kqc", self.N, loc)
        G = np.einsum("fqa,kca->kqcf", self.dN, loc)  # d u_c / d xi_f
        grad = np.einsum("kqef,kqcf->kqce", invjac, G)  # get_gradient, §8a a3
        return val, grad

    # submit_value/submit_gradient + integrate
    def integrate(self, val, grad, invjac, JxW):
        out = 0.0
        if val is not None:
            out = out + np.einsum("qa,kqc,kq->kca", self.N, val, JxW)
        if grad is not None:
            Gs = np.einsum("kqef,kqce->kqcf", invjac, grad) * JxW[:, :, None, None]
            out = out + np.einsum("fqa,kqcf->kca", self.dN, Gs)
        return out


def _sym(g):
    return 0.5 * (g + np.swapaxes(g, -1, -2))


def _tr(g):
    return np.trace(g, axis1=-2, axis2=-1)


# ---------------------------------------------------------------------------
# NeoHookOperator (include/solid_mechanics/mf_nh_operator.h)
# ---------------------------------------------------------------------------


class NeoHookOperator:
    """Restates NeoHookOperator<dim,double>: initialize :331-388, cache
    :392-523, vmult_add :575-625, do_operation_on_cell :651-1019,
    compute_diagonal :1023-1060, precondition_Jacobi :281-291."""

    LEVELS = ("scalar_referential", "tensor2", "tensor4", "tensor4_ns")

    def __init__(self, cells: CellData, material: Material, material_inclusion: Material | None = None):
        self.cd = cells
        self.material = material
        self.material_inclusion = material_inclusion or material
        self.caching = None

    def initialize(self, displacement, caching):
        if caching not in self.LEVELS:
            raise ValueError("Unknown caching")  # :385
        self.caching = caching
        self.displacement = displacement
        self.frame = "referential" if caching in ("scalar_referential", "tensor4_ns") else "deformed"

    def _mat_arrays(self):
        """per-cell material constants (material_id==0 ? material : inclusion,
        :402-404, :656-658)."""
        sel = self.cd.material_id == 0
        g = lambda name: np.where(sel, getattr(self.material, name), getattr(self.material_inclusion, name))
        return {k: g(k)[:, None] for k in ("kappa", "c_1", "mu", "lam", "rho_alpha")}

    def cache(self):
        cd = self.cd
        dim = cd.dim
        I = np.eye(dim)
        m = self._mat_arrays()
        form = self.material.formulation
        loc = cd.read_plain(self.displacement)  # :410
        _, grad_u = cd.evaluate(loc, cd.invjac)  # :411, get_gradient :427
        F = I + grad_u  # Kinematics::F
        detF = np.linalg.det(F)
        if not np.all(detF > 0):
            raise ArithmeticError("det_F is not positive.")  # :430-435
        self.F, self.detF = F, detF
        # MappingQEulerian metric: J_cur = F J_ref
        Finv = np.linalg.inv(F)
        self.invjac_cur = np.swapaxes(Finv, -1, -2) @ cd.invjac
        self.JxW_cur = detF * cd.JxW
        if self.caching == "scalar_referential":
            self.cached_position = loc.copy()  # :416-418
        if form == 0:
            assert self.caching == "tensor2"
            self.cached_scalar = detF ** (-1.0 / dim)  # :437
            self.cached_tensor2 = F  # :438
            return
        scalar = m["mu"] - 2.0 * m["lam"] * np.log(detF)  # :455-456
        if self.caching == "scalar_referential":
            self.cached_scalar = scalar
        elif self.caching in ("tensor2", "tensor4"):
            b = F @ np.swapaxes(F, -1, -2)
            tau = m["mu"][..., None, None] * b - scalar[..., None, None] * I
            self.cached_second_scalar = 1.0 / detF
            self.cached_tensor2 = tau / detF[..., None, None]
            if self.caching == "tensor2":
                self.cached_scalar = scalar * 2.0 / detF
            else:
                S4 = 0.5 * (np.einsum("ik,jl->ijkl", I, I) + np.einsum("il,jk->ijkl", I, I))
                IxI = np.einsum("ij,kl->ijkl", I, I)
                self.cached_tensor4 = (scalar * 2.0 / detF)[..., None, None, None, None] * S4 + (
                    m["lam"] * 2.0 / detF
                )[..., None, None, None, None] * IxI  # :489-493
        elif self.caching == "tensor4_ns":
            FinvT = np.swapaxes(Finv, -1, -2)
            lnJ = np.log(detF)
            t1 = np.einsum("...ij,...kl->...ijkl", FinvT, FinvT)  # outer_product
            t2 = np.einsum("...il,...jk->...ijkl", FinvT, Finv)  # outer_product_iljk :23-46
            t3 = np.einsum("ik,jl->ijkl", I, I)  # IxI_ikjl :275
            self.cached_tensor4_ns = (
                (2.0 * m["lam"])[..., None, None, None, None] * t1
                + (m["mu"] - 2.0 * m["lam"] * lnJ)[..., None, None, None, None] * t2
                + m["mu"][..., None, None, None, None] * t3
            )  # :509-513

    # do_operation_on_cell :651-1019 on already-read local values
    def cell_op(self, loc):
        cd = self.cd
        dim = cd.dim
        I = np.eye(dim)
        m = self._mat_arrays()
        form = self.material.formulation
        if self.frame == "deformed":
            invjac, JxW = self.invjac_cur, self.JxW_cur
        else:
            invjac, JxW = cd.invjac, cd.JxW
        val, g = cd.evaluate(loc, invjac)  # :697
        rho_alpha = m["rho_alpha"][..., None]
        if form == 0:  # :701-840
            F = self.cached_tensor2
            detF = np.linalg.det(F)
            Fbar = F * self.cached_scalar[..., None, None]
            bbar = Fbar @ np.swapaxes(Fbar, -1, -2)
            s = _sym(g)
            kappa, c_1 = m["kappa"], m["c_1"]
            tau_bar = bbar * (2.0 * c_1)[..., None, None]
            tr_tau_bar = _tr(tau_bar)
            tr_tau_bar_dim = tr_tau_bar / dim
            dPsi = (kappa / 2.0) * (detF - 1.0 / detF)
            dPsiJ = dPsi * detF
            d2Psi = (kappa / 2.0) * (1.0 + 1.0 / (detF * detF))
            tau = (dPsiJ - tr_tau_bar_dim)[..., None, None] * I + tau_bar
            tr = _tr(s)
            dev = s - (tr / dim)[..., None, None] * I
            jc = s * (-dPsiJ * 2.0)[..., None, None]
            jc = jc + (detF * (dPsi + detF * d2Psi) * tr)[..., None, None] * I
            tau_iso = tau_bar - tr_tau_bar_dim[..., None, None] * I
            jc = jc + ((2.0 / dim) * tr_tau_bar)[..., None, None] * dev
            jc = jc - ((2.0 / dim) * tr)[..., None, None] * tau_iso
            tis = np.einsum("...ij,...ij->...", tau_iso, s)
            jc = jc - ((2.0 / dim) * tis)[..., None, None] * I
            inv = 1.0 / detF
            sub_g = (jc + g @ tau) * inv[..., None, None]  # :824-832
            sub_v = val * rho_alpha * inv[..., None]  # :833
        elif self.caching == "scalar_referential":  # :848-943
            # second evaluate of the cached displacement values :861-867
            _, gu = cd.evaluate(self.cached_position, cd.invjac)
            F = I + gu
            b = F @ np.swapaxes(F, -1, -2)
            tau = m["mu"][..., None, None] * b - self.cached_scalar[..., None, None] * I
            Finv = np.linalg.inv(F)
            gx = g @ Finv  # :910-922  grad_Nx_v[d][e] = sum_f F_inv[f][e] tmp[f]
            jc = (2.0 * self.cached_scalar)[..., None, None] * _sym(gx)
            jc = jc + (2.0 * m["lam"] * _tr(gx))[..., None, None] * I
            queued = jc + gx @ tau
            sub_g = queued @ np.swapaxes(Finv, -1, -2)  # :938
            sub_v = val * rho_alpha
        elif self.caching == "tensor2":  # :945-974
            s = _sym(g)
            jc = self.cached_scalar[..., None, None] * s
            jc = jc + (2.0 * m["lam"] * self.cached_second_scalar * _tr(s))[..., None, None] * I
            sub_g = jc + g @ self.cached_tensor2
            sub_v = val * rho_alpha * self.cached_second_scalar[..., None]
        elif self.caching == "tensor4":  # :976-994
            s = _sym(g)
            sub_g = g @ self.cached_tensor2 + np.einsum("...ijkl,...kl->...ij", self.cached_tensor4, s)
            sub_v = val * rho_alpha * self.cached_second_scalar[..., None]
        elif self.caching == "tensor4_ns":  # :996-1010
            sub_g = np.einsum("...ijkl,...kl->...ij", self.cached_tensor4_ns, g)
            sub_v = val * rho_alpha
        else:
            raise ValueError("Unknown caching")
        return cd.integrate(sub_v, sub_g, invjac, JxW)  # :1018

    def vmult_add(self, dst, src):
        """:575-625 (single rank: ghost update/compress are no-ops)."""
        loc = self.cd.read(src)
        self.cd.scatter_add(dst, self.cell_op(loc))
        c = self.cd.constrained
        dst[c] += src[c]  # :623-624

    def vmult(self, dst, src):
        dst[:] = 0.0
        self.vmult_add(dst, src)

    Tvmult = vmult  # :556-561

    def compute_diagonal(self):
        """:1023-1060 + MatrixFreeTools::compute_diagonal (A.4): apply the cell
        operator to each local unit vector, keep entry i."""
        cd = self.cd
        nl = cd.ncomp * cd.n_loc
        diag_loc = np.zeros((cd.n_cells, nl))
        for i in range(nl):
            e = np.zeros((cd.n_cells, nl))
            e[:, i] = 1.0
            out = self.cell_op(e.reshape(cd.n_cells, cd.ncomp, cd.n_loc)).reshape(cd.n_cells, nl)
            diag_loc[:, i] = out[:, i]
        diag = np.zeros(cd.n_dofs)
        cd.scatter_add(diag, diag_loc)
        diag[cd.constrained] = 1.0  # :1043-1044
        if not np.all(diag > 0):
            raise ArithmeticError("No diagonal entry in a positive definite operator should be zero or negative")
        self.diagonal = diag
        self.inverse_diagonal = 1.0 / diag
        return diag

    def precondition_Jacobi(self, dst, src, omega):
        dst[:] = omega * self.inverse_diagonal * src  # :281-291

    def el(self, row, col):
        assert row == col
        return self.diagonal[row]


def nh_residual(cd: CellData, material: Material, total_displacement, acceleration, body_force):
    """assemble_residual fast path, cells only
    (include/solid_mechanics/mf_elasticity.h:1764-1846), before compress /
    constraints.set_zero.  Used for the finite-difference pin of the tangent."""
    dim = cd.dim
    I = np.eye(dim)
    loc_u = cd.read_plain(total_displacement)
    loc_a = cd.read_plain(acceleration)
    _, grad_u = cd.evaluate(loc_u, cd.invjac)
    val_a, _ = cd.evaluate(loc_a, cd.invjac)
    F = I + grad_u
    tau = material.get_tau(F)
    Finv = np.linalg.inv(F)
    sub_g = -tau @ np.swapaxes(Finv, -1, -2)  # :1830
    sub_v = (-val_a + np.asarray(body_force)[None, None, :]) * material.rho  # :1831-1834
    loc = cd.integrate(sub_v, sub_g, cd.invjac, cd.JxW)
    rhs = np.zeros(cd.n_dofs)
    # phi.distribute_local_to_global(system_rhs): constrained rows skipped
    cd.scatter_add(rhs, loc)
    return rhs


# ---------------------------------------------------------------------------
# LaplaceOperator (include/heat_transfer/heat_transfer.h:81-250)
# ---------------------------------------------------------------------------


class LaplaceOperator:
    """(M + dt*c*K) u, scalar: do_operation_on_cell :234-250, apply_add
    :194-201 via MatrixFreeOperators::Base (constrained rows = identity),
    compute_diagonal :205-230."""

    def __init__(self, cells: CellData, delta_t: float, coefficient=None):
        assert cells.ncomp == 1
        self.cd = cells
        self.delta_t = delta_t
        self.coefficient = (
            np.ones((cells.n_cells, cells.n_q)) if coefficient is None else coefficient
        )  # evaluate_coefficient :150-166, Coefficient::value == 1 :64-71

    def cell_op(self, loc):
        cd = self.cd
        val, g = cd.evaluate(loc, cd.invjac)
        sub_g = (self.coefficient * self.delta_t)[..., None, None] * g  # :245-247
        return cd.integrate(val, sub_g, cd.invjac, cd.JxW)

    def vmult(self, dst, src):
        dst[:] = 0.0
        self.cd.scatter_add(dst, self.cell_op(self.cd.read(src)))
        c = self.cd.constrained
        dst[c] += src[c]  # Base::vmult: constrained diagonal = 1

    Tvmult = vmult

    def compute_diagonal(self):
        cd = self.cd
        nl = cd.n_loc
        diag_loc = np.zeros((cd.n_cells, nl))
        for i in range(nl):
            e = np.zeros((cd.n_cells, 1, nl))
            e[:, 0, i] = 1.0
            diag_loc[:, i] = self.cell_op(e)[:, 0, i]
        diag = np.zeros(cd.n_dofs)
        cd.scatter_add(diag, diag_loc)
        diag[cd.constrained] = 1.0  # set_constrained_entries_to_one :220
        self.diagonal = diag
        self.inverse_diagonal = 1.0 / diag
        return diag


def heat_rhs_cells(cd: CellData, delta_t, solution_old, source_q=None, source_const=0.0):
    """Cell loop of LaplaceProblem::assemble_rhs (include/heat_transfer/heat_transfer.h:599-615):
    rhs += int (u_old + dt * f(x_q)) phi, on the MatrixFree of the (unconstrained) inhomogeneous operator;
    source_q [n_cells, n_q] = f at the quadrature points (evaluate_function on phi.quadrature_point(q))."""
    loc = cd.read(solution_old)
    val, g = cd.evaluate(loc, cd.invjac)
    f = np.full((cd.n_cells, cd.n_q), float(source_const)) if source_q is None else np.asarray(source_q)
    sub_v = val + delta_t * f[..., None]  # val: [cell, q, component]
    loc_out = cd.integrate(sub_v, np.zeros_like(g), cd.invjac, cd.JxW)
    rhs = np.zeros(cd.n_dofs)
    cd.scatter_add(rhs, loc_out)
    return rhs

