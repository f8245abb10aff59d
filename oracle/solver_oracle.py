"""CPU restatement (numpy/scipy) of the deal.II solver classes the reference instantiates on its operators —
TEST INFRASTRUCTURE ONLY (checker for the device implementations in csrc/solver.cu; never on the product path).

None of this source lives under /root/reference: it is deal.II (>= 9.6; CI pins dealii/dealii:master-noble and
v9.6.2-noble, .github/workflows/solid_mechanics_ci.yml:16, backward_compatibility.yml:16).  The algorithms are restated
from deal.II's documented behaviour and anchored on the reference's call sites:

  PreconditionChebyshev   include/solid_mechanics/mf_elasticity.h:371-372, 1149-1180; include/heat_transfer/heat_transfer.h:764-788
  SolverCG                mf_elasticity.h:2024-2051; heat_transfer.h:724-737, 810-811
  MGTransferMatrixFree    mf_elasticity.h:863-877, 920-922; heat_transfer.h:761-762
  Multigrid/PreconditionMG mf_elasticity.h:1209-1224, 1272-1276

Parity status: PINNED against the reference's end-to-end goldens — oracle/solid_driver.py drives these classes through
the 2-D Q1 Turek-Hron runs and reproduces the average CG iterations per Newton step of all six cases (29 for the four
cache levels with GMG-Chebyshev, 24 for formulation 0, 701 with Jacobi; tests/solid/*/*_serial.output:19) —
tests/test_golden_turek.py.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.sparse as sp

from .np_oracle import gll_nodes, lagrange_eval

INVALID_UNSIGNED_INT = -1


# ---------------------------------------------------------------------------
# SolverCG (A.6)
class CGResult:
    def __init__(self):
        self.iterations = 0
        self.residual = 0.0
        self.initial_residual = 0.0
        self.converged = False
        self.diag, self.offdiag = [], []

    def eigenvalues(self):
        if not self.diag:
            return np.array([1.0])
        return scipy.linalg.eigvalsh_tridiagonal(np.array(self.diag), np.array(self.offdiag))


def solver_cg(A, x, b, precond=None, tol=1e-12, max_it=1000, iteration_number_control=False, dtype=np.float64):
    """A(dst, src) applies the operator; precond(dst, src) or None.  x is the initial guess and is updated in place."""
    res = CGResult()
    r = np.empty_like(b)
    if np.any(x):
        A(r, x)
        r[:] = b - r
    else:
        r[:] = b
    res_norm = float(np.sqrt(abs(np.dot(r.astype(np.float64), r.astype(np.float64)))))
    res.initial_residual = res_norm
    p = np.zeros_like(b)
    v = np.zeros_like(b)
    rPr = alpha = beta = 0.0

    def state(step, val):
        if val <= tol:
            return 1
        if step >= max_it:
            return 1 if iteration_number_control else -1
        return 0

    it = 0
    st = state(0, res_norm)
    while st == 0:
        it += 1
        old = rPr
        if precond is None:
            rPr = res_norm * res_norm
            direction = r
        else:
            precond(v, r)
            rPr = float(np.dot(r.astype(np.float64), v.astype(np.float64)))
            direction = v
        prev_alpha = alpha
        if it > 1:
            beta = rPr / old
            p[:] = direction + dtype(beta) * p
        else:
            p[:] = direction
        A(v, p)
        pAp = float(np.dot(p.astype(np.float64), v.astype(np.float64)))
        alpha = rPr / pAp
        x += dtype(alpha) * p
        r -= dtype(alpha) * v
        res_norm = float(np.sqrt(abs(np.dot(r.astype(np.float64), r.astype(np.float64)))))
        if it > 1:
            res.diag.append(1.0 / alpha + beta / prev_alpha)
            res.offdiag.append(np.sqrt(beta) / prev_alpha)
        else:
            res.diag.append(1.0 / alpha)
        st = state(it, res_norm)
    res.iterations, res.residual, res.converged = it, res_norm, st == 1
    return res


# ---------------------------------------------------------------------------
# PreconditionChebyshev (A.5)
class PreconditionChebyshev:
    def __init__(self, A, inv_diag, constrained, degree=1, smoothing_range=0.0, eig_cg_n_iterations=8, max_eigenvalue=1.0,
                 dtype=np.float64):
        self.A, self.inv, self.con = A, inv_diag.astype(dtype), constrained
        self.degree_in, self.range, self.n_cg, self.max_ev = degree, smoothing_range, eig_cg_n_iterations, max_eigenvalue
        self.dtype = dtype
        self.ready = False

    def estimate_eigenvalues(self):
        n = self.inv.size
        if self.n_cg > 0:
            rhs = (np.arange(n) % 11).astype(self.dtype)
            rhs -= self.dtype(rhs.astype(np.float64).sum() / n)
            rhs[self.con] = 0
            x = np.zeros(n, dtype=self.dtype)

            def P(dst, src):
                dst[:] = self.inv * src

            r = solver_cg(self.A, x, rhs, P, 1e-10, self.n_cg, True, self.dtype)
            ev = r.eigenvalues()
            mn, mx = float(ev[0]), 1.2 * float(ev[-1])
            self.cg_iterations = r.iterations
        else:
            mx = self.max_ev
            mn = mx / self.range
        alpha = mx / self.range if self.range > 1.0 else min(0.9 * mx, mn)
        self.degree = self.degree_in
        if self.degree_in < 0:
            actual = mx / alpha
            sigma = (1.0 - np.sqrt(1.0 / actual)) / (1.0 + np.sqrt(1.0 / actual))
            eps = self.range
            self.degree = 1 + int(np.log(1.0 / eps + np.sqrt(1.0 / eps / eps - 1.0)) / np.log(1.0 / sigma))
        self.min_eig, self.max_eig = mn, mx
        self.delta, self.theta = (mx - alpha) * 0.5, (mx + alpha) * 0.5
        self.ready = True

    def _apply(self, dst, src, zero_start):
        if not self.ready:
            self.estimate_eigenvalues()
        T = self.dtype
        t = np.empty_like(dst)
        if zero_start:
            old = np.zeros_like(dst)
            dst[:] = T(1.0 / self.theta) * self.inv * src
        else:
            self.A(t, dst)
            old = dst.copy()
            dst[:] = dst + T(1.0 / self.theta) * self.inv * (src - t)
        if self.degree < 2 or abs(self.delta) < 1e-40:
            return
        rhok, sigma = self.delta / self.theta, self.theta / self.delta
        for _ in range(self.degree - 1):
            rhokp = 1.0 / (2.0 * sigma - rhok)
            f1, f2 = rhokp * rhok, 2.0 * rhokp / self.delta
            rhok = rhokp
            self.A(t, dst)
            new = dst + T(f1) * (dst - old) + T(f2) * self.inv * (src - t)
            old[:] = dst
            dst[:] = new

    def vmult(self, dst, src):
        self._apply(dst, src, True)

    def step(self, dst, src):
        self._apply(dst, src, False)


# ---------------------------------------------------------------------------
# MGTransferMatrixFree (A.8) as explicit sparse matrices
def embedding_1d(p):
    """P[m, j]: value of coarse basis function j at the fine patch node m (2p+1 nodes of the two children);
    R[i, m]: value at coarse node i of the fine basis function sitting on patch node m (child containing the node)."""
    nd = gll_nodes(p)
    xm = np.concatenate([0.5 * nd, 0.5 + 0.5 * nd[1:]])
    P, _ = lagrange_eval(nd, xm)
    R = np.zeros((p + 1, 2 * p + 1))
    for i in range(p + 1):
        ch = 0 if nd[i] <= 0.5 else 1
        xi = np.array([2.0 * nd[i] - ch])
        V, _ = lagrange_eval(nd, xi)
        R[i, ch * p:ch * p + p + 1] = V[0]
    return P, R


def children_of(coarse_mesh):
    """[n_coarse_cells, 2^dim] fine cell indices (x fastest) for a uniformly refined StructuredMesh."""
    nc = coarse_mesh.ncells
    dim = len(nc)
    ijk = coarse_mesh.cell_ijk
    out = np.zeros((coarse_mesh.n_cells, 1 << dim), dtype=np.int64)
    for ch in range(1 << dim):
        idx = np.zeros(coarse_mesh.n_cells, dtype=np.int64)
        stride = 1
        for d in range(dim):
            idx += (2 * ijk[:, d] + ((ch >> d) & 1)) * stride
            stride *= 2 * nc[d]
        out[:, ch] = idx
    return out


class MGTransfer:
    """Prolongation matrix between two StructuredMesh levels for an ncomp-component interleaved numbering."""

    def __init__(self, coarse_mesh, fine_mesh, ncomp, coarse_constrained, fine_constrained):
        p, dim = coarse_mesh.p, coarse_mesh.dim
        n, M = p + 1, 2 * p + 1
        P1, R1 = embedding_1d(p)
        Pk, Rk = P1, R1
        for _ in range(dim - 1):
            Pk, Rk = np.kron(P1, Pk), np.kron(R1, Rk)  # last factor = x fastest
        ch = children_of(coarse_mesh)
        # patch -> fine node
        loc = np.stack(np.meshgrid(*[np.arange(M)] * dim, indexing="ij")[::-1], axis=-1).reshape(-1, dim)  # x fastest
        rows, cols, vals = [], [], []
        rrows, rcols, rvals = [], [], []
        fn, cn = fine_mesh.cell_nodes, coarse_mesh.cell_nodes
        child = np.zeros(len(loc), dtype=np.int64)
        lidx = np.zeros(len(loc), dtype=np.int64)
        stride = 1
        for d in range(dim):
            c = (loc[:, d] > p).astype(np.int64)
            child += c << d
            lidx += (loc[:, d] - c * p) * stride
            stride *= n
        patch_nodes = fn[ch[:, child], lidx[None, :]]  # [coarse cell, M^dim]
        nzP = np.abs(Pk) > 0
        for cell in range(coarse_mesh.n_cells):
            r, c = np.nonzero(nzP)
            rows.append(patch_nodes[cell][r])
            cols.append(cn[cell][c])
            vals.append(Pk[r, c])
            r2, c2 = np.nonzero(np.abs(Rk) > 0)
            rrows.append(cn[cell][r2])
            rcols.append(patch_nodes[cell][c2])
            rvals.append(Rk[r2, c2])
        rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
        # shared nodes are reached through several parents with identical values: keep one entry per (row, col)
        key = rows * coarse_mesh.n_nodes + cols
        _, first = np.unique(key, return_index=True)
        Pn = sp.csr_matrix((vals[first], (rows[first], cols[first])), shape=(fine_mesh.n_nodes, coarse_mesh.n_nodes))
        rrows, rcols, rvals = np.concatenate(rrows), np.concatenate(rcols), np.concatenate(rvals)
        key = rrows * fine_mesh.n_nodes + rcols
        _, first = np.unique(key, return_index=True)
        # a coarse node on a child interface can be evaluated in either child: keep the entries of one parent/child only
        Rn = sp.lil_matrix((coarse_mesh.n_nodes, fine_mesh.n_nodes))
        done = np.zeros(coarse_mesh.n_nodes, dtype=bool)
        for cell in range(coarse_mesh.n_cells):
            for i in range(n**dim):
                node = cn[cell][i]
                if done[node]:
                    continue
                done[node] = True
                nz = np.nonzero(np.abs(Rk[i]) > 0)[0]
                for m in nz:
                    Rn[node, patch_nodes[cell][m]] = Rk[i, m]
        eye = sp.identity(ncomp, format="csr")
        self.P = sp.kron(Pn, eye, format="csr")
        self.R = sp.kron(Rn.tocsr(), eye, format="csr")
        fc = sp.diags((~fine_constrained).astype(float))
        cc = sp.diags((~coarse_constrained).astype(float))
        self.Pc = (fc @ self.P @ cc).tocsr()  # level Dirichlet DoFs stay zero

    def prolongate(self, fine_dst, coarse_src):
        fine_dst[:] = self.Pc @ coarse_src

    def prolongate_and_add(self, fine_dst, coarse_src):
        fine_dst += self.Pc @ coarse_src

    def restrict_and_add(self, coarse_dst, fine_src):
        coarse_dst += self.Pc.T @ fine_src

    def interpolate(self, coarse_dst, fine_src):
        coarse_dst[:] = self.R @ fine_src


# ---------------------------------------------------------------------------
# Multigrid V-cycle + PreconditionMG (A.7)
class PreconditionMG:
    def __init__(self, level_ops, smoothers, transfers, dtype=np.float64, coarse_cg=None):
        """level_ops[l](dst, src); smoothers[l] PreconditionChebyshev (l = 0: coarse solver); transfers[l-1]: l-1 <-> l.
        coarse_cg = (max_it, abs_tol, reduction): MGCoarseGridIterativeSolver - SolverCG under a ReductionControl,
        preconditioned with the level-0 smoother (mf_elasticity.h:1183-1198, 'MF Chebyshev coarse = false')."""
        self.A, self.S, self.Tr, self.dtype, self.coarse_cg = level_ops, smoothers, transfers, dtype, coarse_cg
        self.coarse_iterations = []

    def _v(self, level, defect):
        sol = np.zeros_like(defect)
        if level == 0:
            if self.coarse_cg is not None:
                max_it, abs_tol, reduction = self.coarse_cg
                tol = max(abs_tol, reduction * float(np.linalg.norm(defect.astype(np.float64))))  # ReductionControl
                r = solver_cg(self.A[0], sol, defect, self.S[0].vmult, tol, max_it, dtype=self.dtype)
                self.coarse_iterations.append(r.iterations)
                return sol
            self.S[0].vmult(sol, defect)
            return sol
        self.S[level].vmult(sol, defect)
        t = np.empty_like(defect)
        self.A[level](t, sol)
        t[:] = defect - t
        dc = np.zeros(self.Tr[level - 1].Pc.shape[1], dtype=self.dtype)
        self.Tr[level - 1].restrict_and_add(dc, t)
        dc = dc.astype(self.dtype)
        sc = self._v(level - 1, dc)
        sol += (self.Tr[level - 1].Pc @ sc).astype(self.dtype)
        self.S[level].step(sol, defect)
        return sol

    def vmult(self, dst, src):
        top = len(self.A) - 1
        dst[:] = self._v(top, src.astype(self.dtype))
