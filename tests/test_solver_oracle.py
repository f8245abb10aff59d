"""CPU pins of the solver oracle (oracle/solver_oracle.py): properties that hold independently of deal.II —
exactness of the level transfer on polynomials, restriction = prolongation^T, Chebyshev eigenvalue bounds,
CG/GMG convergence — so that the GPU parity tests compare against a checked restatement."""
import numpy as np
import pytest

from b200mf.mesh import StructuredMesh, children_of, smooth_mapping
from common import MAT, c_cells, displacement
from oracle import c_oracle as CO
from oracle import solver_oracle as SO


def hierarchy(p, coarse_cells, n_levels, ncomp, mapped=False, clamp=((0, 0),)):
    meshes = []
    for l in range(n_levels):
        nc = tuple(c * 2**l for c in coarse_cells)
        m = StructuredMesh(nc, p, mapping=smooth_mapping(0.04) if mapped else None)
        m.clamp(list(clamp), ncomp)
        meshes.append(m)
    return meshes


@pytest.mark.parametrize("p", [1, 2, 3])
def test_transfer_polynomial_exact_and_transpose(p):
    mc, mf = hierarchy(p, (2, 1, 2), 2, 1, clamp=())
    none_c, none_f = np.zeros(mc.n_nodes, bool), np.zeros(mf.n_nodes, bool)
    tr = SO.MGTransfer(mc, mf, 1, none_c, none_f)
    assert np.array_equal(SO.children_of(mc), children_of(mc))

    def poly(x):
        return 1.0 + x[:, 0] ** p - 2.0 * x[:, 1] ** min(p, 2) * x[:, 2] + 0.5 * x[:, 2] ** p

    uc, uf = poly(mc.node_coordinates()), poly(mf.node_coordinates())
    out = np.zeros(mf.n_nodes)
    tr.prolongate(out, uc)
    assert np.abs(out - uf).max() < 1e-12
    back = np.zeros(mc.n_nodes)
    tr.interpolate(back, uf)
    assert np.abs(back - uc).max() < 1e-12
    # restriction is the transpose of prolongation
    a, b = np.random.default_rng(0).uniform(-1, 1, mf.n_nodes), np.random.default_rng(1).uniform(-1, 1, mc.n_nodes)
    r = np.zeros(mc.n_nodes)
    tr.restrict_and_add(r, a)
    pb = np.zeros(mf.n_nodes)
    tr.prolongate(pb, b)
    assert abs(r @ b - a @ pb) < 1e-12 * abs(a @ pb)


def laplace_level(m, dt=0.1):
    op = CO.LaplaceOperator(c_cells(m, 1), dt)
    d = op.compute_diagonal()

    def A(dst, src):
        tmp = np.zeros(src.size)
        op.vmult(tmp, np.asarray(src, dtype=np.float64))
        dst[:] = tmp

    return op, A, 1.0 / d


def test_chebyshev_eigenvalue_estimate_bounds_spectrum():
    m = hierarchy(2, (2, 2, 2), 1, 1, clamp=((0, 0), (1, 1)))[0]
    op, A, inv = laplace_level(m)
    n = m.n_nodes
    dense = np.zeros((n, n))
    e = np.zeros(n)
    for i in range(n):
        e[:] = 0
        e[i] = 1
        col = np.zeros(n)
        A(col, e)
        dense[:, i] = col
    ev = np.linalg.eigvals(np.diag(inv) @ dense).real
    ch = SO.PreconditionChebyshev(A, inv, m.constrained_dofs(1), degree=5, smoothing_range=15.0, eig_cg_n_iterations=10)
    ch.estimate_eigenvalues()
    assert 0.9 * ev.max() < ch.max_eig < 1.35 * ev.max()
    b = np.random.default_rng(0).uniform(-1, 1, n)
    b[m.constrained_dofs(1)] = 0
    # solver mode on a small problem reduces the residual by the requested factor
    sol = SO.PreconditionChebyshev(A, inv, m.constrained_dofs(1), degree=SO.INVALID_UNSIGNED_INT, smoothing_range=1e-3,
                                   eig_cg_n_iterations=n)
    x = np.zeros(n)
    sol.vmult(x, b)
    r = np.zeros(n)
    A(r, x)
    assert np.linalg.norm(b - r) < 2e-3 * np.linalg.norm(b)


def build_gmg(meshes, ncomp, level_builder):
    ops, As, sms, trs = [], [], [], []
    for l, m in enumerate(meshes):
        op, A, inv = level_builder(m)
        ops.append(op)
        As.append(A)
        con = m.constrained_dofs(ncomp)
        if l == 0:
            sms.append(SO.PreconditionChebyshev(A, inv, con, degree=SO.INVALID_UNSIGNED_INT, smoothing_range=1e-3,
                                                eig_cg_n_iterations=inv.size))
        else:
            sms.append(SO.PreconditionChebyshev(A, inv, con, degree=5, smoothing_range=15.0, eig_cg_n_iterations=10))
            trs.append(SO.MGTransfer(meshes[l - 1], m, ncomp, meshes[l - 1].constrained_dofs(ncomp), con))
    return ops, As, SO.PreconditionMG(As, sms, trs)


def test_gmg_cg_laplace_converges_fast():
    meshes = hierarchy(2, (1, 1, 1), 3, 1, clamp=((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)))
    ops, As, mg = build_gmg(meshes, 1, laplace_level)
    m = meshes[-1]
    b = np.random.default_rng(0).uniform(-1, 1, m.n_nodes)
    b[m.constrained_dofs(1)] = 0
    x = np.zeros(m.n_nodes)
    r = SO.solver_cg(As[-1], x, b, mg.vmult, 1e-10 * np.linalg.norm(b), 100)
    assert r.converged and r.iterations <= 12
    x2 = np.zeros(m.n_nodes)
    inv = 1.0 / ops[-1].compute_diagonal()
    r2 = SO.solver_cg(As[-1], x2, b, lambda d, s: d.__setitem__(slice(None), inv * s), 1e-10 * np.linalg.norm(b), 2000)
    assert r2.converged and r2.iterations > 2 * r.iterations
    assert np.linalg.norm(x - x2) < 1e-7 * np.linalg.norm(x2)
