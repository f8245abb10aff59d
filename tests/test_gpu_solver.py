"""GPU parity of the solver-side device code (csrc/solver.cu) against oracle/solver_oracle.py, through the C ABI:
vector algebra, PreconditionChebyshev (eigenvalue estimate, vmult, step), MGTransferMatrixFree, the V-cycle
(PreconditionMG) and SolverCG.  FP64 level contexts are compared value-wise (the transfer arithmetic is float as in
MGTransferMatrixFree<dim,float>, hence 1e-5); FP32 levels (the reference's LevelNumber, mf_elasticity.h:89-93) are
compared through CG iteration counts, bar +-1 (BASELINE.json north_star)."""
import numpy as np
import pytest

from b200mf import (DeviceVector, LaplaceOperator, MatrixFree, MGTransferMatrixFree, NeoHookOperator, PreconditionChebyshev,
                    PreconditionMG, SolverCG)
from b200mf.mesh import children_of, synthetic_displacement
from common import MAT, c_cells, rel
from oracle import c_oracle as CO
from oracle import solver_oracle as SO
from test_solver_oracle import build_gmg, hierarchy, laplace_level

pytestmark = pytest.mark.gpu
ALL_FACES = ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1))


def test_vector_algebra():
    m = hierarchy(2, (2, 2, 1), 1, 3)[0]
    mf = MatrixFree.from_mesh(m, 3)
    rng = np.random.default_rng(0)
    a, b = rng.uniform(-1, 1, (2, 3 * m.n_nodes))
    da, db = DeviceVector(mf, a), DeviceVector(mf, b)
    assert abs(da.dot(db) - a @ b) < 1e-12 * abs(a @ b)
    assert abs(da.l2_norm() - np.linalg.norm(a)) < 1e-13 * np.linalg.norm(a)
    da.sadd(0.5, -2.0, db)
    assert rel(da.download(), 0.5 * a - 2.0 * b) < 1e-15
    mf32 = MatrixFree.from_mesh(m, 3, dtype="f32")
    d32 = DeviceVector(mf32)
    d32.copy_from(db)
    assert rel(d32.download(), b) < 1e-7


def gpu_laplace_level(m, dtype, dt=0.1):
    mf = MatrixFree.from_mesh(m, 1, dtype=dtype)
    op = LaplaceOperator(mf, dt)
    op.compute_diagonal()
    return mf, op


@pytest.mark.parametrize("p", [1, 2, 4])
def test_chebyshev_laplace(p):
    m = hierarchy(p, (2, 2, 2), 1, 1, mapped=True, clamp=((0, 0), (1, 1)))[0]
    orc, A, inv = laplace_level(m)
    b = np.random.default_rng(p).uniform(-1, 1, m.n_nodes)
    b[m.constrained_dofs(1)] = 0
    mf, op = gpu_laplace_level(m, "f64")
    for kw in (dict(degree=5, smoothing_range=15.0, eig_cg_n_iterations=10),
               dict(degree=SO.INVALID_UNSIGNED_INT, smoothing_range=1e-3, eig_cg_n_iterations=40)):
        ref = SO.PreconditionChebyshev(A, inv, m.constrained_dofs(1), **kw)
        ref.estimate_eigenvalues()
        ch = PreconditionChebyshev(op, **kw)
        info = ch.estimate_eigenvalues()
        assert abs(info["max_eigenvalue_estimate"] - ref.max_eig) < 1e-9 * ref.max_eig
        assert abs(info["min_eigenvalue_estimate"] - ref.min_eig) < 1e-7 * ref.max_eig
        assert info["degree"] == ref.degree and info["cg_iterations"] == ref.cg_iterations
        x_ref = np.zeros(m.n_nodes)
        ref.vmult(x_ref, b)
        s, d = DeviceVector(mf, b), DeviceVector(mf)
        ch.vmult(d, s)
        assert rel(d.download(), x_ref) < 1e-10
        ref.step(x_ref, b)
        ch.step(d, s)
        assert rel(d.download(), x_ref) < 1e-10


@pytest.mark.parametrize("p,ncomp", [(1, 1), (2, 3), (3, 1), (4, 3), (6, 1)])
def test_mg_transfer(p, ncomp):
    mc, mfine = hierarchy(p, (2, 1, 1) if p > 3 else (2, 2, 1), 2, ncomp, mapped=True)
    tr_ref = SO.MGTransfer(mc, mfine, ncomp, mc.constrained_dofs(ncomp), mfine.constrained_dofs(ncomp))
    c_mf, f_mf = MatrixFree.from_mesh(mc, ncomp), MatrixFree.from_mesh(mfine, ncomp)
    tr = MGTransferMatrixFree(c_mf, f_mf, children_of(mc))
    rng = np.random.default_rng(p)
    uc = rng.uniform(-1, 1, ncomp * mc.n_nodes)
    uc[mc.constrained_dofs(ncomp)] = 0
    uf = rng.uniform(-1, 1, ncomp * mfine.n_nodes)
    dc, df = DeviceVector(c_mf, uc), DeviceVector(f_mf, uf)
    out_f = DeviceVector(f_mf)
    tr.prolongate(out_f, dc)
    ref = np.zeros_like(uf)
    tr_ref.prolongate(ref, uc)
    assert rel(out_f.download(), ref) < 1e-6
    base = rng.uniform(-1, 1, uf.size)
    out_f.upload(base)
    tr.prolongate_and_add(out_f, dc)
    assert rel(out_f.download(), base + ref) < 1e-6
    out_c = DeviceVector(c_mf)
    basec = rng.uniform(-1, 1, uc.size)
    out_c.upload(basec)
    tr.restrict_and_add(out_c, df)
    refc = basec.copy()
    tr_ref.restrict_and_add(refc, uf)
    assert rel(out_c.download(), refc) < 1e-6
    tr.interpolate(out_c, df)
    tr_ref.interpolate(refc, uf)
    assert rel(out_c.download(), refc) < 1e-6


def gmg_gpu_laplace(meshes, level_dtype):
    mfs, ops, sms, trs = [], [], [], []
    for l, m in enumerate(meshes):
        mf, op = gpu_laplace_level(m, level_dtype)
        mfs.append(mf)
        ops.append(op)
        if l == 0:
            sms.append(PreconditionChebyshev(op, degree=SO.INVALID_UNSIGNED_INT, smoothing_range=1e-3, eig_cg_n_iterations=m.n_nodes))
        else:
            sms.append(PreconditionChebyshev(op, degree=5, smoothing_range=15.0, eig_cg_n_iterations=10))
            trs.append(MGTransferMatrixFree(mfs[l - 1], mf, children_of(meshes[l - 1])))
    return mfs, ops, sms, trs


@pytest.mark.parametrize("level_dtype", ["f64", "f32"])
def test_gmg_cg_laplace(level_dtype):
    """heat solver: CG + GMG(Chebyshev 5 / range 15 / 10 eig its; level 0 Chebyshev solver 1e-3), heat_transfer.h:757-812"""
    p = 2
    meshes = hierarchy(p, (1, 1, 1), 3, 1, clamp=ALL_FACES)
    ops_ref, As, mg_ref = build_gmg(meshes, 1, laplace_level)
    m = meshes[-1]
    b = np.random.default_rng(0).uniform(-1, 1, m.n_nodes)
    b[m.constrained_dofs(1)] = 0
    tol = 1e-12 * np.linalg.norm(b)  # heat_transfer.h:719-720
    x_ref = np.zeros(m.n_nodes)
    r_ref = SO.solver_cg(As[-1], x_ref, b, mg_ref.vmult, tol, 200)
    mfs, ops, sms, trs = gmg_gpu_laplace(meshes, level_dtype)
    sys_mf = MatrixFree.from_mesh(m, 1)
    sys_op = LaplaceOperator(sys_mf, 0.1)
    mg = PreconditionMG(sms, trs, sys_mf)
    if level_dtype == "f64":  # one V-cycle, value-wise
        z_ref = np.zeros(m.n_nodes)
        mg_ref.vmult(z_ref, b)
        s, d = DeviceVector(sys_mf, b), DeviceVector(sys_mf)
        mg.vmult(d, s)
        assert rel(d.download(), z_ref) < 1e-5
    x, rhs = DeviceVector(sys_mf), DeviceVector(sys_mf, b)
    cg = SolverCG(tol, 200)
    info = cg.solve(sys_op, x, rhs, mg, x_is_zero=True)
    assert info["converged"]
    assert abs(info["iterations"] - r_ref.iterations) <= 1, (info, r_ref.iterations)
    assert rel(x.download(), x_ref) < 1e-8
    gi = mg.graph_info()  # the levels below the finest one ran from the recorded graph, not as single launches
    assert gi["state"] == 1 and gi["top_level"] == len(meshes) - 2 and gi["replays"] >= info["iterations"] - 2, gi


@pytest.mark.parametrize("level_dtype", ["f64", "f32"])
def test_mg_graph_replay_equals_plain_launches(level_dtype):
    """b200mf_mg_set_graph_levels: the V-cycle whose lower levels are replayed from a CUDA graph returns what the plainly
    launched V-cycle returns (first cycle eager, second records, later ones replay), also after the smoothers are
    re-estimated (new Chebyshev constants -> recorded again) and with the recording switched off and on again."""
    p = 2
    meshes = hierarchy(p, (1, 1, 1), 3, 1, clamp=ALL_FACES)
    m = meshes[-1]
    rng = np.random.default_rng(3)
    mfs, ops, sms, trs = gmg_gpu_laplace(meshes, level_dtype)
    sys_mf = MatrixFree.from_mesh(m, 1)
    mg = PreconditionMG(sms, trs, sys_mf)
    plain = PreconditionMG(sms, trs, sys_mf)
    plain.set_graph_levels(-1)
    tol = 1e-12 if level_dtype == "f64" else 2e-5
    s, d, e = DeviceVector(sys_mf), DeviceVector(sys_mf), DeviceVector(sys_mf)
    from b200mf import lib as _lib
    plain.vmult(e, s)  # the smoothers' lazy eigenvalue estimates run in the first application: keep them out of the counts
    for cycle in range(8):
        if cycle == 4:  # new eigenvalue estimates on a changed operator
            for op in ops:
                op.set_delta_t(0.05)
                op.compute_diagonal()
            for sm in sms:
                sm.estimate_eigenvalues()
        if cycle == 6:
            mg.set_graph_levels(-1)
        if cycle == 7:
            mg.set_graph_levels(len(meshes) - 2)
        b = rng.uniform(-1, 1, m.n_nodes)
        b[m.constrained_dofs(1)] = 0
        s.upload(b)
        n0 = _lib.load().b200mf_launch_count()
        mg.vmult(d, s)
        n1 = _lib.load().b200mf_launch_count()
        plain.vmult(e, s)
        n2 = _lib.load().b200mf_launch_count()
        assert n1 - n0 == n2 - n1, (cycle, n1 - n0, n2 - n1)  # replays are accounted with the launches they stand for
        assert rel(d.download(), e.download()) < tol, cycle
    gi = mg.graph_info()
    assert gi["state"] == 1 and gi["records"] >= 3 and gi["replays"] >= 3, gi
    assert plain.graph_info()["records"] == 0


def nh_level(m, u):
    op = CO.NeoHookOperator(c_cells(m), formulation=1, **MAT)
    op.initialize(u, "tensor2")
    op.cache()
    d = op.compute_diagonal()

    def A(dst, src):
        tmp = np.zeros(src.size)
        op.vmult(tmp, np.asarray(src, dtype=np.float64))
        dst[:] = tmp

    return op, A, 1.0 / d


@pytest.mark.parametrize("caching", ["tensor2", "scalar_referential"])
def test_gmg_cg_neo_hooke(caching):
    """solid solver: one linear solve of a Newton step — CG + GMG, Chebyshev(5, range 20, 30 eig its) smoothers on FP32 level
    operators, level-0 Chebyshev solver (mf_elasticity.h:1149-1180, 1981-2070); the level displacement is the fine one
    interpolated down (interpolate_to_mg, :920-922)."""
    p = 2
    meshes = hierarchy(p, (2, 1, 1), 3, 3, mapped=True)
    u_levels = [synthetic_displacement(m.node_coordinates(), 0.01) for m in meshes]  # SPD tangent (0.05 makes it indefinite)
    # oracle
    As, sms_ref, trs_ref = [], [], []
    for l, m in enumerate(meshes):
        _, A, inv = nh_level(m, u_levels[l])
        As.append(A)
        con = m.constrained_dofs(3)
        if l == 0:
            sms_ref.append(SO.PreconditionChebyshev(A, inv, con, degree=SO.INVALID_UNSIGNED_INT, smoothing_range=1e-3, eig_cg_n_iterations=30))
        else:
            sms_ref.append(SO.PreconditionChebyshev(A, inv, con, degree=5, smoothing_range=20.0, eig_cg_n_iterations=30))
            trs_ref.append(SO.MGTransfer(meshes[l - 1], m, 3, meshes[l - 1].constrained_dofs(3), con))
    mg_ref = SO.PreconditionMG(As, sms_ref, trs_ref)
    m = meshes[-1]
    b = np.random.default_rng(0).uniform(-1, 1, 3 * m.n_nodes)
    b[m.constrained_dofs(3)] = 0
    tol = 1e-6 * np.linalg.norm(b)  # tol_lin, mf_elasticity.h:1993
    x_ref = np.zeros(b.size)
    r_ref = SO.solver_cg(As[-1], x_ref, b, mg_ref.vmult, tol, 500)
    # device
    matd = dict(formulation=1, **MAT)
    mfs, ops, sms, trs = [], [], [], []
    for l, mm in enumerate(meshes):
        mf = MatrixFree.from_mesh(mm, 3, dtype="f32")
        op = NeoHookOperator()
        op.set_material(mf, matd)
        op.initialize(u_levels[l], caching)
        op.cache()
        op.compute_diagonal()
        mfs.append(mf)
        ops.append(op)
        if l == 0:
            sms.append(PreconditionChebyshev(op, degree=SO.INVALID_UNSIGNED_INT, smoothing_range=1e-3, eig_cg_n_iterations=30))
        else:
            sms.append(PreconditionChebyshev(op, degree=5, smoothing_range=20.0, eig_cg_n_iterations=30))
            trs.append(MGTransferMatrixFree(mfs[l - 1], mf, children_of(meshes[l - 1])))
    sys_mf = MatrixFree.from_mesh(m, 3)
    sys_op = NeoHookOperator()
    sys_op.set_material(sys_mf, matd)
    sys_op.initialize(u_levels[-1], caching)
    sys_op.cache()
    sys_op.compute_diagonal()
    mg = PreconditionMG(sms, trs, sys_mf)
    x, rhs = DeviceVector(sys_mf), DeviceVector(sys_mf, b)
    info = SolverCG(tol, 500).solve(sys_op, x, rhs, mg, x_is_zero=True)
    assert abs(info["iterations"] - r_ref.iterations) <= 1, (info, r_ref.iterations)
    assert rel(x.download(), x_ref) < 1e-4
    # Jacobi-preconditioned CG (Preconditioner type = jacobi, relaxation 0.65, :2028-2035)
    inv = 1.0 / CO_diag(meshes[-1], u_levels[-1])
    x_j = np.zeros(b.size)
    r_j = SO.solver_cg(As[-1], x_j, b, lambda d, s: d.__setitem__(slice(None), 0.65 * inv * s), tol, 5000)
    x.set_zero()
    info_j = SolverCG(tol, 5000).solve(sys_op, x, rhs, ("jacobi", 0.65), x_is_zero=True)
    assert abs(info_j["iterations"] - r_j.iterations) <= max(1, r_j.iterations // 100), (info_j, r_j.iterations)
    assert rel(x.download(), x_j) < 1e-5
    assert info_j["iterations"] > 2 * info["iterations"]


def CO_diag(m, u):
    op = CO.NeoHookOperator(c_cells(m), formulation=1, **MAT)
    op.initialize(u, "tensor2")
    op.cache()
    return op.compute_diagonal()


def test_cg_reports_no_convergence():
    from b200mf import B200mfError

    m = hierarchy(2, (2, 2, 2), 1, 1, clamp=ALL_FACES)[0]
    mf = MatrixFree.from_mesh(m, 1)
    op = LaplaceOperator(mf, 0.1)
    b = np.random.default_rng(0).uniform(-1, 1, m.n_nodes)
    b[m.constrained_dofs(1)] = 0
    x, rhs = DeviceVector(mf), DeviceVector(mf, b)
    with pytest.raises(B200mfError, match="no convergence"):
        SolverCG(1e-14 * np.linalg.norm(b), 3).solve(op, x, rhs, None, x_is_zero=True)


@pytest.mark.parametrize("level_dtype", ["f64", "f32"])
def test_gmg_cg_with_cg_coarse_solver(level_dtype):
    """'MF Chebyshev coarse = false' (mf_elasticity.h:1183-1198, 1217-1224): MGCoarseGridIterativeSolver - SolverCG under
    ReductionControl(max(m(), 100), 1e-10, 1e-3) on level 0, preconditioned with the level-0 smoother (a degree-5 Chebyshev
    smoother in this mode) - against the same option of the oracle's V-cycle."""
    p = 2
    meshes = hierarchy(p, (2, 2, 1), 3, 1, clamp=ALL_FACES)
    m = meshes[-1]
    # oracle: every level (level 0 included) gets the smoother parameters of the non-Chebyshev-coarse branch
    As, sms_ref, trs_ref = [], [], []
    for l, mm in enumerate(meshes):
        _, A, inv = laplace_level(mm)
        As.append(A)
        con = mm.constrained_dofs(1)
        sms_ref.append(SO.PreconditionChebyshev(A, inv, con, degree=5, smoothing_range=20.0, eig_cg_n_iterations=30))
        if l > 0:
            trs_ref.append(SO.MGTransfer(meshes[l - 1], mm, 1, meshes[l - 1].constrained_dofs(1), con))
    n0 = meshes[0].n_nodes
    mg_ref = SO.PreconditionMG(As, sms_ref, trs_ref, coarse_cg=(max(n0, 100), 1e-10, 1e-3))
    b = np.random.default_rng(1).uniform(-1, 1, m.n_nodes)
    b[m.constrained_dofs(1)] = 0
    tol = 1e-10 * np.linalg.norm(b)
    x_ref = np.zeros(m.n_nodes)
    r_ref = SO.solver_cg(As[-1], x_ref, b, mg_ref.vmult, tol, 200)
    # device
    mfs, ops, sms, trs = [], [], [], []
    for l, mm in enumerate(meshes):
        mf, op = gpu_laplace_level(mm, level_dtype)
        mfs.append(mf)
        ops.append(op)
        sms.append(PreconditionChebyshev(op, degree=5, smoothing_range=20.0, eig_cg_n_iterations=30))
        if l > 0:
            trs.append(MGTransferMatrixFree(mfs[l - 1], mf, children_of(meshes[l - 1])))
    sys_mf = MatrixFree.from_mesh(m, 1)
    sys_op = LaplaceOperator(sys_mf, 0.1)
    mg = PreconditionMG(sms, trs, sys_mf)
    mg.set_coarse_solver_cg(max(n0, 100), 1e-10, 1e-3)
    x, rhs = DeviceVector(sys_mf), DeviceVector(sys_mf, b)
    info = SolverCG(tol, 200).solve(sys_op, x, rhs, mg, x_is_zero=True)
    assert info["converged"]
    assert abs(info["iterations"] - r_ref.iterations) <= 1, (info, r_ref.iterations)
    assert abs(mg.coarse_iterations() - mg_ref.coarse_iterations[-1]) <= 1
    assert rel(x.download(), x_ref) < 1e-7
    # back to the smoother-as-solver coarse grid: still converges
    mg.set_coarse_solver_chebyshev()
    x.set_zero()
    assert SolverCG(tol, 200).solve(sys_op, x, rhs, mg, x_is_zero=True)["converged"]
