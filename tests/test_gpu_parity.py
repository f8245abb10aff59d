"""GPU parity tests (call through the C ABI): CUDA operators vs the CPU oracle
on the same seeded inputs.  Tolerances: FP64 <= 1e-12 relative l2 (BASELINE.json
north_star); FP32 level operators <= 2e-5."""
import numpy as np
import pytest

from b200mf import B200mfError, DeviceVector, LaplaceOperator, MatrixFree, NeoHookOperator
from common import MAT, c_cells, displacement, make_mesh, rel
from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 2e-5
MATD = dict(formulation=1, **MAT)

NCELLS = {1: (5, 4, 3), 2: (4, 3, 3), 3: (3, 3, 2), 4: (3, 2, 2), 5: (2, 2, 2), 6: (2, 2, 1)}


def oracle_nh(m, u, level="tensor2", formulation=1, mu_inclusion=None):
    op = CO.NeoHookOperator(c_cells(m), formulation=formulation, mu_inclusion=mu_inclusion, **MAT)
    op.initialize(u, level)
    op.cache()
    return op


def gpu_nh(m, u, dtype="f64", full_dof_map=False, inclusion=None, explicit_metric=None, level="tensor2", formulation=1):
    mf = MatrixFree.from_mesh(m, 3, dtype=dtype, full_dof_map=full_dof_map, explicit_metric=explicit_metric)
    op = NeoHookOperator()
    mat = dict(MATD, formulation=formulation)
    if inclusion is not None:
        inclusion = dict(inclusion, formulation=formulation)
    op.set_material(mf, mat, inclusion)
    op.initialize(u, level)
    op.cache()
    return mf, op


# (MF caching level, material formulation): every branch of do_operation_on_cell (mf_nh_operator.h:701-1015)
VARIANTS = [("scalar_referential", 1), ("tensor2", 1), ("tensor4", 1), ("tensor4_ns", 1), ("tensor2", 0)]


@pytest.mark.parametrize("level,formulation", VARIANTS)
@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_nh_vmult_all_caching_levels(p, level, formulation):
    """Every cache level / formulation against the oracle's restatement of the same branch, plus the
    reference's own evidence that the four levels are one operator (identical goldens)."""
    m = make_mesh(p, NCELLS[p])
    m.material_id[1::4] = 1
    u = displacement(m)
    incl = dict(MATD, mu=100 * MATD["mu"])
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)
    ref = np.zeros_like(src)
    oracle_nh(m, u, level, formulation, mu_inclusion=incl["mu"]).vmult(ref, src)
    mf, op = gpu_nh(m, u, level=level, formulation=formulation, inclusion=incl)
    s, d = DeviceVector(mf, src), DeviceVector(mf)
    op.vmult(d, s)
    out = d.download()
    assert rel(out, ref) < TOL64
    if formulation == 1 and level != "tensor2":
        ref2 = np.zeros_like(src)
        oracle_nh(m, u, "tensor2", 1, mu_inclusion=incl["mu"]).vmult(ref2, src)
        assert rel(out, ref2) < 1e-11
    if p <= 2:
        op.compute_diagonal()
        assert rel(op.get_matrix_diagonal(), oracle_nh(m, u, level, formulation, mu_inclusion=incl["mu"]).compute_diagonal()) < TOL64


@pytest.mark.parametrize("level,formulation", VARIANTS)
def test_nh_vmult_all_caching_levels_fp32(level, formulation):
    p = 3
    m = make_mesh(p, NCELLS[p])
    u = displacement(m)
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)
    ref = np.zeros_like(src)
    oracle_nh(m, u, level, formulation).vmult(ref, src)
    mf, op = gpu_nh(m, u, dtype="f32", level=level, formulation=formulation)
    out = np.zeros_like(src)
    op.vmult(out, src)
    assert rel(out, ref) < TOL32


def test_formulation0_requires_tensor2():
    m = make_mesh(2, (2, 2, 2))
    mf = MatrixFree.from_mesh(m, 3)
    op = NeoHookOperator()
    op.set_material(mf, dict(MATD, formulation=0))
    with pytest.raises(B200mfError, match="formulation 0 requires tensor2"):
        op.initialize(displacement(m), "tensor4")


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_nh_vmult_fp64(p):
    m = make_mesh(p, NCELLS[p])
    u = displacement(m)
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)
    ref = np.zeros_like(src)
    oracle_nh(m, u).vmult(ref, src)
    mf, op = gpu_nh(m, u)
    out = np.zeros_like(src)
    op.vmult(out, src)  # host-vector entry point
    assert rel(out, ref) < TOL64
    # device-vector entry points: vmult, Tvmult, vmult_add
    s, d = DeviceVector(mf, src), DeviceVector(mf)
    op.vmult(d, s)
    assert rel(d.download(), ref) < TOL64
    op.Tvmult(d, s)
    assert rel(d.download(), ref) < TOL64
    op.vmult_add(d, s)
    assert rel(d.download(), 2 * ref) < TOL64
    assert op.m() == 3 * m.n_nodes
    # smoother hook: the operator pass alone into a zero vector - identity rows left to the consumer
    d.set_zero()
    op.vmult_cells(d, s)
    con = m.constrained_dofs(3)
    want = ref.copy()
    want[con] = 0.0
    assert con.any() and rel(d.download(), want) < TOL64


@pytest.mark.parametrize("p", [2, 3])
def test_nh_cache_tables(p):
    """cache(): deformed metric J_cur = F J_ref and the tensor2 tables vs the oracle's
    (mf_nh_operator.h:440-478)."""
    m = make_mesh(p, NCELLS[p])
    u = displacement(m)
    orc = oracle_nh(m, u)
    mf, op = gpu_nh(m, u)
    import ctypes as C
    nq3 = (p + 1) ** 3
    n = m.n_cells * nq3

    def tab(ptr, k):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n * k,)).reshape(m.n_cells, nq3, k).copy()

    # read the oracle's tables through its struct layout is brittle: recompute from numpy twin instead
    from common import np_cells
    from oracle import np_oracle as NO
    nop = NO.NeoHookOperator(np_cells(m), NO.Material(formulation=1, **MAT))
    nop.initialize(u, "tensor2")
    nop.cache()
    K = np.stack([op.qdata(i) for i in range(9)], -1).reshape(m.n_cells, nq3, 3, 3)
    assert rel(K, nop.invjac_cur) < TOL64
    assert rel(op.qdata(9), nop.JxW_cur) < TOL64
    assert rel(op.qdata(10), nop.cached_scalar) < TOL64
    assert rel(op.qdata(11), nop.cached_second_scalar) < TOL64
    T = nop.cached_tensor2
    for s, (i, j) in enumerate([(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]):
        assert rel(op.qdata(12 + s), T[..., i, j]) < 1e-11


def test_nh_full_dof_map_blocked_numbering():
    """dof_indices path with a non-interleaved caller numbering (component blocks):
    the library renumbers internally and translates at upload/download."""
    p = 2
    m = make_mesh(p, NCELLS[p])
    nn = m.n_nodes
    nl = (p + 1) ** 3
    dof = np.empty((m.n_cells, 3 * nl), dtype=np.int64)
    for c in range(3):
        dof[:, c * nl:(c + 1) * nl] = c * nn + m.cell_nodes
    con_nodes = m.constrained_nodes  # [n_nodes, 3]
    con = np.concatenate([c * nn + np.flatnonzero(con_nodes[:, c]) for c in range(3)]).astype(np.uint32)
    conmask = np.zeros(3 * nn, dtype=bool)
    conmask[con] = True
    u_int = displacement(m)  # interleaved
    u = u_int.reshape(nn, 3).T.reshape(-1).copy()  # blocked
    invjac, jxw = CO.q1_metric(m.vertices, p + 1)
    cells = CO.Cells(p, p + 1, 3, dof, conmask, invjac, jxw)
    orc = CO.NeoHookOperator(cells, formulation=1, **MAT)
    orc.initialize(u, "tensor2")
    orc.cache()
    src = np.random.default_rng(5).uniform(-1, 1, 3 * nn)
    ref = np.zeros_like(src)
    orc.vmult(ref, src)
    mf = MatrixFree(p, 3, 3 * nn, dof_indices=dof, constrained_dofs=con, vertices=m.vertices)
    op = NeoHookOperator()
    op.set_material(mf, MATD)
    op.initialize(u, "tensor2")
    op.cache()
    out = np.zeros_like(src)
    op.vmult(out, src)
    assert rel(out, ref) < TOL64
    op.compute_diagonal()
    d = orc.compute_diagonal()
    assert rel(op.get_matrix_diagonal(), d) < TOL64
    assert abs(op.el(7, 7) - d[7]) < 1e-12 * abs(d[7])


def test_nh_explicit_metric_and_inclusion():
    p = 2
    m = make_mesh(p, NCELLS[p])
    m.material_id[::3] = 2  # inclusion cells (material_id != 0 -> material_inclusion, :656-658)
    u = displacement(m)
    incl = dict(MATD, mu=100 * MATD["mu"])  # mf_elasticity.h:472
    src = np.random.default_rng(7).uniform(-1, 1, 3 * m.n_nodes)
    ref = np.zeros_like(src)
    oracle_nh(m, u, mu_inclusion=incl["mu"]).vmult(ref, src)
    metric = CO.q1_metric(m.vertices, p + 1)
    for em in (None, metric):
        mf, op = gpu_nh(m, u, inclusion=incl, explicit_metric=em)
        out = np.zeros_like(src)
        op.vmult(out, src)
        assert rel(out, ref) < TOL64


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_nh_diagonal_and_jacobi(p):
    m = make_mesh(p, (2, 2, 2))
    u = displacement(m)
    orc = oracle_nh(m, u)
    d = orc.compute_diagonal()
    mf, op = gpu_nh(m, u)
    op.compute_diagonal()
    assert rel(op.get_matrix_diagonal(), d) < TOL64
    assert rel(op.get_matrix_diagonal_inverse(), 1.0 / d) < TOL64
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)
    s, o = DeviceVector(mf, src), DeviceVector(mf)
    op.precondition_Jacobi(o, s, 0.65)
    assert rel(o.download(), 0.65 * src / d) < TOL64


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_nh_vmult_fp32_level_operator(p):
    m = make_mesh(p, NCELLS[p])
    u = displacement(m)
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)
    ref = np.zeros_like(src)
    oracle_nh(m, u).vmult(ref, src)
    mf, op = gpu_nh(m, u, dtype="f32")
    out = np.zeros_like(src)
    op.vmult(out, src)
    assert rel(out, ref) < TOL32


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("with_coef", [False, True])
@pytest.mark.parametrize("mapped", [True, False], ids=["general", "affine"])
def test_laplace_vmult_and_diagonal(p, with_coef, mapped):
    # mapped=False: every cell is affine -> the library's per-cell-metric kernel (LaplaceAffineOp)
    m = make_mesh(p, NCELLS[p], mapped=mapped, ncomp=1, clamp=((0, 0), (1, 1)))
    cells = c_cells(m, 1)
    nq3 = (p + 1) ** 3
    coef = np.random.default_rng(1).uniform(0.5, 2.0, (m.n_cells, nq3)) if with_coef else None
    dt = 0.1
    orc = CO.LaplaceOperator(cells, dt, coef)
    src = np.random.default_rng(p).uniform(-1, 1, m.n_nodes)
    ref = np.zeros_like(src)
    orc.vmult(ref, src)
    mf = MatrixFree.from_mesh(m, 1)
    assert mf.all_cells_affine() == (not mapped)
    op = LaplaceOperator(mf, dt, coef)
    out = np.zeros_like(src)
    op.vmult(out, src)
    assert rel(out, ref) < TOL64
    if p <= 4:
        op.compute_diagonal()
        assert rel(op.get_matrix_diagonal(), orc.compute_diagonal()) < TOL64


@pytest.mark.parametrize("mapped", [True, False], ids=["general", "affine"])
def test_laplace_fp32(mapped):
    p = 4
    m = make_mesh(p, NCELLS[p], mapped=mapped, ncomp=1, clamp=((0, 0),))
    orc = CO.LaplaceOperator(c_cells(m, 1), 0.1)
    src = np.random.default_rng(0).uniform(-1, 1, m.n_nodes)
    ref = np.zeros_like(src)
    orc.vmult(ref, src)
    mf = MatrixFree.from_mesh(m, 1, dtype="f32")
    out = np.zeros_like(src)
    LaplaceOperator(mf, 0.1).vmult(out, src)
    assert rel(out, ref) < TOL32


# enough cells that every persistent block walks several cell groups: exercises the software pipeline
# (TMA staging of the next group's tables, register prefetch of its DoFs) that small meshes never reach
PIPE_CELLS = {1: (40, 40, 36), 2: (20, 20, 20), 3: (24, 20, 16), 4: (12, 12, 10), 5: (10, 8, 8), 6: (8, 8, 8)}


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_nh_vmult_many_groups_per_block(p):
    m = make_mesh(p, PIPE_CELLS[p])
    u = displacement(m)
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)
    ref = np.zeros_like(src)
    level = "scalar_referential" if p % 2 else "tensor2"
    oracle_nh(m, u, level).vmult(ref, src)
    mf, op = gpu_nh(m, u, level=level)
    s, d = DeviceVector(mf, src), DeviceVector(mf)
    for _ in range(2):
        op.vmult(d, s)
        assert rel(d.download(), ref) < TOL64


def test_error_behaviour():
    m = make_mesh(2, (2, 2, 2))
    mf = MatrixFree.from_mesh(m, 3)
    op = NeoHookOperator()
    op.set_material(mf, MATD)
    s, d = DeviceVector(mf), DeviceVector(mf)
    op.initialize(displacement(m), "tensor2")
    with pytest.raises(B200mfError, match="before cache"):
        op.vmult(d, s)
    # det F <= 0 is reported (debug assertion mf_nh_operator.h:430-435)
    op.initialize(-40.0 * displacement(m), "tensor2")
    with pytest.raises(B200mfError, match="det_F is not positive"):
        op.cache()
    with pytest.raises(B200mfError):
        MatrixFree(7, 3, 3 * m.n_nodes, node_indices=m.cell_nodes, vertices=m.vertices)


def test_symmetry_large_random():
    """size-independent property at a larger size: v^T A u == u^T A v."""
    p = 3
    m = make_mesh(p, (12, 10, 9), permute=None)
    u = displacement(m)
    mf, op = gpu_nh(m, u)
    rng = np.random.default_rng(11)
    a, b = rng.uniform(-1, 1, (2, 3 * m.n_nodes))
    Aa, Ab = np.zeros_like(a), np.zeros_like(a)
    op.vmult(Aa, a)
    op.vmult(Ab, b)
    assert abs(b @ Aa - a @ Ab) < 1e-11 * abs(b @ Aa)
    con = m.constrained_dofs(3)
    assert np.array_equal(Aa[con], a[con])


# ---------------------------------------------------------------------------
# cell part of assemble_residual (mf_elasticity.h:1764-1846) against the dense numpy oracle
@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("formulation", [1, 0])
def test_nh_residual_cells(p, formulation):
    from b200mf import DeviceVector
    from common import np_cells
    from oracle import np_oracle as NO

    m = make_mesh(p, NCELLS[p])
    u = displacement(m)
    rng = np.random.default_rng(100 + p)
    acc = rng.uniform(-1, 1, 3 * m.n_nodes) * 50.0
    bf = (0.0, -0.5, 0.25)
    mat = NO.Material(formulation=formulation, **MAT)
    ref = NO.nh_residual(np_cells(m), mat, u, acc, bf)
    mf = MatrixFree.from_mesh(m, 3)
    op = NeoHookOperator()
    op.set_material(mf, dict(formulation=formulation, **MAT))
    rhs = DeviceVector(mf)
    op.assemble_residual(rhs, DeviceVector(mf, u), DeviceVector(mf, acc), bf)
    out = rhs.download()
    assert rel(out, ref) < TOL64
    # constrained rows are not written (distribute_local_to_global), and set_zero keeps them zero
    assert np.all(out[m.constrained_dofs(3)] == 0.0)
    rhs.set_zero_constrained()
    assert rel(rhs.download(), ref) < TOL64


# "Fused solver vector work": p . Ap accumulated inside the operator pass (b200mf_*_vmult_dot)
@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_vmult_dot_fused(p, dtype):
    from b200mf import DeviceVector

    m = make_mesh(p, NCELLS[p])
    u = displacement(m)
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)  # non-zero on constrained rows: identity rows count too
    mf, op = gpu_nh(m, u, dtype=dtype)
    s, d = DeviceVector(mf, src), DeviceVector(mf)
    op.vmult(d, s)
    ref_vec = d.download()
    ref = float(np.dot(s.download(), ref_vec))
    val = op.vmult_dot(d, s)
    assert np.array_equal(d.download(), ref_vec) or rel(d.download(), ref_vec) < (1e-14 if dtype == "f64" else 1e-6)
    assert abs(val - ref) <= (1e-12 if dtype == "f64" else 2e-5) * abs(ref)
    # scalar operator, affine and general cells
    for mapped in (True, False):
        ms = make_mesh(p, NCELLS[p], mapped=mapped, ncomp=1, clamp=((0, 0),))
        mfs = MatrixFree.from_mesh(ms, 1, dtype=dtype)
        lop = LaplaceOperator(mfs, 0.1)
        x = np.random.default_rng(3).uniform(-1, 1, ms.n_nodes)
        xs, ys = DeviceVector(mfs, x), DeviceVector(mfs)
        v = lop.vmult_dot(ys, xs)
        r = float(np.dot(xs.download(), ys.download()))
        assert abs(v - r) <= (1e-12 if dtype == "f64" else 2e-5) * abs(r)


# cell loop of LaplaceProblem::assemble_rhs (heat_transfer.h:599-615)
@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("mapped", [True, False], ids=["general", "affine"])
def test_heat_rhs_cells(p, mapped):
    from b200mf import DeviceVector
    from common import np_cells
    from oracle import np_oracle as NO

    m = make_mesh(p, NCELLS[p], mapped=mapped, ncomp=1, clamp=None)  # the inhomogeneous operator carries no constraints
    cd = np_cells(m, 1)
    rng = np.random.default_rng(7 + p)
    u_old = rng.uniform(-1, 1, m.n_nodes)
    f_q = rng.uniform(-2, 2, (m.n_cells, (p + 1) ** 3))
    dt = 0.1
    mf = MatrixFree.from_mesh(m, 1)
    op = LaplaceOperator(mf, dt)
    for src_q, const in ((f_q, 0.0), (None, -1.5)):
        ref = NO.heat_rhs_cells(cd, dt, u_old, src_q, const)
        rhs = DeviceVector(mf)
        op.assemble_rhs(rhs, DeviceVector(mf, u_old), src_q, const)
        assert rel(rhs.download(), ref) < TOL64


@pytest.mark.parametrize("p,ncells,numbering", [(2, (16, 16, 16), "lexicographic"), (2, (16, 16, 16), "random"),
                                                (2, (16, 16, 16), "morton"), (4, (10, 10, 10), "lexicographic")])
def test_nh_vmult_host_pipelined(p, ncells, numbering):
    """The host-vector entry point on a mesh large enough for the pipelined schedule (sliced H2D || scatter + cell waves ||
    sliced D2H, b200mf_nh_vmult_host): same result as the device-vector vmult and as the oracle, for a locality-preserving
    numbering (frontiers advance with the waves), a cell-wise Morton numbering and a random one (degenerate schedule)."""
    from b200mf.mesh import StructuredMesh

    m = StructuredMesh(ncells, p, permute_seed=3 if numbering == "random" else None)
    m.clamp([(0, 0), (1, 1)], 3)
    if numbering == "morton":
        m.renumber_cellwise_morton()
    u = displacement(m)
    src = np.random.default_rng(p).uniform(-1, 1, 3 * m.n_nodes)  # constrained entries non-zero: dst[c] = src[c] there
    mf, op = gpu_nh(m, u)
    out = np.zeros_like(src)
    op.vmult(out, src)  # host entry point
    s, d = DeviceVector(mf, src), DeviceVector(mf)
    op.vmult(d, s)
    assert rel(out, d.download()) < 1e-14
    out2 = np.full_like(src, np.nan)
    op.vmult(out2, src)  # staging buffers are reused
    assert rel(out2, out) < 1e-14
    ref = np.zeros_like(src)
    oracle_nh(m, u).vmult(ref, src)
    assert rel(out, ref) < TOL64


@pytest.mark.parametrize("p", [5, 6])
@pytest.mark.parametrize("level,formulation", [("tensor2", 1), ("scalar_referential", 1), ("tensor2", 0)])
def test_nh_diagonal_high_degree(p, level, formulation):
    """compute_diagonal at Q5/Q6 (sliced tables: the diagonal pass streams the q-planes once per component)."""
    m = make_mesh(p, (2, 1, 1))
    u = displacement(m)
    d = oracle_nh(m, u, level, formulation).compute_diagonal()
    mf, op = gpu_nh(m, u, level=level, formulation=formulation)
    op.compute_diagonal()
    assert rel(op.get_matrix_diagonal(), d) < TOL64
    assert rel(op.get_matrix_diagonal_inverse(), 1.0 / d) < TOL64


@pytest.mark.parametrize("p", [2, 4, 5])
@pytest.mark.parametrize("level,formulation", VARIANTS)
def test_nh_cache_tables_all_variants(p, level, formulation):
    """cache() tables of every MF caching level against the numpy twin (mf_nh_operator.h:392-523), both table layouts
    (plain for p <= 3, sliced for p >= 4): cached_scalar / cached_second_scalar / cached_tensor2 / cached_tensor4 / cached_tensor4_ns
    and, for the deformed-frame levels, the metric J_cur = F J_ref."""
    from common import np_cells
    from oracle import np_oracle as NO

    m = make_mesh(p, NCELLS[p])
    u = displacement(m)
    nq3 = (p + 1) ** 3
    nop = NO.NeoHookOperator(np_cells(m), NO.Material(formulation=formulation, **MAT))
    nop.initialize(u, level)
    nop.cache()
    mf, op = gpu_nh(m, u, level=level, formulation=formulation)
    q = lambda i: op.qdata(i)  # noqa: E731
    sym = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
    if formulation == 0:  # K (9), JxW, J^(-1/3), F (9)
        K = np.stack([q(i) for i in range(9)], -1).reshape(m.n_cells, nq3, 3, 3)
        assert rel(K, nop.invjac_cur) < TOL64 and rel(q(9), nop.JxW_cur) < TOL64
        assert rel(q(10), nop.cached_scalar) < TOL64
        F = np.stack([q(11 + i) for i in range(9)], -1).reshape(m.n_cells, nq3, 3, 3)
        assert rel(F, nop.cached_tensor2) < TOL64
    elif level == "tensor2":
        assert rel(q(10), nop.cached_scalar) < TOL64 and rel(q(11), nop.cached_second_scalar) < TOL64
        for s, (i, j) in enumerate(sym):
            assert rel(q(12 + s), nop.cached_tensor2[..., i, j]) < 1e-11
    elif level == "tensor4":  # K (9), JxW, c2, T (6), then the 6x6 table
        assert rel(q(10), nop.cached_second_scalar) < TOL64
        for s, (i, j) in enumerate(sym):
            assert rel(q(11 + s), nop.cached_tensor2[..., i, j]) < 1e-11
        for I, (i, j) in enumerate(sym):
            for J, (k, l) in enumerate(sym):
                assert np.abs(q(17 + I * 6 + J) - nop.cached_tensor4[..., i, j, k, l]).max() <= 1e-11 * np.abs(nop.cached_tensor4).max()
    elif level == "tensor4_ns":  # reference metric is not a cached table: the 81 entries only
        assert op_n_fields(op) == 81
        A = np.stack([q(i) for i in range(81)], -1).reshape(m.n_cells, nq3, 3, 3, 3, 3)
        assert np.abs(A - nop.cached_tensor4_ns).max() <= 1e-11 * np.abs(nop.cached_tensor4_ns).max()
    else:  # scalar_referential: c = mu - 2 lambda ln J
        assert op_n_fields(op) == 1
        assert rel(q(0), nop.cached_scalar) < TOL64


def op_n_fields(op):
    from b200mf import lib as L_

    return L_.load().b200mf_nh_n_qfields(op.h)
