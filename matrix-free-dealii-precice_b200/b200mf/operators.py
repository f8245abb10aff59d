"""Host-side mirror of the reference's operator interface on top of the C ABI.

Names follow the reference so that the parity tests read like its call sites:
  MatrixFree          <- dealii::MatrixFree<3,Number>            (mf_elasticity.h:933-992)
  DeviceVector        <- LinearAlgebra::distributed::Vector       (mf_elasticity.h:693-706)
  NeoHookOperator     <- include/solid_mechanics/mf_nh_operator.h:99-237
  LaplaceOperator     <- include/heat_transfer/heat_transfer.h:81-127
All arithmetic runs in libb200mf.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _l


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Communicator:
    """One rank per GPU (the role of MPI_COMM_WORLD, src/solid.cc:30-31): an NCCL communicator owned by the library.
    The 128-byte unique id is created on rank 0 and handed to the other ranks by the host (here: torch.distributed)."""

    def __init__(self, rank, world, unique_id: bytes):
        self.rank, self.world = int(rank), int(world)
        buf = C.create_string_buffer(bytes(unique_id), 128)
        h = C.c_void_p()
        _l.check(_l.load().b200mf_comm_create(C.byref(h), self.rank, self.world, buf))
        self.h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _l.check(_l.load().b200mf_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls):
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(rank, world, box[0])

    def destroy(self):
        if getattr(self, "h", None):
            _l.load().b200mf_comm_destroy(self.h)
            self.h = None


class MatrixFree:
    """Cell-batch data resident in HBM (index maps, constraint masks, metric)."""

    def __init__(self, degree, n_components, n_dofs, *, dof_indices=None, node_indices=None, constrained_dofs=None,
                 vertices=None, inverse_jacobian=None, jxw=None, material_id=None, dtype="f64", n_owned=None,
                 n_q_points_1d=None):
        L = _l.load()
        self.degree, self.n_components, self.n_dofs = int(degree), int(n_components), int(n_dofs)
        self.dtype = dtype
        self.np_dtype = np.float64 if dtype == "f64" else np.float32
        keep = []

        def arr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a

        d = _l.CtxDesc()
        d.degree = self.degree
        d.n_q_points_1d = self.degree + 1 if n_q_points_1d is None else int(n_q_points_1d)
        d.n_components = self.n_components
        d.dtype = _l.F64 if dtype == "f64" else _l.F32
        idx = dof_indices if dof_indices is not None else node_indices
        d.n_cells = int(np.asarray(idx).shape[0])
        d.n_dofs = self.n_dofs
        d.n_owned = self.n_dofs if n_owned is None else int(n_owned)
        d.dof_indices = _ptr(arr(dof_indices, np.uint32))
        d.node_indices = _ptr(arr(node_indices, np.uint32))
        con = arr(constrained_dofs, np.uint32)
        d.constrained_dofs = _ptr(con)
        d.n_constrained = 0 if con is None else con.size
        d.vertices = _ptr(arr(vertices, np.float64))
        d.inverse_jacobian = _ptr(arr(inverse_jacobian, np.float64))
        d.jxw = _ptr(arr(jxw, np.float64))
        d.material_id = _ptr(arr(material_id, np.uint8))
        self.n_cells = d.n_cells
        h = C.c_void_p()
        _l.check(L.b200mf_ctx_create(C.byref(h), C.byref(d)))
        self.h = h

    @classmethod
    def from_mesh(cls, mesh, n_components, dtype="f64", full_dof_map=False, explicit_metric=None):
        """Build from a b200mf.mesh.StructuredMesh (synthetic benchmarks / tests)."""
        con = np.flatnonzero(mesh.constrained_dofs(n_components)).astype(np.uint32)
        kw = dict(constrained_dofs=con, material_id=mesh.material_id, dtype=dtype)
        if explicit_metric is not None:
            kw.update(inverse_jacobian=explicit_metric[0], jxw=explicit_metric[1])
        else:
            kw.update(vertices=mesh.vertices)
        if full_dof_map:
            kw.update(dof_indices=mesh.dof_index(n_components))
        else:
            kw.update(node_indices=mesh.cell_nodes)
        return cls(mesh.p, n_components, mesh.n_nodes * n_components, **kw)

    @classmethod
    def from_partition(cls, part, comm: "Communicator | None", dtype="f64"):
        """Build rank-local data from a b200mf.partition.BrickPartition and attach the vector partitioner
        (import indices / ghost ranges per neighbour)."""
        con = np.flatnonzero(part.constrained_dofs()).astype(np.uint32)
        mf = cls(part.p, part.ncomp, part.n_dofs, node_indices=part.cell_nodes, constrained_dofs=con,
                 vertices=part.vertices, material_id=part.material_id, dtype=dtype, n_owned=part.n_owned)
        if comm is not None and comm.world > 1:
            mf.set_partitioner(comm, part.partitioner_arrays())
        return mf

    def set_partitioner(self, comm, arrays):
        ranks = np.ascontiguousarray(arrays["ranks"], dtype=np.int32)
        so = np.ascontiguousarray(arrays["send_offsets"], dtype=np.int64)
        si = np.ascontiguousarray(arrays["send_indices"], dtype=np.uint32)
        go = np.ascontiguousarray(arrays["ghost_offsets"], dtype=np.int64)
        gc = np.ascontiguousarray(arrays["ghost_counts"], dtype=np.int64)
        d = _l.PartitionerDesc(len(ranks), _ptr(ranks), _ptr(so), _ptr(si), _ptr(go), _ptr(gc))
        _l.check(_l.load().b200mf_ctx_set_partitioner(self.h, comm.h, C.byref(d)))
        self.comm = comm

    def exchange_is_p2p(self):
        """True: ghost exchanges run peer to peer over NVLink (CUDA IPC); False: ncclSend/ncclRecv."""
        return bool(_l.load().b200mf_ctx_exchange_is_p2p(self.h))

    def set_stream(self, cuda_stream):
        _l.check(_l.load().b200mf_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        _l.check(_l.load().b200mf_ctx_synchronize(self.h))

    def set_overlap(self, on=True):
        """Hide the ghost exchanges of vmult behind the interior cells (default on multi-rank contexts)."""
        _l.check(_l.load().b200mf_ctx_set_overlap(self.h, int(on)))

    def all_cells_affine(self):
        return _l.load().b200mf_ctx_cell_type(self.h) == 1

    def time_kernels(self, on=True):
        """Bracket every cell-loop launch with CUDA events on the launching stream (bench.py roofline)."""
        _l.check(_l.load().b200mf_ctx_time_kernels(self.h, int(on)))

    def kernel_time(self):
        """(summed ms, launches) of the cell-loop kernels since time_kernels(True)."""
        ms, n = C.c_double(0), C.c_int64(0)
        _l.check(_l.load().b200mf_ctx_kernel_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def memory_consumption(self):
        return _l.load().b200mf_ctx_memory_consumption(self.h)

    def initialize_dof_vector(self):
        return DeviceVector(self)

    def __del__(self):
        if getattr(self, "h", None):
            try:
                _l.load().b200mf_ctx_destroy(self.h)
            except TypeError:  # interpreter shutdown: the module globals are already gone
                pass
            self.h = None


class DeviceVector:
    """Device-resident vector in the library's internal numbering."""

    def __init__(self, mf: MatrixFree, host=None):
        self.mf = mf
        p = C.c_void_p()
        _l.check(_l.load().b200mf_vec_create(mf.h, C.byref(p)))
        self.ptr = p
        if host is not None:
            self.upload(host)

    def upload(self, host):
        h = np.ascontiguousarray(host, dtype=np.float64)
        assert h.size == self.mf.n_dofs
        _l.check(_l.load().b200mf_vec_upload_f64(self.mf.h, self.ptr, _ptr(h)))
        self.mf.synchronize()
        return self

    def download(self):
        out = np.empty(self.mf.n_dofs, dtype=np.float64)
        _l.check(_l.load().b200mf_vec_download_f64(self.mf.h, _ptr(out), self.ptr))
        return out

    def set_zero_constrained(self):
        """AffineConstraints::set_zero (mf_elasticity.h:1905, 2059)."""
        _l.check(_l.load().b200mf_vec_set_zero_constrained(self.mf.h, self.ptr))

    def update_ghost_values(self):
        _l.check(_l.load().b200mf_vec_update_ghost_values(self.mf.h, self.ptr))

    # LinearAlgebra::distributed::Vector algebra (owned entries, summed over ranks)
    def dot(self, other):
        v = C.c_double()
        _l.check(_l.load().b200mf_vec_dot(self.mf.h, self.ptr, other.ptr, C.byref(v)))
        return v.value

    def l2_norm(self):
        v = C.c_double()
        _l.check(_l.load().b200mf_vec_l2_norm(self.mf.h, self.ptr, C.byref(v)))
        return v.value

    def sadd(self, s, a, x):
        """self = s*self + a*x"""
        _l.check(_l.load().b200mf_vec_sadd(self.mf.h, self.ptr, float(s), float(a), x.ptr))

    def set_zero(self):
        _l.check(_l.load().b200mf_vec_set_zero(self.mf.h, self.ptr))

    def copy_from(self, other):
        """same mesh, possibly another number type (double system <-> float level)"""
        _l.check(_l.load().b200mf_vec_convert(self.mf.h, self.ptr, other.mf.h, other.ptr))

    def compress_add(self):
        _l.check(_l.load().b200mf_vec_compress_add(self.mf.h, self.ptr))

    def zero_out_ghost_values(self):
        _l.check(_l.load().b200mf_vec_zero_out_ghost_values(self.mf.h, self.ptr))

    def __del__(self):
        if getattr(self, "ptr", None) and getattr(self.mf, "h", None):
            try:
                _l.load().b200mf_vec_destroy(self.mf.h, self.ptr)
            except TypeError:  # interpreter shutdown: the module globals are already gone
                pass
            self.ptr = None


def _dv(x):
    return x.ptr if isinstance(x, DeviceVector) else C.c_void_p(int(x))


class NeoHookOperator:
    """mf_nh_operator.h:99-237.  `initialize(displacement, caching)` keeps a
    reference to the displacement vector like the original (:339-341)."""

    def __init__(self):
        self.h = None
        self.mf = None
        self.displacement = None

    def set_material(self, mf: MatrixFree, material: dict, material_inclusion: dict | None = None):
        """mf_nh_operator.h:527-539; material = dict(mu, nu, rho, alpha_1, formulation)."""
        self.clear()
        self.mf = mf
        m0 = _l.Material(material["mu"], material["nu"], material["rho"], material["alpha_1"],
                         int(material.get("formulation", 1)))
        mi = material_inclusion or material
        m1 = _l.Material(mi["mu"], mi["nu"], mi["rho"], mi["alpha_1"], int(mi.get("formulation", 1)))
        h = C.c_void_p()
        _l.check(_l.load().b200mf_nh_create(C.byref(h), mf.h, C.byref(m0), C.byref(m1)))
        self.h = h

    def initialize(self, displacement, caching: str):
        if caching not in _l.CACHING:
            raise ValueError("Unknown caching")  # mf_nh_operator.h:385
        self.displacement = displacement
        _l.check(_l.load().b200mf_nh_initialize(self.h, _l.CACHING[caching]))

    def cache(self):
        bad = C.c_int64(0)
        L = _l.load()
        if isinstance(self.displacement, DeviceVector):
            _l.check(L.b200mf_nh_cache(self.h, self.displacement.ptr, C.byref(bad)))
        else:
            d = np.ascontiguousarray(self.displacement, dtype=np.float64)
            _l.check(L.b200mf_nh_cache_host(self.h, _ptr(d), C.byref(bad)))

    def assemble_residual(self, rhs, total_displacement, acceleration, body_force=(0.0, 0.0, 0.0)):
        """Cell part of Solid::assemble_residual, fast path (mf_elasticity.h:1764-1846): rhs += R_cells(u, a) on
        DeviceVectors; the caller zeroes rhs, adds interface faces, then compress_add() and set_zero_constrained()."""
        bf = (C.c_double * 3)(*[float(x) for x in body_force])
        bad = C.c_int64(0)
        _l.check(_l.load().b200mf_nh_assemble_residual(self.h, _dv(rhs), _dv(total_displacement), _dv(acceleration), bf, C.byref(bad)))

    def vmult(self, dst, src):
        L = _l.load()
        if isinstance(dst, np.ndarray):
            s = np.ascontiguousarray(src, dtype=np.float64)
            assert dst.dtype == np.float64 and dst.flags.c_contiguous
            _l.check(L.b200mf_nh_vmult_host(self.h, _ptr(dst), _ptr(s)))
        else:
            _l.check(L.b200mf_nh_vmult(self.h, _dv(dst), _dv(src)))

    def vmult_dot(self, dst, src):
        """dst = A src and return src . (A src), accumulated inside the operator pass (fused CG dot product)."""
        out = C.c_double(0)
        _l.check(_l.load().b200mf_nh_vmult_dot_host(self.h, _dv(dst), _dv(src), C.byref(out)))
        return out.value

    def vmult_add(self, dst, src):
        _l.check(_l.load().b200mf_nh_vmult_add(self.h, _dv(dst), _dv(src)))

    def vmult_cells(self, dst_zero, src):
        """smoother hook: the operator pass alone (dst zero on entry, constrained rows left 0)"""
        _l.check(_l.load().b200mf_nh_vmult_cells(self.h, _dv(dst_zero), _dv(src)))

    def Tvmult(self, dst, src):
        if isinstance(dst, np.ndarray):
            return self.vmult(dst, src)
        _l.check(_l.load().b200mf_nh_Tvmult(self.h, _dv(dst), _dv(src)))

    def Tvmult_add(self, dst, src):
        _l.check(_l.load().b200mf_nh_Tvmult_add(self.h, _dv(dst), _dv(src)))

    def compute_diagonal(self):
        _l.check(_l.load().b200mf_nh_compute_diagonal(self.h))

    def get_matrix_diagonal(self):
        d = np.empty(self.mf.n_dofs)
        _l.check(_l.load().b200mf_nh_diagonal_host(self.h, _ptr(d), None))
        return d

    def get_matrix_diagonal_inverse(self):
        d = np.empty(self.mf.n_dofs)
        _l.check(_l.load().b200mf_nh_diagonal_host(self.h, None, _ptr(d)))
        return d

    def precondition_Jacobi(self, dst, src, omega):
        _l.check(_l.load().b200mf_nh_precondition_jacobi(self.h, _dv(dst), _dv(src), float(omega)))

    def el(self, row, col):
        assert row == col  # mf_nh_operator.h:1067
        v = C.c_double()
        _l.check(_l.load().b200mf_nh_el(self.h, int(row), C.byref(v)))
        return v.value

    def m(self):
        return _l.load().b200mf_nh_m(self.h)

    n = m

    def memory_consumption(self):
        return _l.load().b200mf_nh_memory_consumption(self.h)

    def qdata(self, field):
        n = self.mf.degree + 1
        out = np.empty((self.mf.n_cells, n**3))
        _l.check(_l.load().b200mf_nh_qdata_host(self.h, int(field), _ptr(out)))
        return out

    def clear(self):
        if getattr(self, "h", None):
            _l.load().b200mf_nh_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass


class LaplaceOperator:
    """heat_transfer.h:81-127 (+ MatrixFreeOperators::Base surface)."""

    def __init__(self, mf: MatrixFree, delta_t, coefficient=None):
        self.mf = mf
        coef = None if coefficient is None else np.ascontiguousarray(coefficient, dtype=np.float64)
        h = C.c_void_p()
        _l.check(_l.load().b200mf_laplace_create(C.byref(h), mf.h, float(delta_t), _ptr(coef)))
        self.h = h

    def set_delta_t(self, dt):
        _l.check(_l.load().b200mf_laplace_set_delta_t(self.h, float(dt)))

    def vmult(self, dst, src):
        L = _l.load()
        if isinstance(dst, np.ndarray):
            s = np.ascontiguousarray(src, dtype=np.float64)
            _l.check(L.b200mf_laplace_vmult_host(self.h, _ptr(dst), _ptr(s)))
        else:
            _l.check(L.b200mf_laplace_vmult(self.h, _dv(dst), _dv(src)))

    Tvmult = vmult

    def vmult_dot(self, dst, src):
        """dst = A src and return src . (A src), accumulated inside the operator pass (fused CG dot product)."""
        out = C.c_double(0)
        _l.check(_l.load().b200mf_laplace_vmult_dot_host(self.h, _dv(dst), _dv(src), C.byref(out)))
        return out.value

    def assemble_rhs(self, rhs, solution_old, source_q=None, source_const=0.0):
        """Cell loop of LaplaceProblem::assemble_rhs (heat_transfer.h:599-615): rhs += int (u_old + dt f) phi."""
        sq = None if source_q is None else np.ascontiguousarray(source_q, dtype=np.float64)
        _l.check(_l.load().b200mf_laplace_assemble_rhs(self.h, _dv(rhs), _dv(solution_old), _ptr(sq), float(source_const)))

    def vmult_add(self, dst, src):
        _l.check(_l.load().b200mf_laplace_vmult_add(self.h, _dv(dst), _dv(src)))

    def vmult_cells(self, dst_zero, src):
        """smoother hook: the operator pass alone (dst zero on entry, constrained rows left 0)"""
        _l.check(_l.load().b200mf_laplace_vmult_cells(self.h, _dv(dst_zero), _dv(src)))

    def Tvmult_add(self, dst, src):
        _l.check(_l.load().b200mf_laplace_Tvmult_add(self.h, _dv(dst), _dv(src)))

    def vmult_interface_down(self, dst, src):
        """MatrixFreeOperators::Base: dst = 0 + refinement-edge rows (none on globally refined meshes)."""
        _l.check(_l.load().b200mf_laplace_vmult_interface_down(self.h, _dv(dst), _dv(src)))

    def vmult_interface_up(self, dst, src):
        _l.check(_l.load().b200mf_laplace_vmult_interface_up(self.h, _dv(dst), _dv(src)))

    def m(self):
        return int(_l.load().b200mf_laplace_m(self.h))

    n = m

    def el(self, row, col=None):
        """Base::el(row, col): only the diagonal is available (deal.II asserts row == col)."""
        if col is not None and col != row:
            raise ValueError("el(): only diagonal entries are available (MatrixFreeOperators::Base::el)")
        out = C.c_double(0)
        _l.check(_l.load().b200mf_laplace_el(self.h, int(row), C.byref(out)))
        return out.value

    def initialize_dof_vector(self):
        """Base::initialize_dof_vector: a zeroed device vector on this operator's MatrixFree."""
        return DeviceVector(self.mf)

    def get_matrix_free(self):
        return self.mf

    def compute_diagonal(self):
        _l.check(_l.load().b200mf_laplace_compute_diagonal(self.h))

    def get_matrix_diagonal(self):
        d = np.empty(self.mf.n_dofs)
        _l.check(_l.load().b200mf_laplace_diagonal_host(self.h, _ptr(d), None))
        return d

    def get_matrix_diagonal_inverse(self):
        d = np.empty(self.mf.n_dofs)
        _l.check(_l.load().b200mf_laplace_diagonal_host(self.h, None, _ptr(d)))
        return d

    def __del__(self):
        if getattr(self, "h", None):
            try:
                _l.load().b200mf_laplace_destroy(self.h)
            except TypeError:  # interpreter shutdown: the module globals are already gone
                pass
            self.h = None


class InterfaceTraction:
    """Interface-face loop of Solid::assemble_residual (mf_elasticity.h:1849-1897) on a set of boundary faces.
    faces: iterable of (direction, side) of a StructuredMesh (every cell face on that side of the box), or an explicit
    list of (cell, face_no) pairs with face_no = 2*direction + side."""

    def __init__(self, mf: MatrixFree, mesh, faces, explicit=None):
        self.mf = mf
        if explicit is None:
            cells, fno = [], []
            for d, side in faces:
                on = np.flatnonzero(mesh.cell_ijk[:, d] == (0 if side == 0 else mesh.ncells[d] - 1))
                cells.append(on)
                fno.append(np.full(on.size, 2 * d + side))
            cells = np.concatenate(cells) if cells else np.zeros(0, np.int64)
            fno = np.concatenate(fno) if fno else np.zeros(0, np.int32)
        else:
            cells, fno = np.asarray(explicit[0]), np.asarray(explicit[1])
        self.cells = np.ascontiguousarray(cells, dtype=np.int64)
        self.face_no = np.ascontiguousarray(fno, dtype=np.int32)
        vtx = np.ascontiguousarray(mesh.vertices[self.cells], dtype=np.float64)
        h = C.c_void_p()
        _l.check(_l.load().b200mf_face_set_create(C.byref(h), mf.h, int(self.cells.size), _ptr(self.cells), _ptr(self.face_no), _ptr(vtx)))
        self.h = h
        self._traction = None

    def n_q_points(self):
        return int(_l.load().b200mf_face_set_n_q_points(self.h))

    def quadrature_points(self):
        out = np.empty((self.n_q_points(), 3))
        _l.check(_l.load().b200mf_face_set_quadrature_points(self.h, _ptr(out)))
        return out

    def set_traction(self, traction):
        """traction: one vector (constant) or [n_faces * nq^2, ncomp] values at the face quadrature points (ncomp = 3: traction
        of the solid; ncomp = 1 on a scalar context: the heat flux)."""
        t = np.ascontiguousarray(traction, dtype=np.float64)
        nc = self.mf.n_components
        if t.size == nc and self.n_q_points() != 1:
            _l.check(_l.load().b200mf_face_set_set_traction(self.h, None, _ptr(t)))
        else:
            assert t.size == nc * self.n_q_points()
            _l.check(_l.load().b200mf_face_set_set_traction(self.h, _ptr(t), None))
        self._traction = t

    def assemble(self, rhs, total_displacement, traction=None):
        if traction is not None and (self._traction is None or not np.array_equal(np.ravel(traction), np.ravel(self._traction))):
            self.set_traction(traction)
        _l.check(_l.load().b200mf_face_set_assemble_traction(self.h, _dv(rhs), _dv(total_displacement)))

    def assemble_flux(self, rhs, scale):
        """Scalar (heat) context: rhs += scale * int_Gamma flux(q) phi dA with the flux given to set_traction, one value per
        face quadrature point (heat_transfer.h:617-652: scale = -delta_t)."""
        _l.check(_l.load().b200mf_face_set_assemble_flux(self.h, _dv(rhs), float(scale)))

    def __del__(self):
        if getattr(self, "h", None):
            try:
                _l.load().b200mf_face_set_destroy(self.h)
            except TypeError:  # interpreter shutdown: the module globals are already gone
                pass
            self.h = None


class BoundaryMassOperator:
    """MatrixFreeOperators::BoundaryMassOperator (include/adapter/boundary_mass_operator.h) on the faces of an
    InterfaceTraction face set, and BoundaryProjector::project (include/adapter/boundary_projector.h:86-110)."""

    def __init__(self, faces: InterfaceTraction):
        self.faces, self.mf = faces, faces.mf
        h = C.c_void_p()
        _l.check(_l.load().b200mf_boundary_mass_create(C.byref(h), faces.h))
        self.h = h

    def vmult(self, dst, src):
        _l.check(_l.load().b200mf_boundary_mass_vmult(self.h, _dv(dst), _dv(src)))

    def vmult_add(self, dst, src):
        _l.check(_l.load().b200mf_boundary_mass_vmult_add(self.h, _dv(dst), _dv(src)))

    def compute_diagonal(self):
        _l.check(_l.load().b200mf_boundary_mass_compute_diagonal(self.h))

    def _diag(self, which):
        d, i = C.c_void_p(), C.c_void_p()
        _l.check(_l.load().b200mf_boundary_mass_diagonal(self.h, C.byref(d), C.byref(i)))
        out = np.empty(self.mf.n_dofs, dtype=np.float64)
        _l.check(_l.load().b200mf_vec_download_f64(self.mf.h, _ptr(out), C.c_void_p(d.value if which == 0 else i.value)))
        return out

    def diagonal(self):
        return self._diag(0)

    def inverse_diagonal(self):
        return self._diag(1)

    def restrict(self, dst, src):
        _l.check(_l.load().b200mf_boundary_mass_restrict(self.h, _dv(dst), _dv(src)))

    def project(self, residual):
        """residual <- M_Gamma^-1 residual|Gamma (CG, ReductionControl(500, 0, 1e-12), Jacobi 0.8, start = residual)"""
        its, res = C.c_int(), C.c_double()
        _l.check(_l.load().b200mf_boundary_mass_project(self.h, _dv(residual), C.byref(its), C.byref(res)))
        return dict(iterations=its.value, residual=res.value)

    def __del__(self):
        if getattr(self, "h", None):
            try:
                _l.load().b200mf_boundary_mass_destroy(self.h)
            except TypeError:  # interpreter shutdown: the module globals are already gone
                pass
            self.h = None
