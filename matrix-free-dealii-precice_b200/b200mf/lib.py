"""ctypes binding of libb200mf.so (include/b200mf.h).  There is no CPU fallback:
a missing library or a missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MF_LIBRARY") or os.path.join(_HERE, "libb200mf.so")  # override: tuning builds only

F64, F32 = 0, 1
CACHING = {"scalar_referential": 0, "tensor2": 1, "tensor4": 2, "tensor4_ns": 3}


class B200mfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200mf status {code}: {msg}")
        self.code = code


class CtxDesc(C.Structure):
    _fields_ = [
        ("degree", C.c_int), ("n_q_points_1d", C.c_int), ("n_components", C.c_int), ("dtype", C.c_int),
        ("n_cells", C.c_int64), ("n_dofs", C.c_int64), ("n_owned", C.c_int64),
        ("dof_indices", C.c_void_p), ("node_indices", C.c_void_p),
        ("constrained_dofs", C.c_void_p), ("n_constrained", C.c_int64),
        ("vertices", C.c_void_p), ("inverse_jacobian", C.c_void_p), ("jxw", C.c_void_p),
        ("material_id", C.c_void_p),
    ]


class PartitionerDesc(C.Structure):
    _fields_ = [("n_neighbors", C.c_int), ("neighbor_ranks", C.c_void_p), ("send_offsets", C.c_void_p),
                ("send_indices", C.c_void_p), ("ghost_offsets", C.c_void_p), ("ghost_counts", C.c_void_p)]


class Material(C.Structure):
    _fields_ = [("mu", C.c_double), ("nu", C.c_double), ("rho", C.c_double), ("alpha_1", C.c_double),
                ("formulation", C.c_int)]


class ChebyshevData(C.Structure):
    _fields_ = [("degree", C.c_int), ("smoothing_range", C.c_double), ("eig_cg_n_iterations", C.c_int),
                ("eig_cg_residual", C.c_double), ("max_eigenvalue", C.c_double), ("first_global_index", C.c_int64),
                ("n_global_dofs", C.c_int64)]


class CGControl(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("max_iterations", C.c_int), ("preconditioner", C.c_int),
                ("jacobi_omega", C.c_double), ("mg", C.c_void_p), ("x_is_zero", C.c_int)]


class CGResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("residual", C.c_double), ("initial_residual", C.c_double),
                ("converged", C.c_int), ("min_eigenvalue", C.c_double), ("max_eigenvalue", C.c_double)]


_lib = None

_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/b200mf.h declares
SIGNATURES = {
    "b200mf_last_error": (C.c_char_p, []),
    "b200mf_launch_count": (_i64, []),
    "b200mf_version": (_i32, []),
    "b200mf_ctx_create": (_i32, [_PP, C.POINTER(CtxDesc)]),
    "b200mf_ctx_destroy": (_i32, [_vp]),
    "b200mf_ctx_set_stream": (_i32, [_vp, _vp]),
    "b200mf_ctx_synchronize": (_i32, [_vp]),
    "b200mf_ctx_time_kernels": (_i32, [_vp, _i32]),
    "b200mf_ctx_kernel_time": (_i32, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "b200mf_ctx_set_overlap": (_i32, [_vp, _i32]),
    "b200mf_ctx_cell_type": (_i32, [_vp]),
    "b200mf_ctx_n_dofs": (_i64, [_vp]),
    "b200mf_ctx_memory_consumption": (C.c_size_t, [_vp]),
    "b200mf_nh_assemble_residual": (_i32, [_vp, _vp, _vp, _vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "b200mf_vec_set_zero_constrained": (_i32, [_vp, _vp]),
    "b200mf_nh_vmult_dot": (_i32, [_vp, _vp, _vp, _vp]),
    "b200mf_nh_vmult_dot_host": (_i32, [_vp, _vp, _vp, C.POINTER(C.c_double)]),
    "b200mf_laplace_vmult_dot_host": (_i32, [_vp, _vp, _vp, C.POINTER(C.c_double)]),
    "b200mf_laplace_vmult_dot": (_i32, [_vp, _vp, _vp, _vp]),
    "b200mf_laplace_assemble_rhs": (_i32, [_vp, _vp, _vp, _vp, _dbl]),
    "b200mf_vec_create": (_i32, [_vp, _PP]),
    "b200mf_vec_destroy": (_i32, [_vp, _vp]),
    "b200mf_vec_upload": (_i32, [_vp, _vp, _vp]),
    "b200mf_vec_download": (_i32, [_vp, _vp, _vp]),
    "b200mf_vec_upload_f64": (_i32, [_vp, _vp, _vp]),
    "b200mf_vec_download_f64": (_i32, [_vp, _vp, _vp]),
    "b200mf_comm_unique_id": (_i32, [_vp]),
    "b200mf_comm_create": (_i32, [_PP, _i32, _i32, _vp]),
    "b200mf_comm_destroy": (_i32, [_vp]),
    "b200mf_ctx_set_partitioner": (_i32, [_vp, _vp, C.POINTER(PartitionerDesc)]),
    "b200mf_ctx_exchange_is_p2p": (_i32, [_vp]),
    "b200mf_vec_update_ghost_values": (_i32, [_vp, _vp]),
    "b200mf_vec_compress_add": (_i32, [_vp, _vp]),
    "b200mf_vec_zero_out_ghost_values": (_i32, [_vp, _vp]),
    "b200mf_nh_create": (_i32, [_PP, _vp, C.POINTER(Material), C.POINTER(Material)]),
    "b200mf_nh_destroy": (_i32, [_vp]),
    "b200mf_nh_initialize": (_i32, [_vp, _i32]),
    "b200mf_nh_cache": (_i32, [_vp, _vp, C.POINTER(_i64)]),
    "b200mf_nh_cache_host": (_i32, [_vp, _vp, C.POINTER(_i64)]),
    "b200mf_nh_vmult": (_i32, [_vp, _vp, _vp]),
    "b200mf_nh_vmult_add": (_i32, [_vp, _vp, _vp]),
    "b200mf_nh_vmult_cells": (_i32, [_vp, _vp, _vp]),
    "b200mf_nh_diagonal_generation": (C.c_uint64, [_vp]),
    "b200mf_laplace_vmult_cells": (_i32, [_vp, _vp, _vp]),
    "b200mf_laplace_diagonal_generation": (C.c_uint64, [_vp]),
    "b200mf_nh_Tvmult": (_i32, [_vp, _vp, _vp]),
    "b200mf_nh_Tvmult_add": (_i32, [_vp, _vp, _vp]),
    "b200mf_nh_vmult_host": (_i32, [_vp, _vp, _vp]),
    "b200mf_nh_compute_diagonal": (_i32, [_vp]),
    "b200mf_nh_diagonal": (_i32, [_vp, _PP, _PP]),
    "b200mf_nh_diagonal_host": (_i32, [_vp, _vp, _vp]),
    "b200mf_nh_precondition_jacobi": (_i32, [_vp, _vp, _vp, _dbl]),
    "b200mf_nh_el": (_i32, [_vp, _i64, C.POINTER(_dbl)]),
    "b200mf_nh_m": (_i64, [_vp]),
    "b200mf_nh_memory_consumption": (C.c_size_t, [_vp]),
    "b200mf_nh_qdata_host": (_i32, [_vp, _i32, _vp]),
    "b200mf_nh_n_qfields": (_i32, [_vp]),
    "b200mf_laplace_create": (_i32, [_PP, _vp, _dbl, _vp]),
    "b200mf_laplace_destroy": (_i32, [_vp]),
    "b200mf_laplace_set_delta_t": (_i32, [_vp, _dbl]),
    "b200mf_laplace_vmult": (_i32, [_vp, _vp, _vp]),
    "b200mf_laplace_vmult_add": (_i32, [_vp, _vp, _vp]),
    "b200mf_laplace_Tvmult": (_i32, [_vp, _vp, _vp]),
    "b200mf_laplace_Tvmult_add": (_i32, [_vp, _vp, _vp]),
    "b200mf_laplace_vmult_interface_down": (_i32, [_vp, _vp, _vp]),
    "b200mf_laplace_vmult_interface_up": (_i32, [_vp, _vp, _vp]),
    "b200mf_laplace_m": (_i64, [_vp]),
    "b200mf_laplace_el": (_i32, [_vp, _i64, C.POINTER(_dbl)]),
    "b200mf_laplace_vmult_host": (_i32, [_vp, _vp, _vp]),
    "b200mf_laplace_compute_diagonal": (_i32, [_vp]),
    "b200mf_laplace_diagonal": (_i32, [_vp, _PP, _PP]),
    "b200mf_laplace_diagonal_host": (_i32, [_vp, _vp, _vp]),
    "b200mf_face_set_create": (_i32, [_PP, _vp, _i64, _vp, _vp, _vp]),
    "b200mf_face_set_destroy": (_i32, [_vp]),
    "b200mf_face_set_n_q_points": (_i64, [_vp]),
    "b200mf_face_set_quadrature_points": (_i32, [_vp, _vp]),
    "b200mf_face_set_set_traction": (_i32, [_vp, _vp, _vp]),
    "b200mf_face_set_assemble_traction": (_i32, [_vp, _vp, _vp]),
    "b200mf_face_set_assemble_flux": (_i32, [_vp, _vp, _dbl]),
    "b200mf_boundary_mass_create": (_i32, [_PP, _vp]),
    "b200mf_boundary_mass_destroy": (_i32, [_vp]),
    "b200mf_boundary_mass_ctx": (_vp, [_vp]),
    "b200mf_boundary_mass_vmult": (_i32, [_vp, _vp, _vp]),
    "b200mf_boundary_mass_vmult_add": (_i32, [_vp, _vp, _vp]),
    "b200mf_boundary_mass_compute_diagonal": (_i32, [_vp]),
    "b200mf_boundary_mass_diagonal": (_i32, [_vp, _PP, _PP]),
    "b200mf_boundary_mass_restrict": (_i32, [_vp, _vp, _vp]),
    "b200mf_boundary_mass_project": (_i32, [_vp, _vp, C.POINTER(_i32), C.POINTER(_dbl)]),
    "b200mf_cg_create_boundary_mass": (_i32, [_PP, _vp]),
    "b200mf_nh_ctx": (_vp, [_vp]),
    "b200mf_laplace_ctx": (_vp, [_vp]),
    "b200mf_vec_dot": (_i32, [_vp, _vp, _vp, C.POINTER(_dbl)]),
    "b200mf_vec_l2_norm": (_i32, [_vp, _vp, C.POINTER(_dbl)]),
    "b200mf_vec_sadd": (_i32, [_vp, _vp, _dbl, _dbl, _vp]),
    "b200mf_vec_set_zero": (_i32, [_vp, _vp]),
    "b200mf_vec_convert": (_i32, [_vp, _vp, _vp, _vp]),
    "b200mf_chebyshev_create_nh": (_i32, [_PP, _vp, C.POINTER(ChebyshevData)]),
    "b200mf_chebyshev_create_laplace": (_i32, [_PP, _vp, C.POINTER(ChebyshevData)]),
    "b200mf_chebyshev_destroy": (_i32, [_vp]),
    "b200mf_chebyshev_estimate_eigenvalues": (_i32, [_vp, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_i32), C.POINTER(_i32)]),
    "b200mf_chebyshev_vmult": (_i32, [_vp, _vp, _vp]),
    "b200mf_chebyshev_step": (_i32, [_vp, _vp, _vp]),
    "b200mf_mg_transfer_create": (_i32, [_PP, _vp, _vp, _vp]),
    "b200mf_mg_transfer_destroy": (_i32, [_vp]),
    "b200mf_mg_transfer_prolongate": (_i32, [_vp, _vp, _vp]),
    "b200mf_mg_transfer_prolongate_and_add": (_i32, [_vp, _vp, _vp]),
    "b200mf_mg_transfer_restrict_and_add": (_i32, [_vp, _vp, _vp]),
    "b200mf_mg_transfer_interpolate": (_i32, [_vp, _vp, _vp]),
    "b200mf_mg_create": (_i32, [_PP, _i32, _PP, _PP, _vp]),
    "b200mf_mg_destroy": (_i32, [_vp]),
    "b200mf_mg_vmult": (_i32, [_vp, _vp, _vp]),
    "b200mf_mg_set_coarse_solver": (_i32, [_vp, _i32, _i32, _dbl, _dbl]),
    "b200mf_mg_coarse_iterations": (_i32, [_vp]),
    "b200mf_mg_set_graph_levels": (_i32, [_vp, _i32]),
    "b200mf_mg_graph_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "b200mf_cg_create_nh": (_i32, [_PP, _vp]),
    "b200mf_cg_create_laplace": (_i32, [_PP, _vp]),
    "b200mf_cg_destroy": (_i32, [_vp]),
    "b200mf_cg_solve": (_i32, [_vp, _vp, _vp, C.POINTER(CGControl), C.POINTER(CGResult)]),
}


def load():
    """Load libb200mf.so and bind every symbol; raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                              "(make -C matrix-free-dealii-precice_b200/csrc); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise B200mfError(rc, load().b200mf_last_error().decode())
    return rc
