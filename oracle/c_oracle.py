"""TEST INFRASTRUCTURE — ctypes front-end of oracle/mf_oracle.c (the
sum-factorised CPU restatement).  Same import restrictions as np_oracle.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LEVELS = {"scalar_referential": 0, "tensor2": 1, "tensor4": 2, "tensor4_ns": 3}


def build(force=False):
    so = os.path.join(_HERE, "libmf_oracle.so")
    src = os.path.join(_HERE, "mf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
        L.orc_shape_1d.argtypes = [i32, i32] + [vp] * 6
        L.orc_q1_metric.argtypes = [i64, vp, i32, vp, vp]
        L.orc_cells_create.restype = vp
        L.orc_cells_create.argtypes = [i32, i32, i32, i64, i64, vp, vp, vp, vp, vp]
        L.orc_cells_destroy.argtypes = [vp]
        L.orc_nh_create.restype = vp
        L.orc_nh_create.argtypes = [vp, dbl, dbl, dbl, dbl, dbl, i32, i32]
        L.orc_nh_destroy.argtypes = [vp]
        L.orc_nh_cache.restype = i64
        L.orc_nh_cache.argtypes = [vp, vp]
        L.orc_nh_vmult.argtypes = [vp, vp, vp]
        L.orc_nh_compute_diagonal.argtypes = [vp, vp, vp]
        L.orc_laplace_vmult.argtypes = [vp, dbl, vp, vp, vp]
        L.orc_laplace_compute_diagonal.argtypes = [vp, dbl, vp, vp, vp]
        L.orc_num_threads.restype = i32
        L.orc_nh_simd_create.restype = vp
        L.orc_nh_simd_create.argtypes = [vp]
        L.orc_nh_simd_destroy.argtypes = [vp]
        L.orc_nh_vmult_simd.argtypes = [vp, vp, vp, vp]
        L.orc_simd_isa.restype = C.c_char_p
        L.orc_stream_triad_gbs.restype = C.c_double
        L.orc_stream_triad_gbs.argtypes = [C.c_long, C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def shape_1d(p, nq):
    n = p + 1
    S, D = np.zeros((nq, n)), np.zeros((nq, n))
    Cc, w, nodes, qp = np.zeros((nq, nq)), np.zeros(nq), np.zeros(n), np.zeros(nq)
    lib().orc_shape_1d(p, nq, _p(S), _p(D), _p(Cc), _p(w), _p(nodes), _p(qp))
    return S, D, Cc, w, nodes, qp


def q1_metric(vertices, nq):
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    nc = v.shape[0]
    invjac = np.zeros((nc, nq**3, 3, 3))
    jxw = np.zeros((nc, nq**3))
    lib().orc_q1_metric(nc, _p(v), nq, _p(invjac), _p(jxw))
    return invjac, jxw


class Cells:
    def __init__(self, p, nq, ncomp, dof_idx, constrained, invjac, jxw, material_id=None):
        self.p, self.nq, self.ncomp = p, nq, ncomp
        self.dof_idx = np.ascontiguousarray(dof_idx, dtype=np.uint32)
        self.constrained = np.ascontiguousarray(constrained, dtype=np.uint8)
        self.invjac = np.ascontiguousarray(invjac, dtype=np.float64)
        self.jxw = np.ascontiguousarray(jxw, dtype=np.float64)
        self.material_id = None if material_id is None else np.ascontiguousarray(material_id, dtype=np.uint8)
        self.n_cells = self.dof_idx.shape[0]
        self.n_dofs = self.constrained.shape[0]
        self.h = lib().orc_cells_create(p, nq, ncomp, self.n_cells, self.n_dofs, _p(self.dof_idx),
                                        _p(self.constrained), _p(self.invjac), _p(self.jxw), _p(self.material_id))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_cells_destroy(self.h)
            self.h = None


class NeoHookOperator:
    """mf_nh_operator.h:99-237 surface on numpy vectors."""

    def __init__(self, cells: Cells, mu, nu, rho, alpha_1, formulation=1, mu_inclusion=None):
        self.cells = cells
        self.params = (mu, nu, rho, alpha_1, mu if mu_inclusion is None else mu_inclusion)
        self.formulation = formulation
        self.h = None

    def initialize(self, displacement, caching):
        if caching not in LEVELS:
            raise ValueError("Unknown caching")
        self.displacement = np.ascontiguousarray(displacement, dtype=np.float64)
        if self.h:
            lib().orc_nh_destroy(self.h)
        self.h = lib().orc_nh_create(self.cells.h, *self.params, LEVELS[caching], self.formulation)

    def cache(self):
        bad = lib().orc_nh_cache(self.h, _p(self.displacement))
        if bad:
            raise ArithmeticError("det_F is not positive.")

    def vmult(self, dst, src):
        src = np.ascontiguousarray(src, dtype=np.float64)
        lib().orc_nh_vmult(self.h, _p(dst), _p(src))

    Tvmult = vmult

    def vmult_simd(self, dst, src):
        """Same operator, SIMD over cells (8 cells per batch, as deal.II's VectorizedArray); tensor2 / formulation 1 only.
        Returns False if the operator has no batched form (other levels)."""
        if getattr(self, "_simd", None) is None:
            self._simd = lib().orc_nh_simd_create(self.h)
        if not self._simd:
            return False
        src = np.ascontiguousarray(src, dtype=np.float64)
        lib().orc_nh_vmult_simd(self.h, self._simd, _p(dst), _p(src))
        return True

    def compute_diagonal(self):
        n = self.cells.n_dofs
        self.diagonal, self.inverse_diagonal = np.zeros(n), np.zeros(n)
        lib().orc_nh_compute_diagonal(self.h, _p(self.diagonal), _p(self.inverse_diagonal))
        return self.diagonal

    def __del__(self):
        if getattr(self, "_simd", None):
            lib().orc_nh_simd_destroy(self._simd)
            self._simd = None
        if getattr(self, "h", None):
            lib().orc_nh_destroy(self.h)
            self.h = None


def simd_isa():
    return lib().orc_simd_isa().decode()


def host_stream_triad_gbs(n=100_000_000, reps=5):
    """STREAM-triad bandwidth of the host with all OpenMP threads (24 B per element), best of `reps` passes."""
    return float(lib().orc_stream_triad_gbs(int(n), int(reps)))


class LaplaceOperator:
    def __init__(self, cells: Cells, delta_t, coefficient=None):
        self.cells, self.delta_t = cells, float(delta_t)
        self.coef = None if coefficient is None else np.ascontiguousarray(coefficient, dtype=np.float64)

    def vmult(self, dst, src):
        src = np.ascontiguousarray(src, dtype=np.float64)
        lib().orc_laplace_vmult(self.cells.h, self.delta_t, _p(self.coef), _p(dst), _p(src))

    Tvmult = vmult

    def vmult_simd(self, dst, src):
        """Same operator, SIMD over cells (8 cells per batch, as deal.II's VectorizedArray); tensor2 / formulation 1 only.
        Returns False if the operator has no batched form (other levels)."""
        if getattr(self, "_simd", None) is None:
            self._simd = lib().orc_nh_simd_create(self.h)
        if not self._simd:
            return False
        src = np.ascontiguousarray(src, dtype=np.float64)
        lib().orc_nh_vmult_simd(self.h, self._simd, _p(dst), _p(src))
        return True

    def compute_diagonal(self):
        n = self.cells.n_dofs
        self.diagonal, self.inverse_diagonal = np.zeros(n), np.zeros(n)
        lib().orc_laplace_compute_diagonal(self.cells.h, self.delta_t, _p(self.coef), _p(self.diagonal),
                                           _p(self.inverse_diagonal))
        return self.diagonal


def num_threads():
    return lib().orc_num_threads()
