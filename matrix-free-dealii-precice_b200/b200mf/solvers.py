"""Host-side mirror of the deal.II solver classes the reference instantiates on its operators
(SURVEY.md §8a a15); each one forwards to the device implementation in libb200mf.so.

  PreconditionChebyshev   <- PreconditionChebyshev<LevelMatrixType, LevelVectorType>   (mf_elasticity.h:371-372, 1149-1180)
  MGTransferMatrixFree    <- MGTransferMatrixFree<dim, float>                          (mf_elasticity.h:872-874)
  PreconditionMG          <- Multigrid<LevelVectorType> + PreconditionMG               (mf_elasticity.h:1209-1224, 1272-1276)
  SolverCG                <- SolverCG<VectorType>                                      (mf_elasticity.h:2024-2051)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _l
from .operators import DeviceVector, LaplaceOperator, NeoHookOperator, _dv

INVALID_UNSIGNED_INT = -1  # numbers::invalid_unsigned_int: Chebyshev in solver mode


class PreconditionChebyshev:
    """AdditionalData keywords as in deal.II; the inner preconditioner is the operator's inverse diagonal
    (`smoother_data[level].preconditioner = op.get_matrix_diagonal_inverse()`, mf_elasticity.h:1176-1177),
    so `op.compute_diagonal()` must have been called."""

    def __init__(self, op, degree=1, smoothing_range=0.0, eig_cg_n_iterations=8, eig_cg_residual=1e-2, max_eigenvalue=1.0,
                 first_global_index=0, n_global_dofs=0):
        self.op = op
        d = _l.ChebyshevData(int(degree), float(smoothing_range), int(eig_cg_n_iterations), float(eig_cg_residual),
                             float(max_eigenvalue), int(first_global_index), int(n_global_dofs))
        h = C.c_void_p()
        L = _l.load()
        if isinstance(op, NeoHookOperator):
            _l.check(L.b200mf_chebyshev_create_nh(C.byref(h), op.h, C.byref(d)))
        elif isinstance(op, LaplaceOperator):
            _l.check(L.b200mf_chebyshev_create_laplace(C.byref(h), op.h, C.byref(d)))
        else:
            raise TypeError("PreconditionChebyshev needs a NeoHookOperator or LaplaceOperator")
        self.h = h

    def estimate_eigenvalues(self):
        mn, mx, its, deg = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        _l.check(_l.load().b200mf_chebyshev_estimate_eigenvalues(self.h, C.byref(mn), C.byref(mx), C.byref(its), C.byref(deg)))
        return dict(min_eigenvalue_estimate=mn.value, max_eigenvalue_estimate=mx.value, cg_iterations=its.value,
                    degree=deg.value)

    def vmult(self, dst, src):
        _l.check(_l.load().b200mf_chebyshev_vmult(self.h, _dv(dst), _dv(src)))

    def step(self, dst, src):
        _l.check(_l.load().b200mf_chebyshev_step(self.h, _dv(dst), _dv(src)))

    def __del__(self):
        if getattr(self, "h", None):
            try:
                _l.load().b200mf_chebyshev_destroy(self.h)
            except TypeError:  # interpreter shutdown: the module globals are already gone
                pass
            self.h = None


class MGTransferMatrixFree:
    """One level pair (coarse, fine) of the transfer."""

    def __init__(self, mf_coarse, mf_fine, children):
        self.coarse, self.fine = mf_coarse, mf_fine
        ch = np.ascontiguousarray(children, dtype=np.int64)
        assert ch.shape == (mf_coarse.n_cells, 8)
        h = C.c_void_p()
        _l.check(_l.load().b200mf_mg_transfer_create(C.byref(h), mf_coarse.h, mf_fine.h, ch.ctypes.data_as(C.c_void_p)))
        self.h = h

    def prolongate(self, fine_dst, coarse_src):
        _l.check(_l.load().b200mf_mg_transfer_prolongate(self.h, _dv(fine_dst), _dv(coarse_src)))

    def prolongate_and_add(self, fine_dst, coarse_src):
        _l.check(_l.load().b200mf_mg_transfer_prolongate_and_add(self.h, _dv(fine_dst), _dv(coarse_src)))

    def restrict_and_add(self, coarse_dst, fine_src):
        _l.check(_l.load().b200mf_mg_transfer_restrict_and_add(self.h, _dv(coarse_dst), _dv(fine_src)))

    def interpolate(self, coarse_dst, fine_src):
        _l.check(_l.load().b200mf_mg_transfer_interpolate(self.h, _dv(coarse_dst), _dv(fine_src)))

    def __del__(self):
        if getattr(self, "h", None):
            try:
                _l.load().b200mf_mg_transfer_destroy(self.h)
            except TypeError:  # interpreter shutdown: the module globals are already gone
                pass
            self.h = None


class PreconditionMG:
    """V-cycle over `smoothers` (level 0 = coarse solver) and `transfers` (level l-1 <-> l) applied to vectors of
    the system context `mf_system` (copy_to_mg / copy_from_mg inside)."""

    def __init__(self, smoothers, transfers, mf_system):
        self.smoothers, self.transfers, self.mf = list(smoothers), list(transfers), mf_system
        n = len(self.smoothers)
        assert len(self.transfers) == n - 1
        sm = (C.c_void_p * n)(*[s.h for s in self.smoothers])
        tr = (C.c_void_p * max(1, n - 1))(*[t.h for t in self.transfers])
        h = C.c_void_p()
        _l.check(_l.load().b200mf_mg_create(C.byref(h), n, sm, tr, mf_system.h))
        self.h = h

    def vmult(self, dst, src):
        _l.check(_l.load().b200mf_mg_vmult(self.h, _dv(dst), _dv(src)))

    def set_coarse_solver_cg(self, max_iterations, abs_tolerance=1e-10, reduction=1e-3):
        """MGCoarseGridIterativeSolver: SolverCG + ReductionControl on level 0, preconditioned with the level-0 smoother
        (mf_elasticity.h:1183-1198, 'MF Chebyshev coarse = false')."""
        _l.check(_l.load().b200mf_mg_set_coarse_solver(self.h, 1, int(max_iterations), float(abs_tolerance), float(reduction)))

    def set_coarse_solver_chebyshev(self):
        _l.check(_l.load().b200mf_mg_set_coarse_solver(self.h, 0, 0, 0.0, 0.0))

    def coarse_iterations(self):
        return int(_l.load().b200mf_mg_coarse_iterations(self.h))

    def set_graph_levels(self, top_level):
        """Levels 0..top_level replayed from a CUDA graph (default n_levels-2; -1 = plain launches)."""
        _l.check(_l.load().b200mf_mg_set_graph_levels(self.h, int(top_level)))

    def graph_info(self):
        top, state = C.c_int(), C.c_int()
        rec, rep, ln = C.c_int64(), C.c_int64(), C.c_int64()
        _l.check(_l.load().b200mf_mg_graph_info(self.h, C.byref(top), C.byref(state), C.byref(rec), C.byref(rep), C.byref(ln)))
        return dict(top_level=top.value, state=state.value, records=rec.value, replays=rep.value, launches_per_replay=ln.value)

    def __del__(self):
        if getattr(self, "h", None):
            try:
                _l.load().b200mf_mg_destroy(self.h)
            except TypeError:  # interpreter shutdown: the module globals are already gone
                pass
            self.h = None


class SolverCG:
    """solve(A, x, b, preconditioner) with preconditioner in {None, ("jacobi", omega), PreconditionMG}."""

    def __init__(self, tolerance, max_iterations):
        self.tolerance, self.max_iterations = float(tolerance), int(max_iterations)
        self.last = None
        self._h = None
        self._op = None

    def _handle(self, A):
        if self._op is not A:
            self._free()
            h = C.c_void_p()
            L = _l.load()
            if isinstance(A, NeoHookOperator):
                _l.check(L.b200mf_cg_create_nh(C.byref(h), A.h))
            else:
                _l.check(L.b200mf_cg_create_laplace(C.byref(h), A.h))
            self._h, self._op = h, A
        return self._h

    def solve(self, A, x, b, preconditioner=None, x_is_zero=False):
        ctl = _l.CGControl(self.tolerance, self.max_iterations, 0, 1.0, None, int(bool(x_is_zero)))
        if isinstance(preconditioner, PreconditionMG):
            ctl.preconditioner, ctl.mg = 2, preconditioner.h
        elif isinstance(preconditioner, tuple) and preconditioner[0] == "jacobi":
            ctl.preconditioner, ctl.jacobi_omega = 1, float(preconditioner[1])
        elif preconditioner is not None:
            raise TypeError("unknown preconditioner")
        res = _l.CGResult()
        rc = _l.load().b200mf_cg_solve(self._handle(A), _dv(x), _dv(b), C.byref(ctl), C.byref(res))
        self.last = dict(iterations=res.iterations, residual=res.residual, initial_residual=res.initial_residual,
                         converged=bool(res.converged), min_eigenvalue=res.min_eigenvalue, max_eigenvalue=res.max_eigenvalue)
        _l.check(rc)
        return self.last

    def last_step(self):
        return self.last["iterations"]

    def _free(self):
        if self._h:
            _l.load().b200mf_cg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass
