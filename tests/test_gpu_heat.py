"""GPU parity of whole backward-Euler steps of the heat solver (BASELINE.json configs[1]): b200mf.heat.
HeatProblem (assemble_rhs cell loop, level diagonals, Chebyshev estimates, GMG-CG - all through the C ABI; the scalar
operator on affine cells runs the stacked warp kernel for p <= 4) against the CPU restatement of the same steps
(oracle/np_oracle.heat_rhs_cells + oracle/solver_oracle.py, heat_transfer.h:577-812)."""
import numpy as np
import pytest

from b200mf.heat import HeatProblem
from common import np_cells, rel
from oracle import np_oracle as NO
from oracle import solver_oracle as SO
from test_solver_oracle import build_gmg, hierarchy, laplace_level

pytestmark = pytest.mark.gpu
ALL_FACES = ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1))


@pytest.mark.parametrize("p,levels,mapped", [(1, 3, False), (2, 2, False), (4, 2, False), (3, 2, True), (5, 1, False)])
def test_heat_time_steps(p, levels, mapped):
    meshes = hierarchy(p, (2, 2, 2), levels, 1, mapped=mapped, clamp=ALL_FACES)
    m = meshes[-1]
    dt, f = 0.1, 2.5
    hp = HeatProblem(meshes, dt, source_const=f)
    assert hp.mf.all_cells_affine() == (not mapped)
    x = m.node_coordinates()
    u0 = np.prod(np.sin(np.pi * x), axis=1)
    con = m.constrained_dofs(1)
    u0[con] = 0.0
    hp.solution.upload(u0)
    ops_ref, As, mg_ref = build_gmg(meshes, 1, lambda mm: laplace_level(mm, dt))
    cells = np_cells(m, 1)
    u_ref = u0.copy()
    for step in range(2):
        res = hp.time_step()
        rhs = NO.heat_rhs_cells(cells, dt, u_ref, None, f)  # (u_old + dt f, phi), constrained rows skipped (:599-615)
        rhs[con] = 0.0
        assert rel(hp.system_rhs.download(), rhs) < 1e-12
        xr = np.zeros_like(rhs)
        r = SO.solver_cg(As[-1], xr, rhs, mg_ref.vmult if levels > 1 else mg_ref.vmult, 1e-12 * np.linalg.norm(rhs), rhs.size)
        xr[con] = 0.0
        u_ref = xr
        assert res["converged"]
        assert abs(res["iterations"] - r.iterations) <= 1, (res["iterations"], r.iterations)
        assert rel(hp.solution.download(), u_ref) < 1e-8
