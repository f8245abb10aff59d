"""Shared builders for the parity tests: synthetic meshes + the oracle objects
(test infrastructure only)."""
import numpy as np

from b200mf.mesh import StructuredMesh, smooth_mapping, synthetic_displacement
from oracle import c_oracle as CO
from oracle import np_oracle as NO

MAT = dict(mu=0.5e6, nu=0.4, rho=1000.0, alpha_1=1.0 / (0.25 * 0.01**2))  # elasticity.prm, mf_elasticity.h:237-240


def make_mesh(p, ncells=(3, 2, 2), mapped=True, permute=1, clamp=((0, 0),), ncomp=3):
    m = StructuredMesh(ncells, p, mapping=smooth_mapping(0.05) if mapped else None, permute_seed=permute)
    if clamp:
        m.clamp(list(clamp), ncomp)
    return m


def np_cells(m, ncomp=3):
    nq = m.p + 1
    invjac, jxw, _ = NO.q1_metric(m.vertices, nq)
    return NO.CellData(m.p, nq, m.dim, ncomp, m.dof_index(ncomp), m.constrained_dofs(ncomp), invjac, jxw,
                       m.material_id)


def c_cells(m, ncomp=3):
    nq = m.p + 1
    invjac, jxw = CO.q1_metric(m.vertices, nq)
    return CO.Cells(m.p, nq, ncomp, m.dof_index(ncomp), m.constrained_dofs(ncomp), invjac, jxw, m.material_id)


def displacement(m, amplitude=0.05):
    return synthetic_displacement(m.node_coordinates(), amplitude)


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
