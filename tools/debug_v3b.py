import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "matrix-free-dealii-precice_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from b200mf import MatrixFree, NeoHookOperator, LaplaceOperator
from common import MAT, c_cells, displacement, make_mesh, rel
from oracle import c_oracle as CO
p = int(sys.argv[1]); nc = tuple(int(x) for x in sys.argv[2].split(","))
m = make_mesh(p, nc, mapped=False, permute=None, clamp=None)
n = p + 1
src = np.random.default_rng(0).uniform(-1, 1, m.n_nodes)
ref = np.zeros_like(src); CO.LaplaceOperator(c_cells(m, 1), 0.1).vmult(ref, src)
mf = MatrixFree.from_mesh(m, 1); out = np.zeros_like(src); LaplaceOperator(mf, 0.1).vmult(out, src)
print("laplace rel", rel(out, ref), "cells", m.n_cells)
bad = np.abs(out - ref) > 1e-9 * np.abs(ref).max()
print("bad nodes", bad.sum(), "of", bad.size)
if bad.any():
    g = m.node_grid_index()[bad]
    print("bad node grid idx sample", g[:10].tolist(), "x range", g[:, 0].min(), g[:, 0].max(), "y", g[:, 1].min(), g[:, 1].max(), "z", g[:, 2].min(), g[:, 2].max())
u = displacement(m)
mf3 = MatrixFree.from_mesh(m, 3)
op = NeoHookOperator(); op.set_material(mf3, dict(formulation=1, **MAT)); op.initialize(u, "tensor2")
try:
    op.cache()
except Exception as ex:
    print("cache:", ex)
orc = CO.NeoHookOperator(c_cells(m), formulation=1, **MAT); orc.initialize(u, "tensor2"); orc.cache()
import ctypes as C
a = op.qdata(11)  # 1/J
from common import np_cells
from oracle import np_oracle as NO
nop = NO.NeoHookOperator(np_cells(m), NO.Material(formulation=1, **MAT)); nop.initialize(u, "tensor2"); nop.cache()
b = nop.cached_second_scalar
badc = np.abs(a - b).max(axis=1) > 1e-9
print("bad cells", badc.sum(), "of", badc.size, "first bad", np.flatnonzero(badc)[:40].tolist())
if badc.any():
    c0 = np.flatnonzero(badc)[0]
    print("cell", c0, "err pattern (k,j,i):\n", np.round(np.abs(a - b)[c0].reshape(n, n, n) / np.abs(b).max(), 3))
