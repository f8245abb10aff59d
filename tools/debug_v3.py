import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "matrix-free-dealii-precice_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from b200mf import MatrixFree, NeoHookOperator, LaplaceOperator
from common import MAT, c_cells, displacement, make_mesh, rel
from oracle import c_oracle as CO
p = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nc = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (1, 1, 1)
m = make_mesh(p, nc, mapped=False, permute=None, clamp=None)
n = p + 1
# Laplace first (scalar, simplest)
src = np.random.default_rng(0).uniform(-1, 1, m.n_nodes)
ref = np.zeros_like(src); CO.LaplaceOperator(c_cells(m, 1), 0.1).vmult(ref, src)
mf = MatrixFree.from_mesh(m, 1); out = np.zeros_like(src); LaplaceOperator(mf, 0.1).vmult(out, src)
print("laplace rel", rel(out, ref))
if rel(out, ref) > 1e-10:
    e = np.abs(out - ref).reshape([c * p + 1 for c in nc][::-1]); print(np.round(e / np.abs(ref).max(), 3))
u = displacement(m)
mf3 = MatrixFree.from_mesh(m, 3)
op = NeoHookOperator(); op.set_material(mf3, dict(formulation=1, **MAT)); op.initialize(u, "tensor2")
try:
    op.cache()
    from common import np_cells
    from oracle import np_oracle as NO
    nop = NO.NeoHookOperator(np_cells(m), NO.Material(formulation=1, **MAT)); nop.initialize(u, "tensor2"); nop.cache()
    for f in (9, 10, 11):
        a = op.qdata(f); b = {9: nop.JxW_cur, 10: nop.cached_scalar, 11: nop.cached_second_scalar}[f]
        print("field", f, "rel", rel(a, b))
        if rel(a, b) > 1e-10: print(np.round((a - b)[0].reshape(n, n, n) / np.abs(b).max(), 4))
except Exception as ex:
    print("cache failed", ex)
