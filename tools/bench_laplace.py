#!/usr/bin/env python
"""Device-resident LaplaceOperator vmult rate (heat_transfer.h:170-250) on a Cartesian cube: python tools/bench_laplace.py --degree 4 --cells 64"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "matrix-free-dealii-precice_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from b200mf import DeviceVector, LaplaceOperator, MatrixFree
from b200mf.mesh import StructuredMesh, smooth_mapping

ap = argparse.ArgumentParser()
ap.add_argument("--degree", type=int, default=4)
ap.add_argument("--cells", type=int, default=64)
ap.add_argument("--mapped", action="store_true", help="smoothly deformed (non-affine) cells: general-metric kernel")
ap.add_argument("--tag", default="")
a = ap.parse_args()
m = StructuredMesh((a.cells,) * 3, a.degree, mapping=smooth_mapping(0.05) if a.mapped else None)
m.clamp([(d, s) for d in range(3) for s in range(2)], 1)
out = {"tag": a.tag, "degree": a.degree, "cells": a.cells, "dofs": m.n_nodes, "mapped": a.mapped}
for dt in ("f64", "f32"):
    mf = MatrixFree.from_mesh(m, 1, dtype=dt)
    op = LaplaceOperator(mf, 0.1)
    x = DeviceVector(mf, np.random.default_rng(0).uniform(-1, 1, m.n_nodes)); y = DeviceVector(mf)
    for _ in range(3):
        op.vmult(y, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        op.vmult(y, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[dt] = {"ms": ms, "dofs_per_s": m.n_nodes / ms * 1e3, "affine": mf.all_cells_affine()}
    del op, x, y, mf
print(json.dumps(out), flush=True)
