#!/usr/bin/env python
"""The `newton_step` block of bench.py alone (for an ncu launch list): prints the number of kernels of the measured step, which are
the LAST kernels of the process.  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/newton_step_profile.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    out = bench.newton_block(stream.cuda_stream)
torch.cuda.synchronize()
print(json.dumps({k: out[k] for k in ("newton_step_s", "update_matrix_free_s", "cg_solve_s", "cg_iterations", "gpu_launches", "mg_graph")}))
os._exit(0)  # no teardown kernels after the measured step
