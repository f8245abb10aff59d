#!/bin/bash
# round 2, GPU call U: MG graph replay tests, boundary mass operator / projector / boundary flux, Newton + heat steps
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_solver.py tests/test_gpu_heat.py tests/test_gpu_newton.py tests/test_gpu_faces.py -q 2>&1 | tail -40
