#!/usr/bin/env python
"""Extracts the reference's own golden numbers for the solid path into tests/golden/turek_solid.json.

Source: /root/reference/tests/solid/<case>/<case>_{serial,parallel}.output (the "blessed" files the reference's
tests/solid/run_tests.sh:78,96 compares against with numdiff) and the matching .prm.  Run in the build container
(where /root/reference exists); the JSON travels with the repo."""
import json
import os
import re

REF = "/root/reference/tests/solid"
CASES = ["turek_gmg_scalar_referential", "turek_gmg_tensor2", "turek_gmg_tensor4", "turek_gmg_tensor4_ns",
         "turek_gmg_tensor2_form0", "turek_jacobi_tensor2"]


def prm_value(txt, key):
    m = re.search(r"set\s+" + re.escape(key) + r"\s*=\s*(\S+)", txt)
    return m.group(1) if m else None


out = {}
for c in CASES:
    serial = open(os.path.join(REF, c, c + "_serial.output")).read()
    parallel = open(os.path.join(REF, c, c + "_parallel.output")).read()
    assert serial == parallel, c  # the reference's goldens are identical for 1 and 4 ranks
    prm = open(os.path.join(REF, c, c + ".prm")).read()
    disp = [[float(a), float(b)] for a, b in re.findall(r"displacement:\s+(\S+)\s+(\S+)", serial)]
    out[c] = {
        "source": f"tests/solid/{c}/{c}_serial.output",
        "reference_volume": float(re.search(r"Reference volume:\s+(\S+)", serial).group(1)),
        "n_cells": int(re.search(r"active cells:\s+(\d+)", serial).group(1)),
        "n_dofs": int(re.search(r"degrees of freedom:\s+(\d+)", serial).group(1)),
        "newton_iterations": [int(x) for x in re.findall(r"Converged in (\d+) Newton", serial)],
        "displacement": disp,
        "average_cg_iterations": int(re.search(r"Average CG iter = (\d+)", serial).group(1)),
        "prm": {"degree": int(prm_value(prm, "Polynomial degree")), "quadrature": int(prm_value(prm, "Quadrature order")),
                "refinement": int(prm_value(prm, "Global refinement")), "caching": prm_value(prm, "MF caching"),
                "preconditioner": prm_value(prm, "Preconditioner type"), "relaxation": float(prm_value(prm, "Preconditioner relaxation")),
                "formulation": int(prm_value(prm, "Formulation")), "mu": float(prm_value(prm, "Shear modulus")),
                "nu": float(prm_value(prm, "Poisson's ratio")), "rho": float(prm_value(prm, "Density")),
                "tol_lin": float(prm_value(prm, "Residual")), "delta_t": float(prm_value(prm, "Time step size")),
                "end_time": float(prm_value(prm, "End time")), "n_cheb_cg": int(prm_value(prm, "MF Chebyshev number CG iterations"))},
    }
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "turek_solid.json"), "w"), indent=1)
print(json.dumps({k: (v["displacement"], v["average_cg_iterations"]) for k, v in out.items()}, indent=1))
