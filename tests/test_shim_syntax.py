"""The deal.II-side shim (contrib/dealii/b200mf_dealii.h) cannot be compiled against deal.II in this image; this test at
least compiles it - every template instantiated, every member used - against tests/mock_dealii/, a stand-in that declares
the deal.II classes and members the shim uses with deal.II's documented signatures.  Catches syntax and type errors and
calls of ABI functions with the wrong arguments (include/b200mf.h is the real header)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_shim_compiles_against_mock_dealii():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter",
           "-I", os.path.join(ROOT, "tests", "mock_dealii"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "contrib", "dealii"), os.path.join(ROOT, "tests", "mock_dealii", "instantiate_shim.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
