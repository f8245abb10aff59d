"""bench.py contract pieces that need no GPU: the algorithmic-bytes table of SURVEY.md §8d and the CPU arm's JSON line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_algorithmic_bytes_table():
    import bench

    # SURVEY.md §8d, "tensor2-sym" row (tau/J stored symmetric: 64 + 80 B per q-point): B/DoF for p = 1..6
    expect = [410.7, 182.5, 132.9, 112.4, 101.2, 94.3]
    for p, e in zip(range(1, 7), expect):
        assert bench.algorithmic_bytes_per_cell(p) / (3 * p**3) == pytest.approx(e, abs=0.06)
    # one launch at the bench workload: 85^3 cells x 10 768 B
    assert bench.algorithmic_bytes_per_cell(3) == 10768
    assert bench.algorithmic_bytes_per_cell(3, 4) == 2 * 4 * 81 + 4 * 64 + 18 * 4 * 64  # FP32 level operator
    # the fatter levels are scored against their own rows
    assert bench.algorithmic_bytes_per_cell(3, 8, "tensor4_ns") > bench.algorithmic_bytes_per_cell(3, 8, "tensor4") > 10768


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-cells", "6"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "nh_tangent_vmult_dofs_per_s" and d["unit"] == "DoFs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("neo-Hookean tangent vmult")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
