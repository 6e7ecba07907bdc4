"""The bench contract on a small mesh: `python bench.py` prints ONE JSON line with the keys the driver reads, the parity
check of the workload passes, and the numbers are consistent with each other.  (The real run is cfg3; this guards the
code path.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_bench_line_small_mesh():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--global-refinements", "5", "--steps", "10",
                        "--warmup", "3", "--cpu-steps", "2", "--no-cfg1"], capture_output=True, text=True, timeout=900,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity"):
        assert key in d, key
    assert d["unit"] == "steps/s" and d["n_gpus"] == 1 and d["steps"] == 10 and d["dtype"] == "f64"
    assert abs(d["value"] - 1000.0 / d["ms_per_step"]) <= 1e-6 * d["value"]
    assert d["gpu_launches"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["value"] < d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    rl = d["roofline"]
    assert rl["bound"] == "hbm" and 0 < rl["frac"] < 1 and abs(rl["frac"] - rl["achieved"] / rl["peak"]) < 1e-9
    assert d["parity"]["ok"] is True and d["parity"]["rhs_rel"] <= 1e-12 and d["parity"]["solve_wait_errors"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] > 0 and len(cb["measured"]) == 2


def test_bench_help_needs_no_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "--impl" in r.stdout
