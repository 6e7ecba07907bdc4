"""bench.py's CPU arm (oracle/cpu_arm.py): oracle assembly + SuperLU as the UMFPACK stand-in, on the oracle's own grid.
CPU only.  (a) its states equal those of the oracle with its own LU -- two direct solvers, one system; (b) the arm
never loads the product library; (c) the JSON line carries what the bench contract asks for."""
import json
import os
import subprocess
import sys

import numpy as np

from helpers import rel_err
from oracle import cpu_arm, grid as ogrid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_superlu_path_equals_oracle_lu():
    path = cpu_arm.CpuReferencePath(3, 1, threads=2).setup()
    path.step(5)
    got = path.states()
    path.close()
    o = ogrid.make_oracle({"global refinements": 3, "local refinements": 1})
    o.setup(1.0, True)
    o.project_initial_conditions()
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    o.step(5)
    for s in range(5):
        want = o.solution(s)
        n = want.size
        if s < 4:
            assert rel_err(got[s][8 * (n // 12):], want[8 * (n // 12):]) <= 1e-10
            assert rel_err(got[s], want) <= 1e-7
        else:
            assert rel_err(got[s], want) <= 1e-10


def test_reference_arm_line_and_independence():
    code = ("import sys; sys.argv = ['bench.py', '--impl', 'reference', '--global-refinements', '4', '--steps', '3', "
            "'--warmup', '1']; sys.path.insert(0, %r); import bench; bench.main(); "
            "bad = [m for m in sys.modules if m.startswith('pecs_b200')]; "
            "maps = open('/proc/self/maps').read(); "
            "assert not bad and 'libpecs_b200' not in maps, (bad, 'product library loaded by the CPU arm')" % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0")  # what torchrun would set: the arm must not inherit it
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "steps/s" and line["value"] > 0
    assert line["config"]["same_config"] is True and line["cpu_baseline"]["same_config"] is True
    assert line["cpu_baseline"]["cores"] == os.cpu_count() and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert len(line["cpu_baseline"]["samples"]) == 2


def test_other_ranks_of_the_reference_arm_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""
