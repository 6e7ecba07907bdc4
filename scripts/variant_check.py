"""Development check (GPU box): the solver variants selected by environment variables must agree."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANT = """
import sys, numpy as np
sys.path.insert(0, %r)
import pecs_b200 as pecs
prob = pecs.SolarCellProblem(pecs.default_input_file(int(sys.argv[2]), 1))
prob.setup_full_system()
prob.step(4)
np.save(sys.argv[1], np.concatenate([prob.get_solution(s) for s in range(5)]))
""" % ROOT
g = sys.argv[1] if len(sys.argv) > 1 else "3"
envs = [{}, {"PECS_B200_NO_SCHUR": "1"}, {"PECS_B200_HOST_FACTOR": "1"}, {"PECS_B200_LEAF_NODES": "3"},
        {"PECS_B200_CELL_SEPARATORS": "1"}, {"PECS_B200_SCHUR_DROP": "0"}, {"PECS_B200_POISSON_CELL_NODES": "1"},
        {"PECS_B200_SOLVE_STAGES": "2"}, {"PECS_B200_SOLVE_PANELS_PER_TILE": "3"}, {"PECS_B200_SOLVE_STAGES": "7"}]
with tempfile.TemporaryDirectory() as tmp:
    script = os.path.join(tmp, "v.py")
    open(script, "w").write(VARIANT)
    ref = None
    for k, env in enumerate(envs):
        out = os.path.join(tmp, f"o{k}.npy")
        r = subprocess.run([sys.executable, script, out, g], env=dict(os.environ, **env), capture_output=True, text=True)
        if r.returncode != 0:
            print(env, "FAILED", r.stderr[-500:])
            continue
        x = np.load(out)
        if ref is None:
            ref = x
        print(env, "max rel diff vs default", np.abs(x - ref).max() / np.abs(ref).max(), flush=True)
