"""Development sweep (GPU box): step time at cfg3 for a list of solve-kernel shapes given as ENV=VAL,ENV=VAL strings."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = """
import sys
sys.path.insert(0, %r)
import pecs_b200 as pecs
from pecs_b200 import solarcell as sc
prob = pecs.SolarCellProblem(pecs.default_input_file(int(sys.argv[1]), 1))
prob.setup_full_system()
prob.step(3); prob.synchronize()
K = 30
ms = prob.step_timed(K)
sec = prob.step_timed(5, sectioned=True)
fb = prob.info(sc.INFO_SOLVE_BYTES_PER_STEP)
print(f"{ms[0]/K:.3f} ms/step  {1000*K/ms[0]:.1f} steps/s  {fb/(ms[0]/K*1e-3)/1e9:.0f} GB/s | LDG {sec[3]/5:.3f} Poisson {sec[5]/5:.3f} | launches {prob.info(0)} bytes {fb/1e9:.2f} GB")
""" % ROOT
g = sys.argv[1]
for spec in sys.argv[2:]:
    env = dict(os.environ)
    if spec != "default":
        for kv in spec.split(","):
            k, v = kv.split("=")
            env["PECS_B200_" + k] = v
    r = subprocess.run([sys.executable, "-c", RUN, g], env=env, capture_output=True, text=True)
    print(f"{spec:60s} {r.stdout.strip() if r.returncode == 0 else 'FAILED ' + r.stderr[-300:]}", flush=True)
