"""GPU box: ms per step of the cfg workload under the current environment switches (tuning A/B; never a bench value).
    PECS_B200_...=... python scripts/tune_step.py [g] [tag]"""
import os
import sys

sys.path.insert(0, ".")
import pecs_b200 as pecs  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 7
tag = sys.argv[2] if len(sys.argv) > 2 else "default"
prob = pecs.SolarCellProblem(pecs.default_input_file(g, 1))
prob.setup_full_system()
prob.step(5)
prob.synchronize()
ms = [prob.step_timed(40)[0] / 40 for _ in range(3)]
solves = prob.step_timed(40, sectioned=2)[0] / 40
lone = []
import ctypes  # noqa: E402
for which in (0, 2, 4):  # lone electron solve, lone (shared) reductant solve, Poisson solve: wall time around 20 calls
    import time
    prob.synchronize()
    t = time.perf_counter()
    for _ in range(20):
        prob.solve_Poisson() if which == 4 else prob.solve_species(which)
    prob.synchronize()
    lone.append((time.perf_counter() - t) / 20 * 1e3)
env = {k: v for k, v in os.environ.items() if k.startswith("PECS_B200_") and k != "PECS_B200_LIB"}
print(f"{tag}: step {min(ms):.4f} ms ({1e3 / min(ms):.1f} steps/s), five solves only {solves:.4f} ms, lone electron solve "
      f"{lone[0]:.4f} ms, lone reductant solve {lone[1]:.4f} ms, Poisson solve {lone[2]:.4f} ms, wait errors {prob.info(7)} {env}", flush=True)
prob.close()
