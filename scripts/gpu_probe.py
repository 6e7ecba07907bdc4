"""Development probe (GPU box): setup time, factor size, step timing for a list of global refinement levels."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import pecs_b200 as pecs  # noqa: E402
from pecs_b200 import solarcell as sc  # noqa: E402

for g in [int(a) for a in sys.argv[1:]] or [4]:
    t0 = time.time()
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, 1))
    prob.setup_full_system_host()
    t1 = time.time()
    prob.setup_full_system()  # repeats the host part; fine for a probe
    prob.synchronize()
    t2 = time.time()
    fb = prob.info(sc.INFO_FACTOR_BYTES)
    print(f"g={g} cells={prob.n_cells(0)} host_setup={t1 - t0:.1f}s full_setup={t2 - t1:.1f}s factor={fb / 1e9:.2f} GB "
          f"launches/step={prob.info(sc.INFO_LAUNCHES_PER_STEP)} levels={prob.info(sc.INFO_TREE_LEVELS_MAX)}", flush=True)
    prob.step(3)
    prob.synchronize()
    K = 20
    ms = prob.step_timed(K)
    print(f"   graph: {ms[0] / K:.3f} ms/step -> {1000 * K / ms[0]:.1f} steps/s ; solve GB/s = {fb / (ms[0] / K * 1e-3) / 1e9:.0f}")
    ms = prob.step_timed(5, sectioned=True)
    names = ["total", "semi rhs", "elec rhs", "solve LDG", "Poisson rhs", "solve Poisson"]
    print("   sections (ms/step):", {n: round(v / 5, 4) for n, v in zip(names, ms)})
    for which, name in enumerate(["carrier rhs", "poisson rhs", "carrier solves", "poisson solve"]):
        t, launches = prob.time_kernel(which, 10)
        print(f"   {name}: {t:.4f} ms, {launches} launches")
    rhs_bytes = 368 * 2 * prob.n_cells(0)
    t, _ = prob.time_kernel(0, 10)
    print(f"   carrier rhs GB/s (L2 flushed) = {rhs_bytes / (t * 1e-3) / 1e9:.0f}")
    u = prob.get_solution(0)
    print("   electrons density range", u[8 * prob.n_cells(0):].min(), u[8 * prob.n_cells(0):].max(), "finite", np.isfinite(u).all())
    prob.close()
