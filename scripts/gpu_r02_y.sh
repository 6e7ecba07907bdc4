#!/bin/bash
# Round 2, GPU call Y: where the numeric factorisation of a system goes, level by level (PECS_B200_SETUP_TIMING=2)
mkdir -p gpurun_out
PECS_B200_SETUP_TIMING=2 timeout 300 python - > gpurun_out/factor_levels.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import pecs_b200 as pecs
for rep in range(2):
    prob = pecs.SolarCellProblem(pecs.default_input_file(7, 1))
    t = time.perf_counter()
    prob.setup_full_system()
    prob.synchronize()
    print(f"setup_full_system total {time.perf_counter() - t:.2f} s (repetition {rep})", flush=True)
    prob.close()
PY
sed -n '/repetition 0/,$p' gpurun_out/factor_levels.log | grep -E "factorize_device|numeric fact|total" | cut -c1-130 | head -90
