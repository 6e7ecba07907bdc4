#!/bin/bash
# Round 2, last GPU call: ELL tables from the preparation threads -- setup timing, full GPU suite, default bench line
mkdir -p gpurun_out
PECS_B200_SETUP_TIMING=1 timeout 300 python - > gpurun_out/setup_timing_final3.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import pecs_b200 as pecs
for rep in range(3):
    prob = pecs.SolarCellProblem(pecs.default_input_file(7, 1))
    t = time.perf_counter()
    prob.setup_full_system()
    prob.synchronize()
    print(f"setup_full_system total {time.perf_counter() - t:.2f} s (repetition {rep})", flush=True)
    prob.step(3); prob.synchronize()
    prob.close()
PY
grep -E "total|wait for the host|system in all|semiconductor:|electrolyte:|Poisson:" gpurun_out/setup_timing_final3.log | cut -c1-110 | tail -12
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final3.log 2>&1; tail -2 gpurun_out/pytest_final3.log
timeout 900 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_final3.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e']['value'], d['parity']['ok'], d['parity']['solve_density_err'], d['config']['setup_seconds'], d['roofline']['traffic'], d['workload_run_full_system']['run_full_system_wall_seconds'])"
