#!/bin/bash
# Round 2, GPU call Q: where the setup time goes (cfg3), then the final-state suite + bench line
tag=${1:-r02q}
mkdir -p gpurun_out
PECS_B200_SETUP_TIMING=1 timeout 300 python - > gpurun_out/setup_timing_$tag.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import pecs_b200 as pecs
t0 = time.perf_counter()
prob = pecs.SolarCellProblem(pecs.default_input_file(7, 1))
t1 = time.perf_counter()
prob.setup_full_system_host()
t2 = time.perf_counter()
print(f"host classes: create {t1 - t0:.2f} s, setup_full_system_host (grids, dofs, maps, matrices) {t2 - t1:.2f} s", flush=True)
prob2 = pecs.SolarCellProblem(pecs.default_input_file(7, 1))
t3 = time.perf_counter()
prob2.setup_full_system()
prob2.synchronize()
print(f"setup_full_system total {time.perf_counter() - t3:.2f} s", flush=True)
PY
cat gpurun_out/setup_timing_$tag.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -2 gpurun_out/pytest_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1]);print('bench', d['value'], d['e2e']['value'], d['parity']['ok'], d['roofline']['traffic'], d['rhs_roofline']['traffic'])"
