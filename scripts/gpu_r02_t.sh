#!/bin/bash
# Round 2, GPU call T: threaded host setup (gather-form LDG assembly, Schur reduction, dissection) -- setup timing, parity of the factor tables (suite)
tag=${1:-r02t}
mkdir -p gpurun_out
PECS_B200_SETUP_TIMING=1 timeout 300 python - > gpurun_out/setup_timing_$tag.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import pecs_b200 as pecs
for rep in range(2):
    prob = pecs.SolarCellProblem(pecs.default_input_file(7, 1))
    t = time.perf_counter()
    prob.setup_full_system()
    prob.synchronize()
    print(f"setup_full_system total {time.perf_counter() - t:.2f} s (repetition {rep})", flush=True)
    prob.step(3); prob.synchronize()
    prob.close()
PY
grep -E "total|factorisation|wait for the host|semiconductor:|electrolyte:|Poisson:|host\)|setup_full_system:" gpurun_out/setup_timing_$tag.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -2 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --no-cpu-baseline --no-cfg1 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1]);print('bench', d['value'], d['e2e']['value'], d['parity']['ok'], d['parity']['solve_density_err'], d['config']['setup_seconds'])"
