#!/bin/bash
# Round 2, GPU call W: device warm-up overlapped with the host setup, host preparation before the first device call, lane handles in parallel -- first-context setup time, GPU suite, bench line
mkdir -p gpurun_out
for p in 1 2; do
PECS_B200_SETUP_TIMING=1 timeout 300 python - > gpurun_out/setup_first_$p.log 2>&1 <<'PY'
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
grep -E "total|set_solvers|wait for the host" gpurun_out/setup_first_$p.log | cut -c1-110
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02w.log 2>&1; tail -2 gpurun_out/pytest_r02w.log
timeout 600 python bench.py --no-cpu-baseline --no-cfg1 > gpurun_out/bench_r02w.json 2> gpurun_out/bench_r02w.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_r02w.json').read().strip().splitlines()[-1]);print('bench', d['value'], d['e2e']['value'], d['parity']['ok'], d['parity']['solve_density_err'], d['config']['setup_seconds'])"
