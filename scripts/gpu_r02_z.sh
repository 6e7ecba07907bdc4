#!/bin/bash
# Round 2, GPU call Z: tiled products of the small fronts in the setup factorisation -- per-level timing, GPU suite, bench line
# (parity.solve_density_err of the bench is a fingerprint of the factor tables: it must not move)
mkdir -p gpurun_out
bash scripts/gpu_r02_y.sh | grep -E "level  [5-9]|level 1[0-9]|numeric fact|total" | head -48
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02z.log 2>&1; tail -2 gpurun_out/pytest_r02z.log
timeout 600 python bench.py --no-cpu-baseline --no-cfg1 > gpurun_out/bench_r02z.json 2> gpurun_out/bench_r02z.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_r02z.json').read().strip().splitlines()[-1]);print('bench', d['value'], d['e2e']['value'], d['parity']['ok'], d['parity']['solve_density_err'], d['config']['setup_seconds'], d['roofline']['traffic'])"
