#!/bin/bash
# Round 2, final GPU call: the state to be judged -- smoke(), full GPU suite, the default bench line (all keys)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -2 gpurun_out/pytest_final.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['traffic'], 'rhs', d['rhs_roofline']['frac'], d['rhs_roofline']['traffic'])
print('parity', {k: d['parity'][k] for k in ('ok', 'rhs_err', 'solve_density_err', 'solve_current_err') if k in d['parity']})
print('setup', d['config']['setup_seconds'], 'clocks', d['clocks'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('cfg1', d.get('cfg1_default_input'))
PY
tail -3 gpurun_out/bench_final.err
