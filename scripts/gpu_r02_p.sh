#!/bin/bash
# Round 2, GPU call P: staggered solves A/B for the end-to-end path, the final bench line and launch list
tag=${1:-r02p}
mkdir -p gpurun_out
b() {  # name, flags..., env via B_ENV
  local n=$1; shift
  env $B_ENV timeout 900 python bench.py "$@" > gpurun_out/bench_${tag}_$n.json 2> gpurun_out/bench_${tag}_$n.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${tag}_$n.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$n', 'steps/s %.1f ms %.4f e2e %.1f frac %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac']))
    if d.get('parity'): print('   parity ok', d['parity']['ok'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'cfg1', d.get('cfg1_default_input',{}).get('run_full_system_wall_seconds'))
except Exception as e: print('$n failed', e)
PY
}
B_ENV="PECS_B200_STAGGER=1" b stagger1 --no-cpu-baseline --no-validate --no-cfg1
B_ENV="PECS_B200_STAGGER=1 PECS_B200_DEFER_CURRENTS=1" b stagger1_defer1 --no-cpu-baseline --no-validate --no-cfg1
B_ENV="PECS_B200_DEFER_CURRENTS=1" b defer1 --no-cpu-baseline --no-validate --no-cfg1
B_ENV="" b default
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:carrier_rhs|poisson_cell_rhs|poisson_face_rhs|level_kernel|ell_|distribute_kernel|interface_current' -c 700 --csv \
  --log-file gpurun_out/launches_$tag.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 900 python -m pytest tests/test_output_path.py tests/test_bench_line.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
