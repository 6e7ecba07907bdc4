#!/bin/bash
# Round 2, GPU call D: resident-wave level kernels.  bash scripts/gpu_r02_d.sh <tag>
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 -s > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
grep -E "passed|failed|error|parity|Error" gpurun_out/pytest_$tag.log | tail -8 | cut -c1-600
b() {  # name, env...
  local n=$1; shift
  env "$@" timeout 900 python bench.py --no-cpu-baseline ${VAL:---no-validate} > gpurun_out/bench_${tag}_$n.json 2> gpurun_out/bench_${tag}_$n.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${tag}_$n.json')); r=d['roofline']
    print('$n', 'steps/s %.1f ms %.4f e2e %.1f solve bytes %.2f GB frac %.3f poisson_ms %.4f ldg_ms %.4f'%(d['value'], d['ms_per_step'], d['e2e']['value'], r['algorithmic_bytes_per_step']/1e9, r['frac'], d['section_ms_per_step']['Solve Poisson system'], d['section_ms_per_step']['Solve LDG Systems']))
    if d.get('parity'): print('   parity', {k:v for k,v in d['parity'].items() if not isinstance(v,(dict,str))})
except Exception as e: print('$n failed', e)
PY
}
VAL=" " b default
b dataflow0 PECS_B200_DATAFLOW=0
b defer1 PECS_B200_DEFER_CURRENTS=1
b unshared PECS_B200_NO_SHARED_FACTORS=1
T=$PWD/pecs_b200/lib/libpecs_b200_trace.so
PECS_B200_LIB=$T timeout 600 python scripts/trace_step.py -g 7 --tag ${tag}_dataflow > gpurun_out/trace_${tag}_dataflow.log 2>&1
grep "==" gpurun_out/trace_${tag}_*.log
python - <<'PY'
import sys
sys.path.insert(0, ".")
import pecs_b200 as pecs
prob = pecs.SolarCellProblem(pecs.default_input_file(7, 1)); prob.setup_full_system(); prob.step(3)
print("time_kernel: carrier solves %.4f ms (%d launches); poisson solve %.4f ms (%d launches)" % (*prob.time_kernel(2, 10), *prob.time_kernel(3, 10)))
print("wait errors", prob.info(7))
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
