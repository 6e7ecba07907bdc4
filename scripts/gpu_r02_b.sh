#!/bin/bash
# Round 2, GPU call B: dataflow level kernels + shared factors.  bash scripts/gpu_r02_b.sh <tag>
tag=${1:-r02b}
mkdir -p gpurun_out
rm -f gpurun_out/race_repro_$tag.jsonl gpurun_out/race_ref_*.npz
run() {  # tag, env...
  local t=$1; shift
  env "$@" timeout 600 python scripts/race_repro.py --g ${G:-4} --reps ${REPS:-200} --steps 25 --tag $t \
      --out gpurun_out/race_repro_$tag.jsonl 2>&1 | tail -1
}
export G=4
run ref_pdl0_dataflow0 PECS_B200_PDL=0 PECS_B200_DATAFLOW=0 PECS_B200_DEFER_CURRENTS=0
run default
run dataflow0 PECS_B200_DATAFLOW=0
run defer1 PECS_B200_DEFER_CURRENTS=1
run legacyset PECS_B200_LEGACY_SET_STATE=1
run legacyset_defer1 PECS_B200_LEGACY_SET_STATE=1 PECS_B200_DEFER_CURRENTS=1
run unshared PECS_B200_NO_SHARED_FACTORS=1
export G=5 REPS=60
run ref_pdl0_dataflow0 PECS_B200_PDL=0 PECS_B200_DATAFLOW=0
run default
run defer1 PECS_B200_DEFER_CURRENTS=1
timeout 1800 python -m pytest tests -m gpu -x -q --durations=10 -s > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
grep -E "passed|failed|error|parity|species:|  [0-3]: |Error" gpurun_out/pytest_$tag.log | tail -20
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
b unshared_dataflow0 PECS_B200_NO_SHARED_FACTORS=1 PECS_B200_DATAFLOW=0
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:carrier_rhs|poisson_cell_rhs|poisson_face_rhs|level_kernel|ell_|distribute_kernel|gather_kernel' -c 700 --csv \
  --log-file gpurun_out/launches_$tag.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
ls -la gpurun_out | tail -4
