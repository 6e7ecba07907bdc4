#!/bin/bash
# Round 2, GPU call H: the state to be judged -- full GPU suite, the default bench line, ncu launch list + full captures
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 -s > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
grep -E "passed|failed|error|parity|Error|  [0-3]: " gpurun_out/pytest_$tag.log | tail -12 | cut -c1-500
timeout 1500 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json')); r=d['roofline']
print('bench: steps/s %.1f ms %.4f e2e %.1f solve bytes %.2f GB frac %.3f rhs frac %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], r['algorithmic_bytes_per_step']/1e9, r['frac'], d['rhs_roofline']['frac']))
print('parity', {k:v for k,v in d['parity'].items() if not isinstance(v,(dict,str))})
print('cpu', d['cpu_baseline'])
print('cfg1', d.get('cfg1_default_input'))
PY
tail -3 gpurun_out/bench_$tag.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:carrier_rhs|poisson_cell_rhs|poisson_face_rhs|level_kernel|ell_|distribute_kernel|gather_kernel' -c 700 --csv \
  --log-file gpurun_out/launches_$tag.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches_$tag.log 2>&1
# full captures: the carrier RHS kernel; the ELL kernels; 50 level kernels of the second step (carriers + Poisson)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:carrier_rhs -c 2 -o gpurun_out/rhs_$tag -f \
  python scripts/profile_step.py --steps 2 > gpurun_out/ncu_rhs_$tag.log 2>&1
ncu -i gpurun_out/rhs_$tag.ncu-rep --page raw --csv > gpurun_out/rhs_${tag}_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:level_kernel --launch-skip 160 -c 60 -o gpurun_out/levels_$tag -f \
  python scripts/profile_step.py --steps 2 > gpurun_out/ncu_levels_$tag.log 2>&1
ncu -i gpurun_out/levels_$tag.ncu-rep --page raw --csv > gpurun_out/levels_${tag}_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:level_kernel|ell_combine|distribute|poisson_cell' --launch-skip 290 -c 45 -o gpurun_out/poisson_$tag -f \
  python scripts/profile_step.py --steps 2 > gpurun_out/ncu_poisson_$tag.log 2>&1
ncu -i gpurun_out/poisson_$tag.ncu-rep --page raw --csv > gpurun_out/poisson_${tag}_raw.csv 2>/dev/null
rm -f gpurun_out/*_$tag.ncu-rep
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
ls -la gpurun_out | grep $tag
