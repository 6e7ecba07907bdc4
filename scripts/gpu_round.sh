#!/bin/bash
# One GPU-box call: GPU tests, bench line, ncu launch list of one step, ncu --set full of the carrier RHS kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh <tag> [skip-tests]
tag=${1:-vX}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_$tag.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
  tail -15 gpurun_out/pytest_$tag.log
fi
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 3000 gpurun_out/bench_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:carrier_rhs|poisson_cell_rhs|poisson_face_rhs|level_kernel|ell_|distribute_kernel|gather_kernel' -c 700 --csv \
  --log-file gpurun_out/launches_$tag.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:carrier_rhs -c 2 -o gpurun_out/rhs_$tag -f \
  python scripts/profile_step.py --steps 2 > gpurun_out/ncu_rhs_$tag.log 2>&1
ncu -i gpurun_out/rhs_$tag.ncu-rep --page raw --csv > gpurun_out/rhs_${tag}_raw.csv 2>/dev/null
if [ -n "$RHS_ALT" ]; then  # the same capture with another production kernel variant
  PECS_B200_RHS_KERNEL=$RHS_ALT timeout 600 ncu --set full --clock-control none --import-source on -k regex:carrier_rhs -c 2 \
    -o gpurun_out/rhs_${tag}_alt$RHS_ALT -f python scripts/profile_step.py --steps 2 > gpurun_out/ncu_rhs_${tag}_alt.log 2>&1
  ncu -i gpurun_out/rhs_${tag}_alt$RHS_ALT.ncu-rep --page raw --csv > gpurun_out/rhs_${tag}_alt${RHS_ALT}_raw.csv 2>/dev/null
fi
ls -la gpurun_out | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
