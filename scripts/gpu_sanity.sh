#!/bin/bash
# Short GPU check: GPU tests, smoke, and one ncu --set full capture of the level kernels of a step (dominant kernels).
tag=${1:-vX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
# launches 0..75 are the initial Poisson solve (38) and the first kernels of step 1; take 24 level-kernel launches from the
# four concurrent carrier solves of step 2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:level_kernel --launch-skip 260 -c 24 \
  -o gpurun_out/levels_$tag -f python scripts/profile_step.py --steps 2 > gpurun_out/ncu_levels_$tag.log 2>&1
ncu -i gpurun_out/levels_$tag.ncu-rep --page raw --csv > gpurun_out/levels_${tag}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -4
