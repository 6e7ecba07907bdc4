#!/bin/bash
# Round 2, GPU call O: final-state check -- full GPU suite and the race reproducer on the final kernels
tag=${1:-r02o}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -12 gpurun_out/pytest_$tag.log | cut -c1-300
rm -f gpurun_out/race_repro_$tag.jsonl gpurun_out/race_ref_*.npz
run() { local t=$1; shift; env "$@" timeout 600 python scripts/race_repro.py --g ${G:-4} --reps ${REPS:-200} --steps 25 --tag $t --out gpurun_out/race_repro_$tag.jsonl 2>&1 | tail -1 | cut -c1-400; }
export G=4 REPS=200
run ref_pdl0_dataflow0 PECS_B200_PDL=0 PECS_B200_DATAFLOW=0
run default
run defer1 PECS_B200_DEFER_CURRENTS=1
run dataflow0 PECS_B200_DATAFLOW=0
export G=6 REPS=30
run ref_pdl0_dataflow0 PECS_B200_PDL=0 PECS_B200_DATAFLOW=0
run default
run defer1 PECS_B200_DEFER_CURRENTS=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
