#!/bin/bash
# Round 2, GPU call C: per-block timestamps of the level kernels (trace build), dataflow on / off
tag=${1:-r02c}
mkdir -p gpurun_out
T=$PWD/pecs_b200/lib/libpecs_b200_trace.so
PECS_B200_LIB=$T timeout 600 python scripts/trace_step.py -g 7 --tag ${tag}_dataflow > gpurun_out/trace_${tag}_dataflow.log 2>&1
PECS_B200_LIB=$T PECS_B200_DATAFLOW=0 timeout 600 python scripts/trace_step.py -g 7 --tag ${tag}_gridwait > gpurun_out/trace_${tag}_gridwait.log 2>&1
PECS_B200_LIB=$T PECS_B200_PDL=0 timeout 600 python scripts/trace_step.py -g 7 --tag ${tag}_nopdl > gpurun_out/trace_${tag}_nopdl.log 2>&1
grep "==" gpurun_out/trace_${tag}_*.log
python - <<'PY'
import sys
sys.path.insert(0, ".")
import pecs_b200 as pecs
prob = pecs.SolarCellProblem(pecs.default_input_file(7, 1)); prob.setup_full_system(); prob.step(3)
print("time_kernel: carrier solves %.4f ms (%d launches); poisson solve %.4f ms (%d launches)" % (*prob.time_kernel(2, 10), *prob.time_kernel(3, 10)))
PY
ls -la gpurun_out | tail -4
