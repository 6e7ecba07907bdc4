#!/bin/bash
# Round 2, GPU call E: tuning switches of the level kernels + the CPU reference arm on the workload itself
tag=${1:-r02e}
mkdir -p gpurun_out
t() { local n=$1; shift; env "$@" timeout 300 python scripts/tune_step.py 7 $n 2>&1 | tail -1; }
( t default
  t stages_warp2 PECS_B200_SOLVE_STAGES_WARP=2
  t stages_warp3 PECS_B200_SOLVE_STAGES_WARP=3
  t ppt8 PECS_B200_SOLVE_PANELS_PER_TILE=8
  t ppt8_sw2 PECS_B200_SOLVE_PANELS_PER_TILE=8 PECS_B200_SOLVE_STAGES_WARP=2
  t waves2 PECS_B200_LEVEL_WAVES=2
  t inflight256 PECS_B200_INFLIGHT_KB=256
  t defer1 PECS_B200_DEFER_CURRENTS=1
  t dataflow0 PECS_B200_DATAFLOW=0
  t unshared PECS_B200_NO_SHARED_FACTORS=1 ) | tee gpurun_out/tune_$tag.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extra.py tests/test_output_path.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
T=$PWD/pecs_b200/lib/libpecs_b200_trace.so
PECS_B200_LIB=$T timeout 600 python scripts/trace_step.py -g 7 --tag ${tag}_dataflow > gpurun_out/trace_${tag}_dataflow.log 2>&1
grep "==" gpurun_out/trace_${tag}_*.log
timeout 1700 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err
tail -c 2500 gpurun_out/bench_${tag}_reference.json; tail -3 gpurun_out/bench_${tag}_reference.err
