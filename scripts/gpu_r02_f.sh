#!/bin/bash
# Round 2, GPU call F: tile-size tuning of the level kernels; SRH parity
tag=${1:-r02f}
mkdir -p gpurun_out
t() { local n=$1; shift; env "$@" timeout 300 python scripts/tune_step.py 7 $n 2>&1 | tail -1; }
( t default
  t tiles1776 PECS_B200_TILE_TARGET=1776
  t tiles1184 PECS_B200_TILE_TARGET=1184
  t tiles888 PECS_B200_TILE_TARGET=888
  t tiles592 PECS_B200_TILE_TARGET=592
  t tiles1184_st3 PECS_B200_TILE_TARGET=1184 PECS_B200_SOLVE_STAGES=3
  t tiles1184_st4 PECS_B200_TILE_TARGET=1184 PECS_B200_SOLVE_STAGES=4
  t tiles1184_sw3 PECS_B200_TILE_TARGET=1184 PECS_B200_SOLVE_STAGES_WARP=3
  t tiles888_sw3 PECS_B200_TILE_TARGET=888 PECS_B200_SOLVE_STAGES_WARP=3
  t tiles1184_defer PECS_B200_TILE_TARGET=1184 PECS_B200_DEFER_CURRENTS=1
  t tiles1184_warps8 PECS_B200_TILE_TARGET=1184 PECS_B200_SOLVE_WARPS=8
  t tiles1184_dataflow0 PECS_B200_TILE_TARGET=1184 PECS_B200_DATAFLOW=0 ) | tee gpurun_out/tune_$tag.log
timeout 600 python -m pytest tests/test_gpu_extra.py -m gpu -x -q -k "other_configurations or shared" > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
