#!/bin/bash
tag=${1:-r02i}
mkdir -p gpurun_out
t() { local n=$1; shift; env "$@" timeout 300 python scripts/tune_step.py 7 $n 2>&1 | tail -1; }
( t default
  t wtw8 PECS_B200_WARP_TILE_WARPS=8
  t wtw16 PECS_B200_WARP_TILE_WARPS=16
  t wtw8_sw2 PECS_B200_WARP_TILE_WARPS=8 PECS_B200_SOLVE_STAGES_WARP=2
  t loop2 PECS_B200_WARP_TILE_LOOP=2
  t loop4 PECS_B200_WARP_TILE_LOOP=4
  t wtw8_loop2 PECS_B200_WARP_TILE_WARPS=8 PECS_B200_WARP_TILE_LOOP=2
  t wtw8_loop4 PECS_B200_WARP_TILE_WARPS=8 PECS_B200_WARP_TILE_LOOP=4 ) | tee gpurun_out/tune_$tag.log
