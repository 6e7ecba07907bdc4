#!/bin/bash
# Round 2, GPU call U: A/B of the setup path on ONE box -- previous library vs current, and OpenMP wait policies
mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" PECS_B200_SETUP_TIMING=1 timeout 300 python - > gpurun_out/setup_ab_$tag.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import pecs_b200 as pecs
for rep in range(3):
    prob = pecs.SolarCellProblem(pecs.default_input_file(7, 1))
    t = time.perf_counter()
    prob.setup_full_system()
    prob.synchronize()
    print(f"setup_full_system total {time.perf_counter() - t:.2f} s (repetition {rep})", flush=True)
    prob.step(3); prob.synchronize()
    prob.close()
PY
  echo "== $tag"; grep -E "total|numeric factorisation|index tables|wait for the host" gpurun_out/setup_ab_$tag.log | cut -c1-120 | tail -14
}
run prev PECS_B200_LIB=$PWD/pecs_b200/lib/libpecs_b200_prev.so
run cur X=1
run cur_passive OMP_WAIT_POLICY=passive
run cur_threads1 PECS_B200_SETUP_THREADS=1
run prev2 PECS_B200_LIB=$PWD/pecs_b200/lib/libpecs_b200_prev.so
