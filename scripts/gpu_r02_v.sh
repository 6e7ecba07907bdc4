#!/bin/bash
# Round 2, GPU call V: factorisation work buffers sized once, parallel ELL tables -- A/B against the previous library on one box,
# then the GPU suite and the bench line
mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" PECS_B200_SETUP_TIMING=1 timeout 300 python - > gpurun_out/setup_ab2_$tag.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import pecs_b200 as pecs
for rep in range(4):
    prob = pecs.SolarCellProblem(pecs.default_input_file(7, 1))
    t = time.perf_counter()
    prob.setup_full_system()
    prob.synchronize()
    print(f"setup_full_system total {time.perf_counter() - t:.2f} s (repetition {rep})", flush=True)
    prob.step(3); prob.synchronize()
    prob.close()
PY
  echo "== $tag"; grep -E "total" gpurun_out/setup_ab2_$tag.log | cut -c1-120
}
run cur X=1
run prev PECS_B200_LIB=$PWD/pecs_b200/lib/libpecs_b200_prev.so
run cur_passive OMP_WAIT_POLICY=passive
run cur2 X=1
sed -n '/repetition 2/,$p' gpurun_out/setup_ab2_cur2.log | grep -E "numeric fact|index tables|system in all|wait for|semiconductor:|electrolyte:|Poisson:" | cut -c1-110
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02v.log 2>&1; tail -2 gpurun_out/pytest_r02v.log
timeout 600 python bench.py --no-cpu-baseline --no-cfg1 > gpurun_out/bench_r02v.json 2> gpurun_out/bench_r02v.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_r02v.json').read().strip().splitlines()[-1]);print('bench', d['value'], d['e2e']['value'], d['parity']['ok'], d['parity']['solve_density_err'], d['config']['setup_seconds'])"
