#!/bin/bash
# Round 2, GPU call A: race reproducer A/B (coherent vs round-1 non-coherent loads), GPU suite, bench with the cfg3
# parity check, ncu launch list.   bash scripts/gpu_r02_a.sh <tag>
tag=${1:-r02a}
mkdir -p gpurun_out
nproc > gpurun_out/host_$tag.txt; free -g >> gpurun_out/host_$tag.txt; nvidia-smi -L >> gpurun_out/host_$tag.txt
rm -f gpurun_out/race_repro_$tag.jsonl gpurun_out/race_ref_*.npz
NC=$PWD/pecs_b200/lib/libpecs_b200_nc.so
run() {  # tag, env...
  local t=$1; shift
  env "$@" timeout 600 python scripts/race_repro.py --g ${G:-4} --reps ${REPS:-300} --steps 25 --tag $t \
      --out gpurun_out/race_repro_$tag.jsonl 2>&1 | tail -1
}
for G in 4 3; do
  export G
  run ok_pdl0_defer0 PECS_B200_PDL=0 PECS_B200_DEFER_CURRENTS=0     # first: writes the reference result
  run ok_default
  run ok_defer1 PECS_B200_DEFER_CURRENTS=1
  run ok_defer2 PECS_B200_DEFER_CURRENTS=2
  run nc_default PECS_B200_LIB=$NC
  run nc_defer1 PECS_B200_LIB=$NC PECS_B200_DEFER_CURRENTS=1
  run nc_defer2 PECS_B200_LIB=$NC PECS_B200_DEFER_CURRENTS=2
  run nc_defer1_pdl0 PECS_B200_LIB=$NC PECS_B200_DEFER_CURRENTS=1 PECS_B200_PDL=0
done
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 -s > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
grep -E "passed|failed|error|parity|species:|  [0-3]: " gpurun_out/pytest_$tag.log | tail -20
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 2500 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
PECS_B200_NO_SHARED_FACTORS=1 timeout 600 python bench.py --no-validate --no-cpu-baseline > gpurun_out/bench_${tag}_unshared.json 2> gpurun_out/bench_${tag}_unshared.err
python -c "import json;d=json.load(open('gpurun_out/bench_${tag}_unshared.json'));print('unshared', d['value'], d['ms_per_step'], d['roofline']['frac'])"
PECS_B200_DEFER_CURRENTS=1 timeout 600 python bench.py --no-validate --no-cpu-baseline > gpurun_out/bench_${tag}_defer1.json 2> gpurun_out/bench_${tag}_defer1.err
python -c "import json;d=json.load(open('gpurun_out/bench_${tag}_defer1.json'));print('defer1', d['value'], d['ms_per_step'], d['roofline']['frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:carrier_rhs|poisson_cell_rhs|poisson_face_rhs|level_kernel|ell_|distribute_kernel|gather_kernel' -c 700 --csv \
  --log-file gpurun_out/launches_$tag.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
ls -la gpurun_out | tail -5
