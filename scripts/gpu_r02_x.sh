#!/bin/bash
# Round 2, GPU call X: ncu launch list of the final kernels (regenerates profiles/r02_solve_traffic.json: the file is tied
# to a hash of the kernel sources), then the default bench line with it
tag=r02x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k 'regex:carrier_rhs|poisson_cell_rhs|poisson_face_rhs|level_kernel|ell_|distribute_kernel|gather_kernel' -c 700 --csv \
  --log-file gpurun_out/launches_$tag.csv python scripts/profile_step.py --steps 2 > gpurun_out/ncu_launches_$tag.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_$tag.csv --out gpurun_out/solve_traffic_$tag.json | tail -25
cp gpurun_out/solve_traffic_$tag.json profiles/r02_solve_traffic.json
timeout 900 python bench.py --no-cfg1 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1]);print('bench', d['value'], d['e2e']['value'], d['parity']['ok'], d['roofline']['frac'], d['roofline']['traffic'], d['config']['setup_seconds'], d['cpu_baseline']['value'])"
