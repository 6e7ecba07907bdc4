#!/bin/bash
# Round 2, final 2-GPU check on the final sources: sharded step bit-identical to the single context (g=5), then the
# 2-GPU bench line (sweep + sharded)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 60 $TR scripts/shard_check.py 5 p2p 2>&1 | grep -E "world|Error|error" | tail -3
timeout 100 $TR bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_final_2gpu.json 2> gpurun_out/bench_final_2gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final_2gpu.json").read().strip().splitlines()[-1])
print('sweep: steps/s %.1f ms %.4f e2e %.1f'%(d['value'], d['ms_per_step'], d['e2e']['value']))
print('sharded', {k: d['sharded'][k] for k in ('steps_per_s', 'speedup_vs_one_gpu', 'setup_seconds') if k in d.get('sharded', {})})
PY
tail -2 gpurun_out/bench_final_2gpu.err
