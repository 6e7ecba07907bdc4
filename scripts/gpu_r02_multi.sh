#!/bin/bash
# Round 2, multi-GPU call: bash scripts/gpu_r02_multi.sh <n_gpus> <tag>
n=${1:-2}; tag=${2:-r02m}; g=${3:-5}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${tag}_$n.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/shard_check.py $g p2p 2>&1 | grep -E "world|Error|error" | tail -3
timeout 600 $TR scripts/shard_check.py $g nccl 2>&1 | grep -E "world|Error|error" | tail -3
timeout 1200 $TR bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_${tag}_$n.json 2> gpurun_out/bench_${tag}_$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}_$n.json").read().strip().splitlines()[-1])
print('sweep: steps/s %.1f ms %.4f e2e %.1f'%(d['value'], d['ms_per_step'], d['e2e']['value']))
print('sharded', d.get('sharded'))
PY
tail -3 gpurun_out/bench_${tag}_$n.err
timeout 900 $TR bench.py --gpus $n --steps 50 --warmup 5 --parallelism $([ $n -eq 2 ] && echo subdomain || echo species) --exchange nccl > gpurun_out/bench_${tag}_${n}_nccl.json 2> gpurun_out/bench_${tag}_${n}_nccl.err
python -c "import json;d=json.load(open('gpurun_out/bench_${tag}_${n}_nccl.json'));print('sharded nccl: steps/s %.1f'%d['value'])"
