"""GPU-box check of the sharded step (launch with torchrun, 2 or 4 ranks): the sharded steps must reproduce a
single-context run on rank 0 (densities, potential), then the step time is printed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/shard_check.py 5
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import pecs_b200 as pecs  # noqa: E402
from pecs_b200 import shard, sweep  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = 6
rank, local, world, dist = sweep.init_distributed("nccl")
torch.cuda.set_device(local)
prob = pecs.SolarCellProblem(pecs.default_input_file(g, 1), device=local)
prob.set_owned_species(shard.owned_mask(rank, world))
prob.setup_full_system()
engine = shard.GpuEngine(prob, local)
exchange = sys.argv[2] if len(sys.argv) > 2 else "nccl"
if exchange == "p2p":
    engine.connect_p2p(dist, rank, world)
stepper = shard.ShardedStepper(engine, dist, rank, world)
stepper.step(steps)
prob.synchronize()
full = shard.gather_states(prob, dist, rank, world, local)
if rank == 0:
    single = pecs.SolarCellProblem(pecs.default_input_file(g, 1), device=local)
    single.setup_full_system()
    single.step(steps)
    worst = 0.0
    for s in range(5):
        ref = single.get_solution(s)
        worst = max(worst, np.abs(full[s] - ref).max() / np.abs(ref).max())
    print(f"world {world} g {g} exchange {exchange}: sharded vs single-context after {steps} steps: max rel diff {worst:.3e}",
          flush=True)
    assert worst <= 1e-12
    single.close()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sweep.barrier(dist, local)
with torch.cuda.stream(engine.stream):
    a.record()
stepper.step(20)
with torch.cuda.stream(engine.stream):
    b.record()
b.synchronize()
ms = sweep.max_over_ranks(a.elapsed_time(b) / 20, dist, local)
if rank == 0:
    print(f"world {world} g {g} exchange {exchange}: {ms:.3f} ms/step sharded ({shard.mode_name(world)})", flush=True)
prob.close()
dist.destroy_process_group()
