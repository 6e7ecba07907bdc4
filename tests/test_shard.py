"""The sharded step (pecs_b200/shard.py) on CPU: the same driver that runs over NCCL on GPUs, here over a world_size-2
gloo group with an engine built on the oracle.  Checks ownership, what is exchanged, and that the sharded steps
reproduce the single-process steps exactly (densities and potential; currents stay with their owner)."""
import os
import subprocess
import sys
import textwrap

import pytest

from pecs_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ownership():
    assert [shard.owner_of(s, 1) for s in range(4)] == [0, 0, 0, 0]
    assert [shard.owner_of(s, 2) for s in range(4)] == [0, 0, 1, 1]      # by subdomain
    assert [shard.owner_of(s, 4) for s in range(4)] == [0, 1, 2, 3]      # by species
    assert [shard.owned_mask(r, 2) for r in range(2)] == [0b0011, 0b1100]
    assert [shard.owned_mask(r, 4) for r in range(4)] == [1, 2, 4, 8]
    assert sum(shard.owned_mask(r, 4) for r in range(4)) == 0xF
    with pytest.raises(ValueError):
        shard.owner_of(0, 3)


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import pecs_b200 as pecs
    from pecs_b200 import shard, sweep
    from helpers import make_oracle

    class OracleEngine:
        # the engine interface of shard.ShardedStepper on the CPU oracle: owned carriers are solved here, the
        # others only receive densities
        def __init__(self, o, mask, n_cells):
            self.o, self.mask, self.n = o, mask, n_cells
        def step_local(self):
            if self.mask & 0b0011: self.o.assemble_semiconductor_rhs()
            if self.mask & 0b1100: self.o.assemble_electrolyte_rhs()
            for s in range(4):
                if self.mask >> s & 1: self.o.solve_species(s)
        def step_finish(self):
            self.o.assemble_Poisson_rhs(); self.o.solve_Poisson()
        def density(self, s):
            return torch.from_numpy(self.o.solution(s)[8 * self.n[s // 2]:].copy())
        def store_density(self, s, t):
            u = self.o.solution(s).copy(); u[8 * self.n[s // 2]:] = t.numpy(); self.o.set_vector(s, 0, u)

    rank, local, world, dist = sweep.init_distributed(backend="gloo")
    prob = pecs.SolarCellProblem(pecs.default_input_file(2, 1))
    prob.setup_full_system_host()
    def fresh():
        o = make_oracle(prob, True)
        o.project_initial_conditions(); o.assemble_Poisson_rhs(); o.solve_Poisson()
        return o
    reference = fresh(); reference.step(4)
    o = fresh()
    mask = shard.owned_mask(rank, world)
    stepper = shard.ShardedStepper(OracleEngine(o, mask, [prob.n_cells(0), prob.n_cells(1)]), dist, rank, world)
    stepper.step(4)
    for s in range(4):
        n = prob.n_cells(s // 2)
        assert np.array_equal(o.solution(s)[8 * n:], reference.solution(s)[8 * n:]), "density %%d" %% s
        if mask >> s & 1:
            assert np.array_equal(o.solution(s), reference.solution(s)), "owned currents %%d" %% s
    assert np.array_equal(o.solution(4), reference.solution(4))
    sys.stdout.write("rank " + str(rank) + " mask " + str(mask) + " sharded == single" + chr(10))
    sys.stdout.flush()
""") % (ROOT, ROOT)


def test_two_rank_subdomain_sharding_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 mask 3 sharded == single" in r.stdout and "rank 1 mask 12 sharded == single" in r.stdout
