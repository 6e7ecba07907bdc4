"""N > 1 path of bench.py on CPU: world_size-2 gloo process group, one applied bias per rank, max-over-ranks timing
(the data path has no collective: contexts of different biases are independent)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    from pecs_b200 import sweep
    import pecs_b200 as pecs
    rank, local, world, dist = sweep.init_distributed(backend="gloo")
    assert world == 2 and dist is not None
    bias = sweep.bias_for_rank(rank, world)
    # every rank builds ITS OWN problem (host part only on a CPU box): different Dirichlet data, same mesh
    prob = pecs.SolarCellProblem(pecs.default_input_file(2, 1, physical__insulated=False, physical__applied_bias=bias))
    prob.setup_full_system_host()
    phi_app = dict(zip(pecs.PARAM_NAMES, prob.params))["phi_app"]
    assert abs(phi_app - bias / 0.02585) < 1e-12
    sweep.barrier(dist)
    t = sweep.max_over_ranks(1.0 + rank, dist)
    total = sweep.sum_over_ranks(prob.n_cells(0), dist)
    assert t == 2.0 and total == 2 * prob.n_cells(0)
    # one I-V point per rank from a host state (constant densities: closed form), gathered into the curve
    import numpy as np
    states = [np.full(12 * prob.n_cells(w // 2), c + rank) for w, c in enumerate((5.0, 3.0, 7.0, 11.0))]
    curve = sweep.gather_iv(dist, bias, prob.interface_currents(states))
    p = prob.params
    length = np.hypot(0.3, 1.0)
    assert [round(b, 12) for b, _, _ in curve] == [0.0, 0.05]
    for r, (b, i_et, i_ht) in enumerate(curve):
        assert abs(i_et - p[9] * (5.0 + r - p[16]) * (11.0 + r) * length) <= 1e-12 * abs(i_et)
        assert abs(i_ht - p[10] * (3.0 + r - p[17]) * (7.0 + r) * length) <= 1e-12 * abs(i_ht)
    sys.stdout.write("rank " + str(rank) + " bias " + str(bias) + " ok" + chr(10))  # ONE write: the ranks share the pipe
    sys.stdout.flush()
""") % ROOT


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 bias 0.0 ok" in r.stdout and "rank 1 bias 0.05 ok" in r.stdout


def test_bias_assignment():
    from pecs_b200 import sweep
    assert [sweep.bias_for_rank(r, 8) for r in range(8)] == [0.05 * r for r in range(8)]
