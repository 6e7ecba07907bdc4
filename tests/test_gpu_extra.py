"""More GPU checks through the C ABI: committed golden fixtures, host-buffer stepping, solver variants."""
import os
import subprocess
import sys

import numpy as np
import pytest

import pecs_b200 as pecs
from helpers import rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name,g,overrides", [
    ("production_g2_l1.npz", 2, {}),
    ("production_g3_l1_biased.npz", 3, {"physical__insulated": False, "physical__applied_bias": 0.1})])
def test_golden_fixtures(name, g, overrides):
    """fixtures were produced by the oracle in the build container (tests/golden/make_golden.py)"""
    gold = np.load(os.path.join(GOLDEN, name))
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, 1, **overrides))
    prob.setup_full_system()
    assert rel_err(prob.get_rhs(pecs.POISSON), gold["poisson_rhs_initial"]) <= 1e-12
    assert rel_err(prob.get_solution(pecs.POISSON), gold["poisson_solution_initial"]) <= 1e-9
    prob.step(int(gold["n_steps"]))
    for s in range(5):
        ref = gold[f"state_after_steps_{s}"]
        got = prob.get_solution(s)
        if s < 4:
            nc = got.size // 12
            assert rel_err(got[8 * nc:], ref[8 * nc:]) <= 1e-9
        else:
            n_rt = prob.n_rt
            assert rel_err(got[n_rt:], ref[n_rt:]) <= 1e-9 and rel_err(got[:n_rt], ref[:n_rt]) <= 1e-9
    prob.close()


def test_device_warmup_is_optional_and_repeatable():
    """pecs_device_warmup pays the first-context costs ahead of pecs_ctx_create; calling it, calling it twice or not at
    all changes nothing; a bad ordinal is refused"""
    pecs.device_warmup(0)
    pecs.device_warmup(0)
    with pytest.raises(pecs.PecsError):
        pecs.device_warmup(pecs.device_count())


def test_step_host_equals_step():
    """the host-buffer entry point (H2D, step, D2H) is the same arithmetic as the resident path"""
    prob = pecs.SolarCellProblem(pecs.default_input_file(3, 1))
    prob.setup_full_system()
    start = [prob.get_solution(s) for s in range(5)]
    prob.step(3)
    resident = [prob.get_solution(s) for s in range(5)]
    states = prob.pinned_states()
    for s in range(5):
        states[s][:] = start[s]
    for _ in range(3):
        prob.step_host(1, states)
    for s in range(5):
        assert np.array_equal(states[s], resident[s])
    assert prob.info(0) > 0 and prob.info(1) > 0
    ms = prob.step_timed(2, sectioned=True)
    assert ms[0] > 0 and abs(ms[1:].sum() - ms[0]) < 1e-6 * ms[0] + 1e-9
    prob.close()


VARIANT = """
import sys, numpy as np
sys.path.insert(0, %r)
import pecs_b200 as pecs
prob = pecs.SolarCellProblem(pecs.default_input_file(3, 1))
prob.setup_full_system()
prob.step(4)
np.save(sys.argv[1], np.concatenate([prob.get_solution(s) for s in range(5)]))
"""


def test_solver_variants_agree(tmp_path):
    """default plan (Schur reduction, unknown-level separators, Poisson edge separators, device factorisation) vs
    full 12-unknown systems, host factorisation, other leaf sizes, cell separators, other pipeline shapes"""
    script = tmp_path / "variant.py"
    script.write_text(VARIANT % ROOT)
    results = []
    for env in ({}, {"PECS_B200_NO_SCHUR": "1"}, {"PECS_B200_HOST_FACTOR": "1"}, {"PECS_B200_LEAF_NODES": "3"},
                {"PECS_B200_CELL_SEPARATORS": "1", "PECS_B200_SCHUR_DROP": "0", "PECS_B200_POISSON_CELL_NODES": "1"},
                {"PECS_B200_SOLVE_STAGES": "2", "PECS_B200_SOLVE_PANELS_PER_TILE": "3"}):
        out = tmp_path / ("out_" + "_".join(env) + ".npy")
        r = subprocess.run([sys.executable, str(script), str(out)], env=dict(os.environ, **env), capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        results.append(np.load(out))
    for other in results[1:]:
        assert rel_err(other, results[0]) <= 1e-9


CASES = {
    # BASELINE.json config 4: strongly non-affine conic wire, two levels of interface refinement (2:1 smoothing layer)
    "conic_l2": (3, 2, {"mesh__radius_one": 0.2, "mesh__radius_two": 0.6}),
    # no local refinement: no hanging faces, no Poisson constraints besides the Neumann edges
    "uniform_l0": (3, 0, {}),
    # dark, no Schottky contact, not insulated with an applied bias: every boundary branch the default input skips
    "dark_biased": (2, 1, {"physical__illumination_status": False, "physical__schottky_status": False,
                           "physical__insulated": False, "physical__applied_bias": 0.2}),
    # non-zero Shockley-Read-Hall recombination (the formula the reference carries as a comment), O(1) here
    "srh_on": (3, 1, {"physical__srh_recombination": True, "physical__intrinsic_density": 0.5e16,
                      "electrons__recombination_time": 2e-12, "holes__recombination_time": 1e-12}),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_other_configurations_match_oracle(case):
    """RHS vectors (1e-12) and densities / potential after 10 steps (1e-9) against the oracle on the meshes and
    switches of the other BASELINE configurations"""
    from helpers import SPECIES, block_rel_err, make_oracle
    g, l, overrides = CASES[case]
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **overrides))
    prob.setup_full_system()
    o = make_oracle(prob, True)
    o.project_initial_conditions()
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    assert rel_err(prob.get_solution(pecs.POISSON), o.solution(4)) <= 1e-9
    prob.step(3)
    o.step(3)
    prob.assemble_semiconductor_rhs()
    prob.assemble_electrolyte_rhs()
    o.assemble_semiconductor_rhs()
    o.assemble_electrolyte_rhs()
    for s in SPECIES:
        # both sides assemble from their own (1e-12-close) states: the RHS tolerance is that of the states here
        assert block_rel_err(prob.get_rhs(s), o.rhs(s)) <= 1e-9, f"rhs of species {s}"
    prob.step(7)
    o.step(7)
    for s in SPECIES:
        ug, uo = prob.get_solution(s), o.solution(s)
        nc = ug.size // 12
        assert rel_err(ug[8 * nc:], uo[8 * nc:]) <= 1e-9, f"density of species {s}"
        assert block_rel_err(ug, uo) <= 1e-7, f"currents of species {s}"
    n_rt = prob.n_rt
    xg, xo = prob.get_solution(pecs.POISSON), o.solution(4)
    assert rel_err(xg[n_rt:], xo[n_rt:]) <= 1e-9 and rel_err(xg[:n_rt], xo[:n_rt]) <= 1e-9
    prob.close()


@pytest.mark.parametrize("g,l,overrides", [(4, 1, {}), (3, 2, {"mesh__radius_one": 0.2})])
def test_production_rhs_kernels_agree(g, l, overrides, monkeypatch):
    """The production carrier kernels (0: point-by-point, 1: sum-factorised = default, its 64- and 32-thread launch
    shapes, 2: sum-factorised streaming kernel, also with several tiles per block) and the static-table Poisson rows,
    from one perturbed state: each within 1e-12 of the oracle, the sum-factorised ones bit-identical to each other."""
    from helpers import SPECIES, block_rel_err, make_oracle, perturbed
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **overrides))
    prob.setup_full_system()
    o = make_oracle(prob, True)
    o.project_initial_conditions()
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    for s in SPECIES:
        u = perturbed(o.solution(s), 4321 + s)
        prob.set_solution(s, u)
        o.set_vector(s, 0, u)
    X = perturbed(o.solution(4), 77)
    prob.set_solution(pecs.POISSON, X)
    o.set_vector(4, 0, X)
    o.assemble_semiconductor_rhs()
    o.assemble_electrolyte_rhs()
    o.assemble_Poisson_rhs()
    got = {}
    for name, env in {"points": {"PECS_B200_RHS_KERNEL": "0"}, "direct": {"PECS_B200_RHS_KERNEL": "1"},
                      "direct64": {"PECS_B200_RHS_KERNEL": "14"}, "direct32": {"PECS_B200_RHS_KERNEL": "15"},
                      "stream": {"PECS_B200_RHS_KERNEL": "2"},
                      "stream3": {"PECS_B200_RHS_KERNEL": "2", "PECS_B200_RHS_GRID": "3"}}.items():
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for s in SPECIES:
            prob.set_rhs(s, np.full(prob.n_dofs(s), np.nan))
        prob.assemble_semiconductor_rhs()
        prob.assemble_electrolyte_rhs()
        got[name] = [prob.get_rhs(s) for s in SPECIES]
        for s in SPECIES:
            assert block_rel_err(got[name][s], o.rhs(s)) <= 1e-12, f"{name}: species {s}"
        monkeypatch.delenv("PECS_B200_RHS_GRID", raising=False)
    for s in SPECIES:
        for other in ("direct64", "direct32", "stream", "stream3"):
            assert np.array_equal(got["direct"][s], got[other][s]), other
    prob.assemble_Poisson_rhs()
    assert rel_err(prob.get_rhs(pecs.POISSON), o.rhs(4)) <= 1e-12
    prob.close()


def test_shared_factors_bit_identical(monkeypatch):
    """Reductants and oxidants have the same constant matrix at equal mobilities (reference source/LDG.cpp:624-678,
    input_file.prm:77,128): the default context factorises it once and solves the pair with two right-hand sides per
    pass over the tables.  Per right-hand side the arithmetic and its order are those of the one-vector kernels, so
    the states must equal those of two separate factorisations BIT FOR BIT -- through the graph, the five calls and
    pecs_solve_species."""
    def run(shared):
        if shared:
            monkeypatch.delenv("PECS_B200_NO_SHARED_FACTORS", raising=False)
        else:
            monkeypatch.setenv("PECS_B200_NO_SHARED_FACTORS", "1")
        prob = pecs.SolarCellProblem(pecs.default_input_file(4, 1))
        prob.setup_full_system()
        bytes_per_step = prob.info(2)
        prob.step(6)
        after_graph = [prob.get_solution(s) for s in range(5)]
        prob.assemble_semiconductor_rhs()
        prob.assemble_electrolyte_rhs()
        for s in (3, 2, 1, 0):  # one species at a time: the single-vector path on the shared tables
            prob.solve_species(s)
        prob.assemble_Poisson_rhs()
        prob.solve_Poisson()
        after_calls = [prob.get_solution(s) for s in range(5)]
        prob.close()
        return bytes_per_step, after_graph, after_calls
    b1, g1, c1 = run(True)
    b0, g0, c0 = run(False)
    assert b1 < 0.85 * b0, "the shared factorisation must stream fewer bytes per step"
    for s in range(5):
        assert np.array_equal(g1[s], g0[s]), f"graph, vector {s}"
        assert np.array_equal(c1[s], c0[s]), f"calls, vector {s}"
    # different mobilities: no sharing, and still the oracle's answer (covered by the other tests through electrons / holes)


SCHEDULING = """
import sys, numpy as np
sys.path.insert(0, %r)
import pecs_b200 as pecs
prob = pecs.SolarCellProblem(pecs.default_input_file(4, 1))
prob.setup_full_system()
out = []
start = [prob.get_solution(s) for s in range(5)]
for rep in range(8):
    for s in range(5):
        prob.set_solution(s, start[s])
    prob.step(25)
    out.append(np.concatenate([prob.get_solution(s) for s in range(5)]))
assert all(np.array_equal(o, out[0]) for o in out), "repetitions differ"
np.save(sys.argv[1], out[0])
"""


def test_step_scheduling_variants_match_oracle(tmp_path):
    """Regression test for the round-1 overlap failure: the step graph with programmatic
    dependent launches on / off and with the deferred recovery of the currents on / off is the same arithmetic in a
    different schedule -- bit-identical states after 25 steps, 8 repetitions each, and all of them within the parity
    tolerance of the oracle.  (scripts/race_repro.py is the long version of this with hundreds of repetitions.)"""
    from helpers import SPECIES, make_oracle
    script = tmp_path / "sched.py"
    script.write_text(SCHEDULING % ROOT)
    results = {}
    for name, env in {"default": {}, "pdl_off": {"PECS_B200_PDL": "0"}, "defer_on": {"PECS_B200_DEFER_CURRENTS": "1"},
                      "defer_off": {"PECS_B200_DEFER_CURRENTS": "0"},
                      "defer_on_pdl_off": {"PECS_B200_DEFER_CURRENTS": "1", "PECS_B200_PDL": "0"}}.items():
        out = tmp_path / f"{name}.npy"
        r = subprocess.run([sys.executable, str(script), str(out)], env=dict(os.environ, **env), capture_output=True,
                           text=True, timeout=900)
        assert r.returncode == 0, f"{name}: {r.stderr[-2000:]}"
        results[name] = np.load(out)
    for name, v in results.items():
        assert np.array_equal(v, results["default"]), name
    prob = pecs.SolarCellProblem(pecs.default_input_file(4, 1))
    prob.setup_full_system_host()
    o = make_oracle(prob, True)
    o.project_initial_conditions()
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    o.step(25)
    off = 0
    for s in range(5):
        n = o.n_dofs(s)
        got, want = results["default"][off:off + n], o.solution(s)
        off += n
        if s < 4:
            assert rel_err(got[8 * (n // 12):], want[8 * (n // 12):]) <= 1e-9, f"density of species {s}"
        else:
            assert rel_err(got[prob.n_rt:], want[prob.n_rt:]) <= 1e-9
    prob.close()
