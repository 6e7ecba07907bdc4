"""GPU parity tests proper: the CUDA path, driven through the C ABI, against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): assembled right-hand-side vectors <= 1e-12 relative, carrier densities and
potential <= 1e-9 relative after N IMEX steps.  "Relative" is max|a-b| / max|b| per component block.
"""
import numpy as np
import pytest

import pecs_b200 as pecs
from helpers import SPECIES, block_rel_err, make_oracle, perturbed, rel_err

pytestmark = pytest.mark.gpu

RHS_TOL = 1e-12
STATE_TOL = 1e-9
CURRENT_TOL = 1e-7  # LDG currents: derived quantities, see test_rhs_vectors_perturbed_state


@pytest.fixture(scope="module")
def production():
    """cfg1: the reference's default input_file.prm (g=4, l=1)."""
    prob = pecs.SolarCellProblem(pecs.default_input_file(4, 1))
    prob.setup_full_system()
    o = make_oracle(prob, True)
    o.project_initial_conditions()
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    yield prob, o
    prob.close()


def test_initial_poisson(production):
    prob, o = production
    assert rel_err(prob.get_rhs(pecs.POISSON), o.rhs(4)) <= RHS_TOL
    n_rt = prob.n_rt
    xg, xo = prob.get_solution(pecs.POISSON), o.solution(4)
    assert rel_err(xg[:n_rt], xo[:n_rt]) <= STATE_TOL
    assert rel_err(xg[n_rt:], xo[n_rt:]) <= STATE_TOL


def test_rhs_vectors_perturbed_state(production):
    """every assembled vector of one step, from an identical non-trivial state on both sides"""
    prob, o = production
    saved = [prob.get_solution(s) for s in range(5)]
    try:
        for s in SPECIES:
            u = perturbed(o.solution(s), 1234 + s)
            prob.set_solution(s, u)
            o.set_vector(s, 0, u)
        X = perturbed(o.solution(4), 99)
        prob.set_solution(pecs.POISSON, X)
        o.set_vector(4, 0, X)
        prob.assemble_semiconductor_rhs()
        prob.assemble_electrolyte_rhs()
        prob.assemble_Poisson_rhs()
        o.assemble_semiconductor_rhs()
        o.assemble_electrolyte_rhs()
        o.assemble_Poisson_rhs()
        for s in SPECIES:
            assert block_rel_err(prob.get_rhs(s), o.rhs(s)) <= RHS_TOL, f"species {s}"
        assert rel_err(prob.get_rhs(pecs.POISSON), o.rhs(4)) <= RHS_TOL
        # and the five solves on those right-hand sides
        prob.solve_full_system()
        prob.solve_Poisson()
        o.solve_full_system()
        o.solve_Poisson()
        # Densities and potential carry the north_star tolerance.  The currents of this deliberately rough state are
        # small differences of large terms (q = A^-1 (r_q - G u)): a density error is amplified ~150x for the redox
        # species (mobility 2.6e-6).  Against an extended-precision refinement of the same system the oracle's LU is
        # itself 5e-10 off in that block, SuperLU 2e-10, the Schur-reduced solve 2.6e-9 -- hence CURRENT_TOL there.
        for s in SPECIES:
            ug, uo = prob.get_solution(s), o.solution(s)
            nc = ug.size // 12
            assert rel_err(ug[8 * nc:], uo[8 * nc:]) <= STATE_TOL, f"density of species {s}"
            assert block_rel_err(ug, uo) <= CURRENT_TOL, f"currents of species {s}"
        n_rt = prob.n_rt
        xg, xo = prob.get_solution(pecs.POISSON), o.solution(4)
        assert rel_err(xg[:n_rt], xo[:n_rt]) <= STATE_TOL and rel_err(xg[n_rt:], xo[n_rt:]) <= STATE_TOL
    finally:
        for s in range(5):
            prob.set_solution(s, saved[s])
            o.set_vector(s, 0, saved[s])


def test_states_after_n_steps(production):
    """N IMEX steps through the captured graph vs N oracle steps"""
    prob, o = production
    n = 25
    prob.step(n)
    o.step(n)
    for s in SPECIES:
        ug, uo = prob.get_solution(s), o.solution(s)
        nc = ug.size // 12
        assert rel_err(ug[8 * nc:], uo[8 * nc:]) <= STATE_TOL, f"density of species {s}"
        assert block_rel_err(ug, uo) <= CURRENT_TOL, f"currents of species {s}"
    n_rt = prob.n_rt
    xg, xo = prob.get_solution(pecs.POISSON), o.solution(4)
    assert rel_err(xg[n_rt:], xo[n_rt:]) <= STATE_TOL
    assert rel_err(xg[:n_rt], xo[:n_rt]) <= STATE_TOL


def test_step_equals_five_calls(production):
    """the graph replay and the five reference-named calls are the same arithmetic: bit-identical states"""
    prob, o = production
    saved = [prob.get_solution(s) for s in range(5)]
    prob.step(2)
    a = [prob.get_solution(s) for s in range(5)]
    for s in range(5):
        prob.set_solution(s, saved[s])
    for _ in range(2):
        prob.assemble_semiconductor_rhs()
        prob.assemble_electrolyte_rhs()
        prob.solve_full_system()
        prob.assemble_Poisson_rhs()
        prob.solve_Poisson()
    b = [prob.get_solution(s) for s in range(5)]
    for s in range(5):
        assert np.array_equal(a[s], b[s])
        prob.set_solution(s, saved[s])


def test_interface_currents_cancel(production):
    """structural identity: the interface terms of reductants and oxidants cancel exactly, so
    rhs_r + rhs_o has no interface contribution: compare against a run with the transfer rates switched off
    is not available per call, so check the antisymmetry directly on the density rows of interface cells."""
    prob, o = production
    prob.assemble_electrolyte_rhs()
    r, ox = prob.get_rhs(pecs.REDUCTANTS), prob.get_rhs(pecs.OXIDANTS)
    o.assemble_electrolyte_rhs()
    assert block_rel_err(r + ox, o.rhs(2) + o.rhs(3)) <= 1e-11


@pytest.mark.parametrize("kind,level", [(pecs.KIND_TEST_STEADY, 3), (pecs.KIND_TEST_TRANSIENT, 3),
                                        (pecs.KIND_TEST_DD_POISSON, 3), (pecs.KIND_TEST_DD_POISSON, 4)])
def test_manufactured_problem_matches_oracle(kind, level):
    """cfg2: the reference's convergence tests, GPU vs oracle states and the L2 errors against the analytic solution"""
    prob = pecs.SolarCellProblem(None, test_defaults=True)
    errors = prob.run_test(kind, level)
    o = make_oracle(prob, False, 0.0 if kind == pecs.KIND_TEST_STEADY else 1.0)
    dt = prob.delta_t
    if kind == pecs.KIND_TEST_STEADY:
        o.assemble_test_steady_rhs()
        o.solve_Poisson()
        o.solve_species(0)
        e_ldg, t = o.ldg_errors(0, 0.0), 0.0
    else:
        o.project_test_initial_condition()
        t = 0.0
        while t < 1.0:
            if kind == pecs.KIND_TEST_DD_POISSON:
                o.assemble_coupled_Poisson_test_rhs(t)
                o.solve_Poisson()
                o.assemble_coupled_DD_test_rhs(t)
            else:
                o.assemble_test_transient_rhs(t)
            o.solve_species(0)
            t += dt
        e_ldg = o.ldg_errors(1 if kind == pecs.KIND_TEST_TRANSIENT else 2, t)
    ug, uo = prob.get_solution(pecs.ELECTRONS), o.solution(0)
    assert block_rel_err(ug, uo) <= STATE_TOL
    assert abs(errors["u"] - e_ldg[0]) <= 1e-9 * e_ldg[0] + 1e-12
    assert abs(errors["J"] - e_ldg[1]) <= 1e-8 * e_ldg[1] + 1e-12
    if kind != pecs.KIND_TEST_TRANSIENT:
        e_mix = o.mixed_errors()
        xg, xo = prob.get_solution(pecs.POISSON), o.solution(4)
        assert rel_err(xg, xo) <= STATE_TOL
        assert abs(errors["Phi"] - e_mix[0]) <= 1e-9 * e_mix[0]
    prob.close()


def test_convergence_order_on_gpu():
    """cfg2 gate: L2 error of u reduces with order k+1 = 2 (reference include/SolarCell.hpp:179-181, 229-235)"""
    errs = []
    for level in (3, 4, 5):
        prob = pecs.SolarCellProblem(None, test_defaults=True)
        errs.append(prob.run_test(pecs.KIND_TEST_DD_POISSON, level))
        prob.close()
    for a, b in zip(errs[:-1], errs[1:]):
        assert np.log2(a["u"] / b["u"]) >= 1.9
        assert np.log2(a["Phi"] / b["Phi"]) >= 0.95


@pytest.mark.parametrize("kind,levels,what", [
    (pecs.KIND_TEST_STEADY, (4, 5, 6, 7), ("u", "Phi")),      # Poisson_test: LDG Poisson + mixed FEM, steady
    (pecs.KIND_TEST_DD_POISSON, (4, 5, 6), ("u", "Phi"))])    # DD_Poisson_test: the coupled problem
def test_convergence_gate_refinements_4_to_7(kind, levels, what):
    """BASELINE config 2: the reference's manufactured-solution programs at the refinements it runs them on, on the
    GPU path, against the analytic solutions: L2 order k+1 = 2 for the density, 1 for the DG0 potential
    (reference include/SolarCell.hpp:179-181, 229-235; dt = h^(k+1) makes the transient levels 4x longer each)"""
    errs = []
    for level in levels:
        prob = pecs.SolarCellProblem(None, test_defaults=True)
        errs.append(prob.run_test(kind, level))
        prob.close()
    for a, b in zip(errs[:-1], errs[1:]):
        assert np.log2(a["u"] / b["u"]) >= 1.9
        if "Phi" in what:
            assert np.log2(a["Phi"] / b["Phi"]) >= 0.95


# L2 density errors of the CPU oracle on IMEX_LDG_test (tests/test_oracle_convergence.py::_transient, run here on the
# CPU; refinement 6 takes 8 minutes there).  The reference's program runs refinements 2..5 (tests/IMEX_LDG_test.cpp:19-20).
# The restated scheme converges with order 2.00 (3->4), 1.82 (4->5) and 1.07 (5->6): past refinement 5 the error of this
# problem stops following h^2 in the oracle as well -- the gate for the GPU path is therefore "the oracle's errors".
TRANSIENT_ORACLE_ERRORS = {3: 0.04173435, 4: 0.010446326, 5: 0.0029485025, 6: 0.00140736}


def test_transient_gate_matches_oracle_errors():
    """BASELINE config 2, IMEX_LDG_test at refinements 4-6 on the GPU path: the L2 error against the analytic solution
    equals the oracle's at every level, and the order is k+1 where the oracle's is (3->4->5)."""
    errs = {}
    for level in (3, 4, 5, 6):
        prob = pecs.SolarCellProblem(None, test_defaults=True)
        errs[level] = prob.run_test(pecs.KIND_TEST_TRANSIENT, level)["u"]
        prob.close()
        assert abs(errs[level] - TRANSIENT_ORACLE_ERRORS[level]) <= 1e-5 * TRANSIENT_ORACLE_ERRORS[level], level
    assert np.log2(errs[3] / errs[4]) >= 1.9
    assert np.log2(errs[4] / errs[5]) >= 1.8
