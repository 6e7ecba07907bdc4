"""GPU parity at production scale: the meshes the tests of round 1 stopped 64x short of.

* g=6 (245 760 DoF per carrier): one pass of the hot path from a perturbed state -- five right-hand sides against the
  oracle's assembly at 1e-12, the five solves through host-CSR residuals, finiteness (helpers.validate_workload; the
  same routine bench.py runs at cfg3, g=7, and reports as `parity`).
* g=5: states after N steps against the oracle WITH its sparse LU (1e-9), and the LDG currents of one solve against
  an EXTENDED-PRECISION solution of the same full 12-unknowns-per-cell system (SuperLU + iterative refinement with
  long-double residuals): the currents are what the output path writes, and q = A^-1 (r_q - G u) amplifies density
  errors by 1/mobility-like factors that grow with 1/h.
"""
import numpy as np
import pytest

import pecs_b200 as pecs
from helpers import (SPECIES, block_rel_err, csr_matvec_longdouble, make_oracle, perturbed, rel_err,
                     validate_workload)

pytestmark = pytest.mark.gpu

RHS_TOL = 1e-12
STATE_TOL = 1e-9
CURRENT_TOL = 1e-7
BACKWARD_TOL = 1e-10   # normwise backward error |b - A x| / (|A| |x| + |b|) of a solve


def extended_precision_solve(A, b, sweeps=4):
    """x with A x = b to (nearly) long-double residual accuracy: SuperLU in fp64 + iterative refinement"""
    import scipy.sparse.linalg as spl
    lu = spl.splu(A.tocsc())
    x = lu.solve(b).astype(np.longdouble)
    for _ in range(sweeps):
        r = b.astype(np.longdouble) - csr_matvec_longdouble(A, x)
        x = x + lu.solve(r.astype(np.float64)).astype(np.longdouble)
    return x


def test_workload_parity_g6():
    prob = pecs.SolarCellProblem(pecs.default_input_file(6, 1))
    prob.setup_full_system()
    try:
        # a few real steps first so that the state is not the trivial initial one
        prob.step(3)
        v = validate_workload(prob)
        print("g=6 workload parity:", {k: v[k] for k in v if not isinstance(v[k], dict)})
        assert v["finite"] and prob.info(pecs.solarcell.INFO_SOLVE_WAIT_ERRORS) == 0
        assert v["rhs_rel"] <= RHS_TOL, v["rhs_rel_per_vector"]
        assert v["backward_err"] <= BACKWARD_TOL, v["residual_rel_per_system"]
        assert v["last_step_residual_rel"] <= 1e-9, v["last_step_residual_rel_per_system"]
        # the north_star quantities of one solve from the (deliberately rough) perturbed state
        assert v["solve_density_err"] <= STATE_TOL, v["solve_density_err_per_species"]
        assert v["solve_potential_err"] <= STATE_TOL and v["solve_field_err"] <= STATE_TOL
        assert v["solve_current_err"] <= CURRENT_TOL, v["solve_current_err_per_species"]
    finally:
        prob.close()


@pytest.fixture(scope="module")
def g5():
    prob = pecs.SolarCellProblem(pecs.default_input_file(5, 1))
    prob.setup_full_system()
    o = make_oracle(prob, True)
    o.project_initial_conditions()
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    yield prob, o
    o.close()
    prob.close()


def test_states_after_n_steps_g5(g5):
    prob, o = g5
    n = 10
    prob.step(n)
    o.step(n)
    for s in SPECIES:
        ug, uo = prob.get_solution(s), o.solution(s)
        nc = ug.size // 12
        assert np.isfinite(ug).all()
        assert rel_err(ug[8 * nc:], uo[8 * nc:]) <= STATE_TOL, f"density of species {s}"
        assert block_rel_err(ug, uo) <= CURRENT_TOL, f"currents of species {s}"
    n_rt = prob.n_rt
    xg, xo = prob.get_solution(pecs.POISSON), o.solution(4)
    assert rel_err(xg[n_rt:], xo[n_rt:]) <= STATE_TOL and rel_err(xg[:n_rt], xo[:n_rt]) <= STATE_TOL


def test_currents_against_extended_precision_g5(g5):
    """one solve of every carrier from a perturbed state; reference = long-double-refined solution of the full system"""
    prob, o = g5
    saved = [prob.get_solution(s) for s in range(5)]
    try:
        for s in SPECIES:
            prob.set_solution(s, perturbed(saved[s], 1234 + s))
        prob.set_solution(pecs.POISSON, perturbed(saved[4], 99))
        prob.assemble_semiconductor_rhs()
        prob.assemble_electrolyte_rhs()
        rhs = [prob.get_rhs(s) for s in SPECIES]
        prob.solve_full_system()
        report = {}
        for s in SPECIES:
            x = prob.get_solution(s)
            ref = extended_precision_solve(prob.matrix(s), rhs[s]).astype(np.float64)
            nc = x.size // 12
            e_density = rel_err(x[8 * nc:], ref[8 * nc:])
            e_current = max(rel_err(x[:4 * nc], ref[:4 * nc]), rel_err(x[4 * nc:8 * nc], ref[4 * nc:8 * nc]))
            # the oracle's own LU on the same right-hand side, for scale
            o.set_vector(s, 1, rhs[s])
            o.solve_species(s)
            xo = o.solution(s)
            report[s] = (e_density, e_current, rel_err(xo[8 * nc:], ref[8 * nc:]),
                         max(rel_err(xo[:4 * nc], ref[:4 * nc]), rel_err(xo[4 * nc:8 * nc], ref[4 * nc:8 * nc])))
        print("species: (GPU density, GPU current, oracle-LU density, oracle-LU current) rel. error vs extended precision")
        for s, r in report.items():
            print(f"  {s}: " + ", ".join(f"{v:.2e}" for v in r))
        for s, r in report.items():
            assert r[0] <= STATE_TOL, f"density of species {s}: {r[0]:.2e}"
            assert r[1] <= CURRENT_TOL, f"currents of species {s}: {r[1]:.2e}"
    finally:
        for s in range(5):
            prob.set_solution(s, saved[s])
