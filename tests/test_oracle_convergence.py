"""Pins the CPU oracle: the reference's own manufactured-solution tests (tests/Poisson_test.cpp,
tests/IMEX_LDG_test.cpp, tests/DD_Poisson_test.cpp) hold no numbers, only analytic solutions and the documented
expectation of L2 order k+1 for the primary variables with dt = h^(k+1) (reference include/SolarCell.hpp:179-181,
229-235).  This is all the reference offers as known answers: raw-vector parity stays unpinned."""
import numpy as np

import pecs_b200 as pecs
from helpers import make_oracle


def _steady(level):
    prob = pecs.SolarCellProblem(None, test_defaults=True)
    prob.setup_test_host(pecs.KIND_TEST_STEADY, level)
    o = make_oracle(prob, False, 0.0)
    o.assemble_test_steady_rhs()
    o.solve_Poisson()
    o.solve_species(0)
    return np.concatenate([o.ldg_errors(0, 0.0), o.mixed_errors()])


def _transient(level, coupled):
    prob = pecs.SolarCellProblem(None, test_defaults=True)
    prob.setup_test_host(pecs.KIND_TEST_DD_POISSON if coupled else pecs.KIND_TEST_TRANSIENT, level)
    o = make_oracle(prob, False, 1.0)
    o.project_test_initial_condition()
    dt, t = prob.delta_t, 0.0
    while t < 1.0:
        if coupled:
            o.assemble_coupled_Poisson_test_rhs(t)
            o.solve_Poisson()
            o.assemble_coupled_DD_test_rhs(t)
        else:
            o.assemble_test_transient_rhs(t)
        o.solve_species(0)
        t += dt
    e = o.ldg_errors(2 if coupled else 1, t)
    return np.concatenate([e, o.mixed_errors()]) if coupled else e


def _rates(errors):
    return np.log2(errors[:-1] / errors[1:])


def test_steady_state_orders():
    e = np.array([_steady(n) for n in (2, 3, 4)])
    r = _rates(e)
    assert r[-1, 0] > 1.9          # u:   k+1 = 2
    assert r[-1, 1] > 0.9          # J:   k (penalty tau/h)
    assert r[-1, 2] > 0.95         # Phi: RT0 x DG0 -> 1
    assert r[-1, 3] > 1.9          # D at the trapezoid points


def test_ldg_imex_orders():
    e = np.array([_transient(n, False) for n in (2, 3, 4)])
    r = _rates(e)
    assert r[-1, 0] > 1.9 and r[-1, 1] > 0.9


def test_dd_poisson_orders():
    e = np.array([_transient(n, True) for n in (2, 3, 4)])
    r = _rates(e)
    assert r[-1, 0] > 1.9 and r[-1, 2] > 0.95 and r[-1, 3] > 1.9
