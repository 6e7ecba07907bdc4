"""Generate tests/golden/*.npz from the CPU oracle.

The reference itself ships no golden vectors and cannot be built in this container (deal.II / UMFPACK / TBB are
absent), so these fixtures are ORACLE outputs, frozen so that (a) any later change of the oracle is noticed and
(b) the GPU path can be checked on the GPU box against numbers that were produced here.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pecs_b200 as pecs  # noqa: E402
from helpers import make_oracle, perturbed  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def production_case(g, l, n_steps, **overrides):
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **overrides))
    prob.setup_full_system_host()
    o = make_oracle(prob, True)
    o.project_initial_conditions()
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    out = {"poisson_rhs_initial": o.rhs(4), "poisson_solution_initial": o.solution(4)}
    # one assembly pass from a perturbed state
    saved = [o.solution(s) for s in range(5)]
    for s in range(4):
        o.set_vector(s, 0, perturbed(saved[s], 1234 + s))
    o.set_vector(4, 0, perturbed(saved[4], 99))
    o.assemble_semiconductor_rhs()
    o.assemble_electrolyte_rhs()
    o.assemble_Poisson_rhs()
    for s in range(5):
        out[f"rhs_perturbed_{s}"] = o.rhs(s)
    for s in range(5):
        o.set_vector(s, 0, saved[s])
    o.step(n_steps)
    for s in range(5):
        out[f"state_after_steps_{s}"] = o.solution(s)
    out["n_steps"] = np.array(n_steps)
    return out


def output_case(g, l, n_steps, **overrides):
    """output path and I-V post-processing of the oracle state after n_steps (oracle/output.py)"""
    from oracle import output as oracle_output
    state = production_case(g, l, n_steps, **overrides)
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **overrides))
    prob.setup_full_system_host()
    sc = prob.output_scales
    u = [state[f"state_after_steps_{s}"] for s in range(5)]
    field, potential = oracle_output.poisson_patches(prob.mesh(2)["vertices"], prob.poisson_face_dofs(), prob.n_rt, u[4],
                                                     sc[1], sc[0])
    p = prob.params
    out = {"field": field, "potential": potential, "n_steps": np.array(n_steps),
           "interface_currents": oracle_output.interface_currents(prob.mesh(0)["vertices"], prob.interface_pairs(), u[:4],
                                                                  p[9], p[10], p[16], p[17])}
    for s in range(4):
        out[f"current_{s}"], out[f"density_{s}"] = oracle_output.carrier_patches(u[s], sc[3])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "production_g2_l1.npz"), **production_case(2, 1, 10))
    np.savez_compressed(os.path.join(HERE, "production_g3_l1_biased.npz"),
                        **production_case(3, 1, 5, physical__insulated=False, physical__applied_bias=0.1))
    np.savez_compressed(os.path.join(HERE, "output_g3_l1_biased.npz"),
                        **output_case(3, 1, 5, physical__insulated=False, physical__applied_bias=0.1))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
