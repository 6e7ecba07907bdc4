"""Target of the ncu passes (GPU box): set up the cfg workload, run `--steps` IMEX steps.  Never a timing source."""
import argparse
import sys

sys.path.insert(0, ".")
import pecs_b200 as pecs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("-g", type=int, default=7)
ap.add_argument("-l", type=int, default=1)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--five-calls", action="store_true", help="use the five reference-named calls instead of the graph")
a = ap.parse_args()
prob = pecs.SolarCellProblem(pecs.default_input_file(a.g, a.l))
prob.setup_full_system()
prob.synchronize()
for _ in range(a.steps):
    if a.five_calls:
        prob.assemble_semiconductor_rhs()
        prob.assemble_electrolyte_rhs()
        prob.solve_full_system()
        prob.assemble_Poisson_rhs()
        prob.solve_Poisson()
    else:
        prob.step(1)
prob.synchronize()
print("done", prob.info(0), "launches per step")
