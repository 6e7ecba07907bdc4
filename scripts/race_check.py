"""GPU box: determinism of the step graph.  Every kernel has a fixed summation order, so repeating the same N steps
from the same state must reproduce the result bit for bit; a mismatch means a race.
    python scripts/race_check.py [g] [reps] [steps]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import pecs_b200 as pecs  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 25
prob = pecs.SolarCellProblem(pecs.default_input_file(g, 1))
prob.setup_full_system()
start = [prob.get_solution(s) for s in range(5)]
first, bad, worst = None, 0, 0.0
for r in range(reps):
    for s in range(5):
        prob.set_solution(s, start[s])
    prob.step(steps)
    got = [prob.get_solution(s) for s in range(5)]
    if first is None:
        first = got
        continue
    same = all(np.array_equal(a, b) for a, b in zip(got, first))
    if not same:
        bad += 1
        worst = max(worst, max(np.abs(a - b).max() / np.abs(b).max() for a, b in zip(got, first)))
print(f"g {g}: {bad} of {reps - 1} repetitions of {steps} steps differ from the first (worst rel diff {worst:.3e})", flush=True)
prob.close()
