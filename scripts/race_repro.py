"""GPU box: reproducer for the round-1 overlap failure (VERDICT W2).

One process = one configuration (the switches are read once per process / at graph capture):

    PECS_B200_LIB=...           which library (default build, or lib/libpecs_b200_nc.so = round-1 non-coherent loads)
    PECS_B200_DEFER_CURRENTS=1  recovery of the LDG currents overlaps the Poisson part
    PECS_B200_PDL=0             plain launches instead of programmatic dependent launches

    python scripts/race_repro.py --g 4 --reps 300 --steps 25 --tag defer1 --out gpurun_out/race_repro.jsonl

Every repetition: reset the five states -> `steps` IMEX steps through the step graph -> compare
  (a) bit for bit with the result stored by the FIRST configuration that ran (--ref file; the scheduling switches do
      not change any arithmetic, so every configuration must reproduce it exactly), and
  (b) with the CPU oracle's states after the same steps (densities and potential, 1e-9: the parity tolerance).
A stale read shows up in (a) whatever its size; (b) says whether it would have failed the parity tests.
Appends one JSON line per configuration to --out.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pecs_b200 as pecs  # noqa: E402
from helpers import make_oracle, rel_err  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--g", type=int, default=4)
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--tag", default="default")
    ap.add_argument("--ref", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "race_repro.jsonl"))
    ap.add_argument("--no-oracle", action="store_true",
                    help="bitwise comparison only (the oracle's sparse LU takes 156 s at g=5 and hours at g=6)")
    a = ap.parse_args()
    ref_path = a.ref or os.path.join(ROOT, "gpurun_out", f"race_ref_g{a.g}_s{a.steps}.npz")

    prob = pecs.SolarCellProblem(pecs.default_input_file(a.g, 1))
    prob.setup_full_system()
    start = [prob.get_solution(s) for s in range(5)]
    # the oracle's states after the same steps from the same start
    use_oracle = not a.no_oracle and a.g <= 5
    if use_oracle:
        o = make_oracle(prob, True)
        for s in range(5):
            o.set_vector(s, 0, start[s])
        o.step(a.steps)
        want = [o.solution(s) for s in range(5)]
    n_rt = prob.n_rt

    def oracle_err(got):
        if not use_oracle:
            return 0.0
        worst = 0.0
        for s in range(4):
            nc = got[s].size // 12
            worst = max(worst, rel_err(got[s][8 * nc:], want[s][8 * nc:]))
        return max(worst, rel_err(got[4][n_rt:], want[4][n_rt:]), rel_err(got[4][:n_rt], want[4][:n_rt]))

    ref = None
    if os.path.exists(ref_path):
        z = np.load(ref_path)
        ref = [z[f"s{s}"] for s in range(5)]
    bit_bad, par_bad, worst_bit, worst_par, first_bad = 0, 0, 0.0, 0.0, None
    for r in range(a.reps):
        for s in range(5):
            prob.set_solution(s, start[s])
        prob.step(a.steps)
        got = [prob.get_solution(s) for s in range(5)]
        if ref is None:
            ref = got
            os.makedirs(os.path.dirname(ref_path), exist_ok=True)
            np.savez(ref_path, **{f"s{s}": got[s] for s in range(5)})
        if not all(np.array_equal(x, y) for x, y in zip(got, ref)):
            bit_bad += 1
            if first_bad is None:
                first_bad = r
            worst_bit = max(worst_bit, max(rel_err(x, y) for x, y in zip(got, ref)))
        e = oracle_err(got)
        worst_par = max(worst_par, e)
        par_bad += e > 1e-9
    line = {"tag": a.tag, "g": a.g, "reps": a.reps, "steps": a.steps,
            "lib": os.path.basename(os.environ.get("PECS_B200_LIB", "libpecs_b200.so")),
            "defer_currents": os.environ.get("PECS_B200_DEFER_CURRENTS", "default"),
            "pdl": os.environ.get("PECS_B200_PDL", "1"),
            "bitwise_mismatches": bit_bad, "first_mismatch_rep": first_bad, "worst_bitwise_rel_diff": worst_bit,
            "oracle_parity_failures": int(par_bad) if use_oracle else None,
            "worst_oracle_rel_err": worst_par if use_oracle else None}
    print(json.dumps(line), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "a") as f:
        f.write(json.dumps(line) + "\n")
    prob.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
