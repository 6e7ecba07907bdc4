"""Development check (GPU box): step graph vs solve-only graph timing, alternated."""
import sys
sys.path.insert(0, ".")
import pecs_b200 as pecs
prob = pecs.SolarCellProblem(pecs.default_input_file(int(sys.argv[1]) if len(sys.argv) > 1 else 7, 1))
prob.setup_full_system()
prob.step(5); prob.synchronize()
for rep in range(3):
    a = prob.step_timed(30)[0] / 30
    b = prob.step_timed(30, sectioned=2)[0] / 30
    c = prob.step_timed(30, sectioned=3)[0] / 30
    prob.step(3)
    print(f"rep {rep}: step graph {a:.3f} ms, solve-only graph {b:.3f} ms, assembly-only graph {c:.3f} ms", flush=True)
