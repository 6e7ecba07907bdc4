"""GPU box: which cells of the production carrier kernels differ from the point-by-point kernel (variant 0)?"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import pecs_b200 as pecs  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prob = pecs.SolarCellProblem(pecs.default_input_file(g, 1))
prob.setup_full_system()
prob.step(2)
res = {}
for v in ("0", "1", "2"):
    os.environ["PECS_B200_RHS_KERNEL"] = v
    for s in range(4):
        prob.set_rhs(s, np.full(prob.n_dofs(s), np.nan))
    prob.assemble_semiconductor_rhs()
    prob.assemble_electrolyte_rhs()
    res[v] = [prob.get_rhs(s) for s in range(4)]
for v in ("1", "2"):
    for s in range(4):
        a, b = res[v][s], res["0"][s]
        n = a.size // 12
        bad = np.zeros(n, bool)
        for blk in range(3):
            x, y = a[4 * n * blk:4 * n * (blk + 1)].reshape(n, 4), b[4 * n * blk:4 * n * (blk + 1)].reshape(n, 4)
            scale = np.abs(y).max()
            bad |= ~(np.abs(x - y).max(axis=1) <= 1e-12 * scale)
        mesh = prob.mesh(s // 2)
        bdry = (mesh["face_kind"] == 1).any(axis=1) if "face_kind" in mesh else None
        print(f"variant {v} species {s}: {bad.sum()} of {n} cells differ; nan cells {np.isnan(a.reshape(3, n, 4)).any(axis=(0, 2)).sum()}",
              "first bad:", np.flatnonzero(bad)[:10])
        if bad.any():
            c = np.flatnonzero(bad)[0]
            print("   got ", a.reshape(3, n, 4)[:, c, :].ravel())
            print("   want", b.reshape(3, n, 4)[:, c, :].ravel())
