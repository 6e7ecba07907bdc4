"""Shared helpers of the test-suite: build the CPU oracle for a pecs_b200 problem and compare vectors."""
import numpy as np

import pecs_b200 as pecs
from oracle import Oracle

SPECIES = (pecs.ELECTRONS, pecs.HOLES, pecs.REDUCTANTS, pecs.OXIDANTS)


def make_oracle(prob, full_system, transient_or_steady=1.0, factor=True):
    """The oracle gets the MESH TABLES and the scaled parameters (what deal.II's Triangulation and the .prm file
    are to the reference) and rebuilds everything else itself: dofs, maps, matrices, factorisations."""
    o = Oracle(prob.params, full_system)
    o.set_mesh(0, prob.mesh(0))
    if full_system:
        o.set_mesh(1, prob.mesh(1))
    o.set_mesh(2, prob.mesh(2))
    o.setup(transient_or_steady, factor)
    return o


def rel_err(a, b):
    """max-norm of the difference relative to the max-norm of the reference vector b."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


def block_rel_err(a, b, n_blocks=3):
    """worst relative error over the component blocks [Jx | Jy | rho] (each block has its own scale)."""
    a, b = np.asarray(a), np.asarray(b)
    return max(rel_err(x, y) for x, y in zip(np.split(a, n_blocks), np.split(b, n_blocks)))


def perturbed(values, seed):
    """u <- u (1 + 1e-3 U(-1,1)) + 1e-3 U(-1,1): a non-trivial, reproducible state (seed 1234)."""
    rng = np.random.default_rng(seed)
    v = np.asarray(values)
    return v * (1.0 + 1e-3 * rng.uniform(-1, 1, v.size)) + 1e-3 * rng.uniform(-1, 1, v.size)


def validate_workload(prob, oracle=None, seed=1234):
    """Parity of ONE pass of the hot path on whatever mesh `prob` holds (used by bench.py at the benchmarked
    configuration and by tests/test_gpu_large.py), outside any timed region.  From a perturbed state (seed 1234):

      rhs_rel      all five assembled right-hand sides against the oracle's assembly of the same state, worst block,
                   max-norm relative (north_star: 1e-12).  The oracle is built WITHOUT its LU (factor=False).
      residual_rel max over the five systems of |b - A x|_inf / |b|_inf with the host CSR matrices (scipy mat-vec),
                   x = the vector pecs_solve_* returned for right-hand side b: reference Carrier.cpp:34-40 semantics
                   x = A^-1 b, checked without any second solver.  Constrained Poisson rows (hanging / Neumann
                   fluxes) are excluded: the solver eliminates them and `distribute` overwrites them.
      backward_err the same residuals normwise: |r|_inf / (|A|_inf |x|_inf + |b|_inf)
      finite       every state entry is finite
    The caller's state is restored."""
    import time
    t0 = time.perf_counter()
    saved = [prob.get_solution(s) for s in range(5)]
    o = oracle if oracle is not None else make_oracle(prob, True, factor=False)
    out = {}
    try:
        state = [perturbed(saved[s], seed + s) for s in range(4)] + [perturbed(saved[4], 99)]
        for s in range(5):
            prob.set_solution(s, state[s])
            o.set_vector(s, 0, state[s])
        prob.assemble_semiconductor_rhs()
        prob.assemble_electrolyte_rhs()
        prob.assemble_Poisson_rhs()
        o.assemble_semiconductor_rhs()
        o.assemble_electrolyte_rhs()
        o.assemble_Poisson_rhs()
        b = [prob.get_rhs(s) for s in range(5)]
        per = {f"species_{s}": block_rel_err(b[s], o.rhs(s)) for s in range(4)}
        per["poisson"] = rel_err(b[4], o.rhs(4))
        out["rhs_rel"] = max(per.values())
        out["rhs_rel_per_vector"] = per
        prob.solve_full_system()
        prob.solve_Poisson()
        x = [prob.get_solution(s) for s in range(5)]
        out["finite"] = bool(all(np.isfinite(v).all() for v in x))
        free = np.ones(b[4].size, bool)
        cdof = prob.constraints()[0]
        free[cdof] = False
        res, bwd = {}, {}
        for s in range(5):
            A = prob.matrix(s)
            r = b[s] - A @ x[s]
            if s == 4:
                r = r[free]
            name = f"species_{s}" if s < 4 else "poisson"
            res[name] = float(np.abs(r).max() / np.abs(b[s]).max())
            bwd[name] = float(np.abs(r).max() / (abs(A).sum(axis=1).max() * np.abs(x[s]).max() + np.abs(b[s]).max()))
        out["residual_rel"] = max(res.values())
        out["residual_rel_per_system"] = res
        out["backward_err"] = max(bwd.values())
        out["n_dofs_per_carrier"] = int(b[0].size)
    finally:
        for s in range(5):
            prob.set_solution(s, saved[s])
        if oracle is None:
            o.close()
    out["seconds"] = time.perf_counter() - t0
    return out
