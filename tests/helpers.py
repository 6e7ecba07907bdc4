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


def csr_matvec_longdouble(A, x):
    """A @ x with products and row sums in long double (80-bit on x86): residuals of the refinements below"""
    A = A.tocsr()
    prod = A.data.astype(np.longdouble) * np.asarray(x, np.longdouble)[A.indices]
    out = np.zeros(A.shape[0], np.longdouble)
    nonempty = np.diff(A.indptr) > 0
    out[nonempty] = np.add.reduceat(prod, A.indptr[:-1][nonempty])
    return out


def refine_through_abi(prob, which, A, b, x0, free=None, sweeps=2):
    """Solution of A x = b to working accuracy WITHOUT a second solver: iterative refinement with long-double residuals
    whose correction solves go through the device solver itself (pecs_set_rhs / pecs_solve_*): if the device solve is
    accurate to 1e-6 or better, two sweeps leave an error of 1e-12 times the first one's.  What the FIRST solve was off by
    is then simply x0 - x.  (At the benchmarked size no CPU LU is affordable; this is how its accuracy is measured.)"""
    x = np.asarray(x0, np.longdouble).copy()
    zero = np.zeros(b.size)
    for _ in range(sweeps):
        r = np.asarray(b, np.longdouble) - csr_matvec_longdouble(A, x)
        if free is not None:
            r[~free] = 0.0
        prob.set_rhs(which, r.astype(np.float64))
        prob.set_solution(which, zero)
        if which == 4:
            prob.solve_Poisson()
        else:
            prob.solve_species(which)
        dx = prob.get_solution(which)
        if free is not None:
            dx[~free] = 0.0
        x = x + dx
    return x


def validate_workload(prob, oracle=None, seed=1234, refine=True):
    """Parity of ONE pass of the hot path on whatever mesh `prob` holds (used by bench.py at the benchmarked
    configuration and by tests/test_gpu_large.py), outside any timed region.

    First the step the caller ran LAST (the right-hand sides and states are still on the device):
      last_step_residual_rel  max over the five systems of |b - A x|_inf / |b|_inf with the host CSR matrices
    then, from a perturbed state (seed 1234):
      rhs_rel      all five assembled right-hand sides against the oracle's assembly of the same state, worst block,
                   max-norm relative (north_star: 1e-12).  The oracle is built WITHOUT its LU (factor=False).
      residual_rel max over the five systems of |b - A x|_inf / |b|_inf (scipy mat-vec), x = what pecs_solve_* returned
                   for right-hand side b: reference Carrier.cpp:34-40 semantics x = A^-1 b.  Constrained Poisson rows
                   (hanging / Neumann fluxes) are excluded: the solver eliminates them, `distribute` overwrites them.
      backward_err the same residuals normwise: |r|_inf / (|A|_inf |x|_inf + |b|_inf)
      solve_density_err / solve_current_err / solve_potential_err / solve_field_err
                   error of that ONE solve per block against the refined solution of the same system
                   (refine_through_abi), max-norm relative: the quantity north_star bounds by 1e-9 (densities, potential)
      finite       every state entry is finite
    The caller's state is restored."""
    import time
    t0 = time.perf_counter()
    saved = [prob.get_solution(s) for s in range(5)]
    saved_rhs = [prob.get_rhs(s) for s in range(5)]
    o = oracle if oracle is not None else make_oracle(prob, True, factor=False)
    out = {}
    free = np.ones(saved[4].size, bool)
    free[prob.constraints()[0]] = False
    names = [f"species_{s}" for s in range(4)] + ["poisson"]
    mats = [prob.matrix(s) for s in range(5)]

    def residuals(b, x):
        res, bwd = {}, {}
        for s in range(5):
            r = b[s] - mats[s] @ x[s]
            if s == 4:
                r = r[free]
            res[names[s]] = float(np.abs(r).max() / np.abs(b[s]).max())
            bwd[names[s]] = float(np.abs(r).max() /
                                  (abs(mats[s]).sum(axis=1).max() * np.abs(x[s]).max() + np.abs(b[s]).max()))
        return res, bwd

    try:
        if all(np.abs(v).max() > 0 for v in saved_rhs):
            res, _ = residuals(saved_rhs, saved)
            out["last_step_residual_rel"] = max(res.values())
            out["last_step_residual_rel_per_system"] = res
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
        per = {names[s]: block_rel_err(b[s], o.rhs(s)) for s in range(4)}
        per["poisson"] = rel_err(b[4], o.rhs(4))
        out["rhs_rel"] = max(per.values())
        out["rhs_rel_per_vector"] = per
        prob.solve_full_system()
        prob.solve_Poisson()
        x = [prob.get_solution(s) for s in range(5)]
        out["finite"] = bool(all(np.isfinite(v).all() for v in x))
        res, bwd = residuals(b, x)
        out["residual_rel"] = max(res.values())
        out["residual_rel_per_system"] = res
        out["backward_err"] = max(bwd.values())
        out["n_dofs_per_carrier"] = int(b[0].size)
        if refine:
            dens, cur = {}, {}
            for s in range(4):
                ref = refine_through_abi(prob, s, mats[s], b[s], x[s]).astype(np.float64)
                nc = ref.size // 12
                dens[names[s]] = float(rel_err(x[s][8 * nc:], ref[8 * nc:]))
                cur[names[s]] = float(max(rel_err(x[s][:4 * nc], ref[:4 * nc]), rel_err(x[s][4 * nc:8 * nc], ref[4 * nc:8 * nc])))
            ref = refine_through_abi(prob, 4, mats[4], b[4], x[4], free).astype(np.float64)
            n_rt = prob.n_rt
            rt_free = free[:n_rt]
            out["solve_density_err"] = max(dens.values())
            out["solve_current_err"] = max(cur.values())
            out["solve_potential_err"] = float(rel_err(x[4][n_rt:], ref[n_rt:]))
            out["solve_field_err"] = float(rel_err(x[4][:n_rt][rt_free], ref[:n_rt][rt_free]))
            out["solve_density_err_per_species"] = dens
            out["solve_current_err_per_species"] = cur
    finally:
        for s in range(5):
            prob.set_solution(s, saved[s])
            prob.set_rhs(s, saved_rhs[s])
        if oracle is None:
            o.close()
    out["seconds"] = time.perf_counter() - t0
    return out
