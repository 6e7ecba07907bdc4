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
