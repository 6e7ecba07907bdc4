"""oracle/cpu_arm.py -- TEST / MEASUREMENT INFRASTRUCTURE ONLY: the CPU arm of bench.py.

Times the reference's per-step path on the host cores, restated WITHOUT the product library (nothing here imports
pecs_b200; `--impl reference` must not load libpecs_b200.so):

  * mesh and scaled parameters: oracle/grid.py (reference source/Grid.cpp, include/Parameters.hpp)
  * dofs, maps, the constant matrices and the five per-step assembly passes: the C++ oracle (oracle/pecs_oracle.cpp,
    OpenMP cell loops with an ordered scatter = the reference's WorkStream, reference SolarCell.cpp:430-815, 1037-1726)
  * the five fixed-matrix direct solves: SuperLU (scipy.sparse.linalg.splu, COLAMD ordering), factorised once, one
    substitution per system and step -- the stand-in for deal.II's SparseDirectUMFPACK (reference Carrier.cpp:26-40,
    Poisson.cpp:92-105), which is absent here.  Like UMFPACK it is a supernodal sparse LU with a fill-reducing column
    ordering and partial pivoting.  The four carrier substitutions run concurrently on four threads (the reference's
    four TBB tasks, SolarCell.cpp:1763-1781), the Poisson one on one thread.  The oracle's own Gilbert-Peierls LU
    (oracle/sparse_lu.hpp, the parity checker) is a scalar left-looking code: 156 s to factorise at g=5 where SuperLU
    takes 2.4 s, and hours at g=6 -- it cannot reach the benchmarked size, SuperLU can (measured in the build
    container, one thread per system: g=5 factor 2.8 s / substitution 0.06 s, g=6 23.5 s / 0.30 s, fill x5.9 per level).
"""
import concurrent.futures as cf
import os
import time

import numpy as np

from . import grid as ogrid

SECTIONS = ["Assemble semiconductor rhs", "Assemble electrolyte rhs", "Solve LDG Systems", "Assemble Poisson rhs",
            "Solve Poisson system"]


def cells_per_subdomain(g, l):
    return 4 ** g + (4 ** (g + l) if l > 0 else 0)


def available_memory_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return None


class CpuReferencePath:
    """oracle assembly + SuperLU solves on one mesh; setup() is untimed (the reference's set_solvers), step() timed"""

    def __init__(self, g, l, prm=None, threads=None):
        self.g, self.l = g, l
        self.threads = int(threads or os.cpu_count())
        os.environ["OMP_NUM_THREADS"] = str(self.threads)  # never inherit torchrun's OMP_NUM_THREADS=1
        p = {"global refinements": g, "local refinements": l}
        p.update(prm or {})
        self.prm = p
        self.o = None
        self.lu = [None] * 5
        self.setup_seconds = {}

    def setup(self):
        import scipy.sparse.linalg as spl
        t0 = time.perf_counter()
        self.o = ogrid.make_oracle(self.prm)
        self.o.setup(1.0, False)  # dofs, maps, matrices; no oracle LU
        t1 = time.perf_counter()
        mats = [self.o.matrix(s).tocsc() for s in range(5)]
        t2 = time.perf_counter()

        def factor(s):
            self.lu[s] = spl.splu(mats[s], permc_spec="COLAMD")

        with cf.ThreadPoolExecutor(5) as ex:  # SuperLU releases the GIL: five concurrent factorisations
            list(ex.map(factor, range(5)))
        t3 = time.perf_counter()
        self.setup_seconds = {"oracle_setup": t1 - t0, "matrices": t2 - t1, "factor": t3 - t2}
        self.factor_nnz = [int(lu.L.nnz + lu.U.nnz) for lu in self.lu]
        self.pool = cf.ThreadPoolExecutor(4)
        o = self.o
        o.project_initial_conditions()
        o.assemble_Poisson_rhs()
        self._solve_poisson()
        return self

    def _solve_poisson(self):
        o = self.o
        o.set_vector(4, 0, self.lu[4].solve(o.rhs(4)))
        o.distribute_Poisson()

    def step(self, n=1):
        """n IMEX steps in the reference's order (SolarCell.cpp:2057-2075); returns section seconds"""
        o, sec = self.o, np.zeros(5)
        for _ in range(n):
            t = time.perf_counter()
            o.assemble_semiconductor_rhs()
            t1 = time.perf_counter()
            o.assemble_electrolyte_rhs()
            t2 = time.perf_counter()
            rhs = [o.rhs(s) for s in range(4)]
            sols = list(self.pool.map(lambda s: self.lu[s].solve(rhs[s]), range(4)))
            for s in range(4):
                o.set_vector(s, 0, sols[s])
            t3 = time.perf_counter()
            o.assemble_Poisson_rhs()
            t4 = time.perf_counter()
            self._solve_poisson()
            t5 = time.perf_counter()
            sec += [t1 - t, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
        return sec

    def states(self):
        return [self.o.solution(s) for s in range(5)]

    def close(self):
        if getattr(self, "pool", None):
            self.pool.shutdown()
        if self.o is not None:
            self.o.close()
            self.o = None
        self.lu = [None] * 5


def run(g, l, steps, warmup, threads=None, prm=None):
    """steps/s of the CPU path on refinement g, with its section times and setup cost"""
    path = CpuReferencePath(g, l, prm, threads).setup()
    path.step(max(warmup, 1))
    t0 = time.perf_counter()
    sections = path.step(steps)
    dt = time.perf_counter() - t0
    finite = bool(all(np.isfinite(v).all() for v in path.states()))
    out = {"g": g, "l": l, "cells_per_subdomain": cells_per_subdomain(g, l), "dofs_per_carrier": 12 * cells_per_subdomain(g, l),
           "steps": steps, "steps_per_s": steps / dt, "seconds_per_step": dt / steps, "threads": path.threads,
           "section_seconds_per_step": dict(zip(SECTIONS, (sections / steps).tolist())),
           "setup_seconds": path.setup_seconds, "factor_nnz": path.factor_nnz, "finite": finite}
    path.close()
    return out
