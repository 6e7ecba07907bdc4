"""Host logic on CPU: the nested-dissection plan and the factor tables (host numeric factorisation) solve the
constant systems; the oracle's own sparse LU is cross-checked with SuperLU (scipy) as a third, unrelated solver."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import pecs_b200 as pecs
from helpers import make_oracle


@pytest.fixture(scope="module")
def problem():
    prob = pecs.SolarCellProblem(pecs.default_input_file(3, 1))
    prob.setup_full_system_host()
    return prob


@pytest.mark.parametrize("which", [0, 1, 2, 4])
@pytest.mark.parametrize("leaf", [0, 1, 9])
def test_plan_and_tables_solve_the_system(problem, which, leaf):
    A = problem.matrix(which)
    rng = np.random.default_rng(which)
    b = rng.standard_normal(A.shape[0])
    x = problem.selftest_direct_solve(which, b, leaf)
    assert np.linalg.norm(A @ x - b) <= 1e-10 * np.linalg.norm(b)
    st = problem.plan_stats(which, leaf)
    assert st["fwd_entries"] > 0 and st["levels"] >= 3


def test_poisson_plan_never_needs_cross_front_pivoting():
    """the saddle-point matrix has a zero (Phi, Phi) block: pairing each potential with fluxes of its own cell
    keeps every pivot block invertible (host/SolverSetup.hpp) -- on the hanging-node mesh and on the test grids"""
    for prm, test_defaults, setup in ((pecs.default_input_file(4, 1), False, None), (None, True, 4)):
        prob = pecs.SolarCellProblem(prm, test_defaults=test_defaults)
        if setup is None:
            prob.setup_full_system_host()
        else:
            prob.setup_test_host(pecs.KIND_TEST_DD_POISSON, setup)
        A = prob.matrix(pecs.POISSON)
        b = np.ones(A.shape[0])
        x = prob.selftest_direct_solve(pecs.POISSON, b)
        assert np.linalg.norm(A @ x - b) <= 1e-9 * np.linalg.norm(b)


def test_oracle_lu_against_superlu(problem):
    o = make_oracle(problem, True)
    rng = np.random.default_rng(7)
    for which in (0, 3, 4):
        A = o.matrix(which).tocsc()
        b = rng.standard_normal(A.shape[0])
        o.set_vector(which, 1, b)
        if which == 4:
            # PoissonData::solve also distributes the constraints; compare the unconstrained rows
            o.solve_Poisson()
        else:
            o.solve_species(which)
        x_ref = spla.splu(A).solve(b)
        x = o.solution(which)
        if which == 4:
            dof, _, _ = problem.constraints()
            keep = np.ones(A.shape[0], bool)
            keep[dof] = False
            assert np.abs(x[keep] - x_ref[keep]).max() <= 1e-9 * np.abs(x_ref).max()
        else:
            assert np.abs(x - x_ref).max() <= 1e-10 * np.abs(x_ref).max()


@pytest.mark.parametrize("g,l,overrides", [
    (3, 2, {"mesh__radius_one": 0.2, "mesh__radius_two": 0.6}),   # conic wire, two interface refinements (2:1 smoothing)
    (3, 0, {}),                                                   # no hanging faces
    (2, 1, {"physical__insulated": False, "physical__applied_bias": 0.2, "physical__schottky_status": False})])
def test_plans_on_the_other_configurations(g, l, overrides):
    """unknown-level separators (carriers) and edge-flux separators with delayed potentials (Poisson) on the meshes of
    the other BASELINE configurations: plan + host factor tables solve every constant system"""
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **overrides))
    prob.setup_full_system_host()
    for which in range(5):
        A = prob.matrix(which)
        b = np.random.default_rng(which).standard_normal(A.shape[0])
        x = prob.selftest_direct_solve(which, b)
        assert np.linalg.norm(A @ x - b) <= 1e-9 * np.linalg.norm(b)
    fronts = prob.plan_fronts(pecs.POISSON)
    assert (fronts[:, 1] >= 0).all() and fronts[:, 1].sum() == prob.matrix(pecs.POISSON).shape[0]


def test_host_preparation_does_not_depend_on_the_thread_count(monkeypatch):
    """Schur reduction, nested dissection, symbolic plan and the permuted matrices are built on several threads at setup
    (SolverSetup.cpp: preparation_threads); every table must come out bit-identical whatever the number of threads."""
    prob = pecs.SolarCellProblem(pecs.default_input_file(4, 1))
    prob.setup_full_system_host()
    try:
        for which in (pecs.ELECTRONS, pecs.OXIDANTS, pecs.POISSON):
            seen = {}
            for threads in (1, 2, 3, 8):
                monkeypatch.setenv("PECS_B200_SETUP_THREADS", str(threads))
                seen[threads] = prob.selftest_prepared_hashes(which)
            assert len(set(seen.values())) == 1, (which, seen)
            assert seen[1][4] != 0 and seen[1][6] != seen[1][7]
    finally:
        prob.close()


@pytest.mark.parametrize("table,cols", [(0, 4), (1, 8), (2, 8), (3, 4)])
def test_ell_tables_reproduce_the_csr_product(problem, table, cols):
    """the ELL layout the device kernels stream (groups of four columns per slot where that is smaller) is the matrix:
    S, T1, A^-1 and T2 of a carrier, with the kernel's own slot-by-slot arithmetic, against the CSR product"""
    n = problem.n_cells(0)
    x = np.random.default_rng(11 + table).standard_normal(cols * n)
    y_ell, y_csr, (rows, width, block) = problem.selftest_ell_matvec(pecs.HOLES, table, x)
    assert rows == y_csr.size and width > 0 and block in (1, 4)
    assert abs(y_ell - y_csr).max() <= 1e-13 * abs(y_csr).max()
