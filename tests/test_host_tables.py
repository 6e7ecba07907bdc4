"""Host logic on CPU: meshes, dof numbering, mappings and the constant matrices built by pecs_b200/csrc/host,
checked against SURVEY App. D counts, against the independently written oracle, and through structural identities."""
import os

import numpy as np
import pytest

import pecs_b200 as pecs
from helpers import make_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def default_problem():
    prob = pecs.SolarCellProblem(pecs.default_input_file(4, 1))
    prob.setup_full_system_host()
    return prob, make_oracle(prob, True, factor=False)


@pytest.mark.parametrize("g,l,cells,rt,pairs", [(4, 1, 1280, 5280, 32), (3, 1, 320, 1360, 16), (4, 0, 256, None, 16),
                                                (3, 2, None, None, 32)])
def test_mesh_counts(g, l, cells, rt, pairs):
    """cells/subdomain = 4^g + 4^(g+l), RT0 count incl. parent and child edges on hanging lines"""
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l))
    prob.setup_full_system_host()
    if cells is not None:
        assert prob.n_cells(0) == cells and prob.n_cells(1) == cells and prob.n_cells(2) == 2 * cells
    if rt is not None:
        assert prob.n_rt == rt
    assert len(prob.interface_pairs()[0]) == pairs
    for which in range(3):
        m = prob.mesh(which)
        # 2:1 balance (deal.II smoothing): neighbours differ by at most one level
        for c in range(m["n_cells"]):
            for f in range(4):
                k = m["face_kind"][c, f]
                if k == 0:
                    assert m["level"][m["neighbor"][c, f]] == m["level"][c]
                elif k == 2:
                    assert m["level"][m["neighbor"][c, f]] == m["level"][c] + 1
                    assert m["level"][m["neighbor2"][c, f]] == m["level"][c] + 1
                elif k == 3:
                    assert m["level"][m["neighbor"][c, f]] == m["level"][c] - 1
        # positive Jacobian at the vertices
        v = m["vertices"]
        det = (v[:, 1, 0] - v[:, 0, 0]) * (v[:, 2, 1] - v[:, 0, 1]) - (v[:, 2, 0] - v[:, 0, 0]) * (v[:, 1, 1] - v[:, 0, 1])
        assert (det > 0).all()


def test_boundary_tagging_default_input(default_problem):
    """reference Grid.cpp:340-461 incl. the radius-one quirk on the bottom edge"""
    prob, _ = default_problem
    m = prob.mesh(0)
    on_bdry = m["face_kind"] == 1
    ids = m["boundary_id"][on_bdry]
    assert set(np.unique(ids)) == {0, 1, 2, 3}  # interface, Dirichlet, Neumann, Schottky all present
    assert (ids == 0).sum() == 32
    # bottom faces with 0.3 < x < 0.6 are Neumann, x < 0.3 Dirichlet
    v = m["vertices"]
    for c in range(m["n_cells"]):
        if m["face_kind"][c, 2] == 1 and v[c, 0, 1] == 0.0:
            xc = 0.5 * (v[c, 0, 0] + v[c, 1, 0])
            assert m["boundary_id"][c, 2] == (2 if xc > 0.3 else 1)


def test_dofs_and_maps_match_oracle(default_problem):
    prob, o = default_problem
    assert o.n_rt == prob.n_rt
    assert (o.poisson_face_dofs(prob.n_cells(2)) == prob.poisson_face_dofs()).all()
    assert (o.cell_map(0, prob.n_cells(0)) == prob.cell_map(0)).all()
    assert (o.cell_map(1, prob.n_cells(1)) == prob.cell_map(1)).all()
    dof, master, w = prob.constraints()
    assert (w[master >= 0] == 0.5).all() and len(dof) > 0


@pytest.mark.parametrize("which", range(7))
def test_matrices_match_oracle(default_problem, which):
    """block-form host assembly vs the oracle's FEValues loops (hanging faces included: l = 1)"""
    prob, o = default_problem
    A, B = prob.matrix(which), o.matrix(which)
    assert A.nnz == B.nnz
    assert abs(A - B).max() <= 1e-14 * abs(B).max()  # measured 4e-16 .. 2e-15 (VERDICT r1 item 7 asks for 1e-14)


def test_carrier_matrix_structure(default_problem):
    """symmetric part = diag(mu^-1 A, M/dt + penalty); B blocks are skew"""
    prob, _ = default_problem
    A = prob.matrix(0).tocsr()
    n = A.shape[0] // 3
    S = 0.5 * (A + A.T)
    off = S[: 2 * n, 2 * n:]
    assert abs(off).max() <= 1e-13 * abs(A).max()
    # current-current block is the cell mass matrix / mu: symmetric positive definite, block diagonal
    JJ = A[: 2 * n, : 2 * n]
    assert abs(JJ - JJ.T).max() <= 1e-13 * abs(JJ).max()
    assert JJ.diagonal().min() > 0


def test_mass_matrix_row_sums(default_problem):
    """sum of M = |Omega| / dt (partition of unity)"""
    prob, _ = default_problem
    M = prob.matrix(5)
    area = 0.5 * (0.3 + 0.6) * 1.0  # trapezoid semiconductor
    assert abs(M.sum() * prob.delta_t - area) <= 1e-12


def test_oracle_matches_golden():
    """the committed fixtures are oracle outputs: this notices any change of the oracle (and of the host tables)"""
    for name, g, overrides in (("production_g2_l1.npz", 2, {}),
                               ("production_g3_l1_biased.npz", 3,
                                {"physical__insulated": False, "physical__applied_bias": 0.1})):
        gold = np.load(os.path.join(GOLDEN, name))
        prob = pecs.SolarCellProblem(pecs.default_input_file(g, 1, **overrides))
        prob.setup_full_system_host()
        o = make_oracle(prob, True)
        o.project_initial_conditions()
        o.assemble_Poisson_rhs()
        o.solve_Poisson()
        scale = abs(gold["poisson_solution_initial"]).max()
        assert abs(o.solution(4) - gold["poisson_solution_initial"]).max() <= 1e-11 * max(scale, 1e-300)
        o.step(int(gold["n_steps"]))
        for s in range(5):
            ref = gold[f"state_after_steps_{s}"]
            assert abs(o.solution(s) - ref).max() <= 1e-10 * abs(ref).max()


def test_ldg_matrices_do_not_depend_on_the_thread_count(tmp_path):
    """host/LDG.cpp assembles the system matrices cell by cell in gather form (every cell evaluates all terms of its own
    12 rows, in the order a sequential face loop would insert them), the cells in parallel: bit-identical whatever the
    number of threads"""
    import subprocess
    import sys
    code = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import pecs_b200 as pecs\n"
        "p = pecs.SolarCellProblem(pecs.default_input_file(3, 2, mesh__radius_one=0.2))\n"
        "p.setup_full_system_host()\n"
        "h = hashlib.sha256()\n"
        "for w in range(4):\n"
        "    m = p.matrix(w)\n"
        "    for a in (m.indptr, m.indices, m.data):\n"
        "        h.update(np.ascontiguousarray(a).tobytes())\n"
        "print(h.hexdigest())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    digests = []
    for threads in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=threads)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
        digests.append(out.stdout.strip().splitlines()[-1])
    assert digests[0] == digests[1] == digests[2]


@pytest.mark.parametrize("g,l,overrides", [
    (3, 1, {}), (3, 2, {"mesh__radius_one": 0.2}),
    (2, 1, {"physical__illumination_status": False, "physical__schottky_status": False, "physical__insulated": False,
            "physical__applied_bias": 0.2}),
    # Shockley-Read-Hall recombination switched on (the formula the reference carries as a comment, SolarCell.hpp:86-98)
    # with lifetimes and intrinsic density that make it an O(1) term of the right-hand side
    (3, 1, {"physical__srh_recombination": True, "physical__intrinsic_density": 0.5e16,
            "electrons__recombination_time": 2e-12, "holes__recombination_time": 1e-12})])
def test_production_rhs_arithmetic_matches_oracle_on_cpu(g, l, overrides):
    """The arithmetic of the production RHS kernels -- csrc/rhs_math.hpp: static generation integrals, sum-factorised
    cell terms, static face geometry and the Dirichlet / interface / Schottky face terms, the very inline functions
    the CUDA kernels call -- evaluated on the CPU against the oracle's assembly from a perturbed state: all four
    carrier right-hand sides, 1e-12 relative per block (the tolerance of the GPU parity tests).  What only the GPU
    tests can cover is the kernels' data movement: loads, stores, launch shapes, boundary blocks."""
    from helpers import block_rel_err, perturbed
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **overrides))
    prob.setup_full_system_host()
    o = make_oracle(prob, True)
    o.project_initial_conditions()
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    for s in range(4):
        o.set_vector(s, 0, perturbed(o.solution(s), 4321 + s))
    o.set_vector(4, 0, perturbed(o.solution(4), 77))
    o.assemble_semiconductor_rhs()
    o.assemble_electrolyte_rhs()
    u, X = [o.solution(s) for s in range(4)], o.solution(4)
    for w in range(2):
        got = prob.selftest_carrier_rhs(w, u[2 * w], u[2 * w + 1], X, u[2 - 2 * w], u[3 - 2 * w])
        for k in range(2):
            assert block_rel_err(got[k], o.rhs(2 * w + k)) <= 1e-12, (w, k)
        # without the other subdomain's vectors: cell terms only, equal on the cells without boundary faces
        cells_only = prob.selftest_carrier_rhs(w, u[2 * w], u[2 * w + 1], X)
        interior = ~(prob.mesh(w)["face_kind"] == 1).any(axis=1)  # FACE_BOUNDARY == 1
        n = interior.size
        assert 0 < interior.sum() < n
        for k in range(2):
            a, b = cells_only[k].reshape(3, n, 4), got[k].reshape(3, n, 4)
            assert np.array_equal(a[:, interior], b[:, interior]) and not np.array_equal(a, b)
    prob.close()


def test_poisson_rows_and_output_field_arithmetic_on_cpu():
    """same idea for the other two device formulas of the path: the Poisson charge rows on the static int N_a table
    against the oracle's Poisson assembly (1e-12), and the RT0 field at the patch vertices of the output path against
    the numpy restatement in oracle/output.py (1e-14)"""
    from helpers import perturbed, rel_err
    from oracle import output as oracle_output
    prob = pecs.SolarCellProblem(pecs.default_input_file(3, 2, mesh__radius_one=0.2))
    prob.setup_full_system_host()
    o = make_oracle(prob, True)
    o.project_initial_conditions()
    for s_ in range(4):
        o.set_vector(s_, 0, perturbed(o.solution(s_), 99 + s_))
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    n_rt = prob.n_rt
    rows = prob.selftest_poisson_rows([o.solution(s_) for s_ in range(4)])
    assert rel_err(rows, o.rhs(4)[n_rt:]) <= 1e-12
    X = o.solution(4)
    got = prob.selftest_field_patches(X, 2.5)
    want, _ = oracle_output.poisson_patches(prob.mesh(2)["vertices"], prob.poisson_face_dofs(), n_rt, X, 2.5, 1.0)
    assert np.abs(got - want[:, :2]).max() <= 1e-14 * np.abs(want).max()
    prob.close()
