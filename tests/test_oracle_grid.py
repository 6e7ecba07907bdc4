"""The oracle-side mesh generator and parameter scaling (oracle/grid.py, numpy, closed-form index arithmetic) against
the product's host classes (own quadtree, host/Triangulation.cpp + host/Grid.cpp + host/Parameters.cpp): two
independent restatements of reference source/Grid.cpp and include/Parameters.hpp must produce the same tables, and
both must reproduce the sizes the reference prints (derived from Grid.cpp:44-133).  CPU only."""
import numpy as np
import pytest

import pecs_b200 as pecs
from oracle import grid as ogrid

CASES = [
    (2, 0, {}),
    (2, 1, {}),
    (3, 1, {"physical__insulated": False, "physical__applied_bias": 0.1}),
    (3, 1, {"physical__schottky_status": False}),
    (4, 1, {}),
    (3, 1, {"mesh__radius_one": 0.2, "mesh__radius_two": 0.6}),
]


def to_prm(g, l, overrides):
    names = {"physical__insulated": "insulated", "physical__applied_bias": "applied bias",
             "physical__schottky_status": "schottky status", "mesh__radius_one": "radius one",
             "mesh__radius_two": "radius two"}
    prm = {"global refinements": g, "local refinements": l}
    prm.update({names[k]: v for k, v in overrides.items()})
    return prm


@pytest.mark.parametrize("g,l,overrides", CASES)
def test_oracle_grid_equals_product_tables(g, l, overrides):
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **overrides))
    prob.setup_full_system_host()
    meshes = ogrid.make_grids(to_prm(g, l, overrides), True)
    for which, mine in enumerate(meshes):
        theirs = prob.mesh(which)
        assert mine["n_cells"] == theirs["n_cells"]
        np.testing.assert_allclose(mine["vertices"], theirs["vertices"], rtol=0, atol=4e-16)
        for key in ("material_id", "level", "face_kind", "neighbor", "neighbor2", "boundary_id"):
            assert np.array_equal(mine[key], theirs[key]), (which, key)
        np.testing.assert_allclose(mine["nb_parent_diameter"], theirs["nb_parent_diameter"], rtol=1e-15, atol=0)
    np.testing.assert_allclose(ogrid.scaled_parameters(to_prm(g, l, overrides))[:29], prob.params[:29], rtol=1e-15)
    prob.close()


@pytest.mark.parametrize("g,l,cells,poisson_cells,rt_dofs,interface_faces", [
    (4, 1, 1280, 2560, 5280, 32), (5, 1, 5120, 10240, 20800, 64), (6, 1, 20480, 40960, 82560, 128)])
def test_sizes_of_survey_appendix_d(g, l, cells, poisson_cells, rt_dofs, interface_faces):
    """cells per subdomain = 4^g + 4^(g+l); RT0 dofs = 4N(N+1) + 4M(M+1) - M, N = 2^g, M = 2^(g+l) (parent and child
    edges both carry dofs on the two hanging lines); interface faces = 2^(g+l)"""
    semi, elec, poisson = ogrid.make_grids({"global refinements": g, "local refinements": l})
    assert semi["n_cells"] == cells == elec["n_cells"] and poisson["n_cells"] == poisson_cells
    assert int((semi["boundary_id"] == ogrid.INTERFACE).sum()) == interface_faces
    assert int((elec["boundary_id"] == ogrid.INTERFACE).sum()) == interface_faces
    # count edges of the Poisson mesh: every face once (same-level pairs counted once, hanging: parent + 2 children)
    fk = poisson["face_kind"]
    edges = (fk == ogrid.FACE_BOUNDARY).sum() + (fk == ogrid.FACE_SAME_LEVEL).sum() // 2 + \
        (fk == ogrid.FACE_HAS_CHILDREN).sum() + (fk == ogrid.FACE_COARSER).sum()
    assert int(edges) == rt_dofs


def test_default_scaled_parameters_match_the_survey():
    """SURVEY App. D, computed from include/Parameters.hpp:181-242 on input_file.prm"""
    p = dict(zip(ogrid.PARAM_SLOTS, ogrid.scaled_parameters()))
    assert p["mu_n"] == pytest.approx(3.48975e-3, rel=1e-6) and p["mu_p"] == pytest.approx(1.2408e-3, rel=1e-6)
    assert p["mu_r"] == pytest.approx(2.585e-6, rel=1e-6) and p["mu_o"] == p["mu_r"]
    assert p["lambda2"] == pytest.approx(1.41218e-4, rel=1e-5)
    assert p["k_et"] == pytest.approx(1e-11, rel=1e-12) and p["k_ht"] == pytest.approx(1e-6, rel=1e-12)
    assert p["v_n"] == pytest.approx(3.0e-3) and p["v_p"] == pytest.approx(2.9e-3)
    assert p["gen_flux"] == pytest.approx(1.2e-11) and p["gen_alpha"] == pytest.approx(17.4974)
    assert p["phi_bi"] == pytest.approx(15.8607, rel=1e-5) and p["delta_t"] == 0.05


def test_oracle_runs_without_the_product_library():
    """bench.py's CPU arm: the oracle steps on its own grid; same states as the oracle fed with the product's tables"""
    from helpers import make_oracle, rel_err
    o = ogrid.make_oracle({"global refinements": 2, "local refinements": 1})
    o.setup(1.0, True)
    o.project_initial_conditions()
    o.assemble_Poisson_rhs()
    o.solve_Poisson()
    o.step(3)
    prob = pecs.SolarCellProblem(pecs.default_input_file(2, 1))
    prob.setup_full_system_host()
    o2 = make_oracle(prob, True)
    o2.project_initial_conditions()
    o2.assemble_Poisson_rhs()
    o2.solve_Poisson()
    o2.step(3)
    for s in range(5):
        assert rel_err(o.solution(s), o2.solution(s)) <= 1e-12
    prob.close()
