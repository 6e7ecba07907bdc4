"""Output path: the reference's print_results (source/SolarCell.cpp:1826-1858) = DataOut patches
+ PostProcessor rescaling + write_vtu.  CPU tests cover the host half (names, scales, file layout); the GPU tests
compare the device-made patch values with the numpy restatement in oracle/output.py and read the files back."""
import os

import numpy as np
import pytest

import pecs_b200 as pecs
from oracle import output as oracle_output
from pecs_b200.vtu import read_vtu


def _problem(g=2, l=1):
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l))
    prob.setup_full_system_host()
    return prob


def test_scales_match_postprocessor_formulas():
    prob = _problem()
    # default input: characteristic length 1e-4, density 1e16, time 1e-12 (reference input_file.prm)
    expect = oracle_output.scales(1.0e-4, 1.0e16, 1.0e-12)
    assert np.allclose(prob.output_scales, expect, rtol=1e-15, atol=0)
    assert expect[3] == 1.6e-19 * 1.0e16 * 1.0e-4 / 1.0e-12
    prob.close()


@pytest.mark.parametrize("which,name,fields", [
    (0, "Semiconductor-007.vtu", [("Electrons Current", 3), ("Electrons Density", 1), ("Holes Current", 3), ("Holes Density", 1)]),
    (1, "Electrolyte-007.vtu", [("Reductants Current", 3), ("Reductants Density", 1), ("Oxidants Current", 3), ("Oxidants Density", 1)]),
    (2, "Poisson-007.vtu", [("Field", 3), ("Potential", 1)])])
def test_vtu_files_round_trip(tmp_path, which, name, fields):
    """file names and field names of the reference (LDG.cpp:1195-1232, MixedFEM.cpp:297-320, PostProcessor.cpp:20-55),
    one 4-vertex patch per cell, values bit-exact through the base64 stream"""
    prob = _problem()
    n = prob.n_cells(which)
    rng = np.random.default_rng(5 + which)
    patches = rng.standard_normal((32 if which < 2 else 16) * n)
    prob.write_patches(which, patches, 7, str(tmp_path))
    f = read_vtu(os.path.join(tmp_path, name))
    assert f["n_cells"] == n and f["n_points"] == 4 * n
    verts = prob.mesh(which)["vertices"]
    assert np.array_equal(f["points"][:, :2], verts.reshape(-1, 2)) and not f["points"][:, 2].any()
    base = 4 * np.arange(n)[:, None]
    assert np.array_equal(f["connectivity"].reshape(n, 4), base + np.array([0, 1, 3, 2]))  # VTK_QUAD from lexicographic
    assert np.array_equal(f["offsets"], 4 * (np.arange(n) + 1)) and set(f["types"]) == {9}
    assert list(f["point_data"]) == [k for k, _ in fields]
    off = 0
    for key, comps in fields:
        a = f["point_data"][key]
        assert np.array_equal(a.ravel(), patches[off:off + 4 * n * comps])
        off += 4 * n * comps
    # every quad is counter-clockwise in the VTK order
    p = f["points"][f["connectivity"].reshape(n, 4)]
    area2 = sum(p[:, i, 0] * p[:, (i + 1) % 4, 1] - p[:, (i + 1) % 4, 0] * p[:, i, 1] for i in range(4))
    assert (area2 > 0).all()
    prob.close()


def test_print_results_needs_the_device():
    """no CPU fallback: without a context print_results fails loudly"""
    prob = _problem()
    with pytest.raises(pecs.PecsError):
        prob.print_results(0)
    prob.close()


def _expected_patches(prob):
    sc = prob.output_scales
    out = []
    for w, (a, b) in enumerate([(pecs.ELECTRONS, pecs.HOLES), (pecs.REDUCTANTS, pecs.OXIDANTS)]):
        c1, d1 = oracle_output.carrier_patches(prob.get_solution(a), sc[3])
        c2, d2 = oracle_output.carrier_patches(prob.get_solution(b), sc[3])
        out.append({"current_1": c1, "density_1": d1, "current_2": c2, "density_2": d2})
    field, potential = oracle_output.poisson_patches(prob.mesh(2)["vertices"], prob.poisson_face_dofs(), prob.n_rt,
                                                     prob.get_solution(pecs.POISSON), sc[1], sc[0])
    out.append({"field": field, "potential": potential})
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("g,l,overrides", [(3, 1, {}), (3, 2, {"mesh__radius_one": 0.2})])
def test_snapshot_matches_output_oracle(g, l, overrides):
    """device patch values after a few steps against the numpy restatement of DataOut + PostProcessor"""
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **overrides))
    prob.setup_full_system()
    prob.step(3)
    got, want = prob.output_snapshot(), _expected_patches(prob)
    for g_, w_ in zip(got, want):
        for key in w_:
            scale = np.abs(w_[key]).max() or 1.0
            assert np.abs(g_[key] - w_[key]).max() <= 1e-14 * scale, key
    prob.close()


@pytest.mark.gpu
def test_print_results_files_and_time_loop_overlap(tmp_path):
    """print_results returns before the files exist, steps go on meanwhile, and every stamp's files hold the state of
    ITS stamp (two host slots in flight)"""
    prob = pecs.SolarCellProblem(pecs.default_input_file(3, 1))
    prob.setup_full_system()
    prob.set_output(str(tmp_path))
    expected = []
    for k in range(4):
        prob.synchronize()
        expected.append(_expected_patches(prob))
        prob.print_results(k)
        prob.step(2)
    prob.finish_output()
    for k in range(4):
        f = read_vtu(os.path.join(tmp_path, f"Semiconductor-{k:03d}.vtu"))
        assert np.allclose(f["point_data"]["Electrons Density"], expected[k][0]["density_1"], rtol=1e-14, atol=0)
        assert np.allclose(f["point_data"]["Holes Current"], expected[k][0]["current_2"], rtol=1e-14, atol=1e-300)
        f = read_vtu(os.path.join(tmp_path, f"Electrolyte-{k:03d}.vtu"))
        assert np.allclose(f["point_data"]["Oxidants Density"], expected[k][1]["density_2"], rtol=1e-14, atol=0)
        f = read_vtu(os.path.join(tmp_path, f"Poisson-{k:03d}.vtu"))
        scale = np.abs(expected[k][2]["field"]).max()
        assert np.abs(f["point_data"]["Field"] - expected[k][2]["field"]).max() <= 1e-14 * scale
        assert np.allclose(f["point_data"]["Potential"], expected[k][2]["potential"], rtol=1e-14, atol=0)
    prob.close()


@pytest.mark.gpu
def test_run_full_system_writes_the_reference_files(tmp_path):
    """run_full_system with the reference's cadence (source/SolarCell.cpp:2036-2092): stamp 0 = initial values, one
    set of files per time stamp, restart files at the end; the last stamp holds the state after all steps"""
    prm = pecs.default_input_file(2, 1, computational__end_time=0.2, computational__time_stamps=2)
    prob = pecs.SolarCellProblem(prm)
    prob.set_output(str(tmp_path))
    prob.run_full_system()
    names = sorted(os.listdir(tmp_path))
    for k in range(3):
        for stem in ("Poisson-", "Semiconductor-", "Electrolyte-"):
            assert f"{stem}{k:03d}.vtu" in names
    assert not any(n.endswith("003.vtu") for n in names)
    for carrier in ("Electrons", "Holes", "Reductants", "Oxidants"):
        assert f"{carrier}.dofs" in names
    # 0.2 / 0.05 = 4 steps (the reference's floating-point loop: time < timeStamps[k]); compare with a stepped problem
    other = pecs.SolarCellProblem(prm)
    other.setup_full_system()
    first = read_vtu(os.path.join(tmp_path, "Semiconductor-000.vtu"))
    assert np.allclose(first["point_data"]["Electrons Density"],
                       oracle_output.carrier_patches(other.get_solution(pecs.ELECTRONS), 1.0)[1], rtol=1e-14, atol=0)
    n_steps = 0
    t = 0.0
    for stamp in (0.1, 0.2):
        while t < stamp:
            t += 0.05
            n_steps += 1
    other.step(n_steps)
    last = read_vtu(os.path.join(tmp_path, "Electrolyte-002.vtu"))
    assert np.allclose(last["point_data"]["Reductants Density"],
                       oracle_output.carrier_patches(other.get_solution(pecs.REDUCTANTS), 1.0)[1], rtol=1e-13, atol=0)
    # restart file = dealii::Vector::block_write layout: "<n>\\n[" + raw doubles + "]"
    raw = open(os.path.join(tmp_path, "Oxidants.dofs"), "rb").read()
    n = other.n_dofs(pecs.OXIDANTS)
    head = f"{n}\n[".encode()
    assert raw.startswith(head) and raw.endswith(b"]")
    assert np.array_equal(np.frombuffer(raw[len(head):-1], np.float64), other.get_solution(pecs.OXIDANTS))
    prob.close()
    other.close()


def test_interface_currents_on_cpu():
    """I-V post-processing: the interface integrals against the numpy restatement from random states, and
    exactly for constant densities: k (rho - rho^e) rho' x the length of the interface (the straight line from
    (radius one, height) to (radius two, 0) of the default wire)"""
    prob = _problem(3, 1)
    rng = np.random.default_rng(11)
    p = prob.params
    k_et, k_ht, rho_n_e, rho_p_e = p[9], p[10], p[16], p[17]  # PECS_P_K_ET, K_HT, RHO_N_E, RHO_P_E
    states = [rng.uniform(0.5, 3.0, 12 * prob.n_cells(w // 2)) for w in range(4)]
    pairs = prob.interface_pairs()
    want = oracle_output.interface_currents(prob.mesh(0)["vertices"], pairs, states, k_et, k_ht, rho_n_e, rho_p_e)
    got = prob.interface_currents(states)
    assert np.allclose(got, want, rtol=1e-13, atol=0)
    const = [np.full(12 * prob.n_cells(w // 2), c) for w, c in enumerate((5.0, 3.0, 7.0, 11.0))]
    length = np.hypot(0.6 - 0.3, 1.0)
    exact = np.array([k_et * (5.0 - rho_n_e) * 11.0, k_ht * (3.0 - rho_p_e) * 7.0]) * length
    assert np.allclose(prob.interface_currents(const), exact, rtol=1e-12, atol=0)
    prob.close()


@pytest.mark.gpu
def test_interface_currents_from_the_device_state():
    """one I-V point integrated ON the device (pecs_interface_currents: a kernel over the interface cells + an ordered
    sum) equals the host integral over the downloaded vectors (same formula, other summation order: 1e-13)"""
    prob = pecs.SolarCellProblem(pecs.default_input_file(3, 1, physical__insulated=False, physical__applied_bias=0.1))
    prob.setup_full_system()
    prob.step(5)
    got = prob.interface_currents()
    want = prob.interface_currents([prob.get_solution(s) for s in range(4)])
    assert np.isfinite(got).all() and np.all(want != 0.0)
    assert np.allclose(got, want, rtol=1e-13, atol=0)
    assert np.array_equal(got, prob.interface_currents())  # deterministic
    prob.close()


GOLDEN_OUTPUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "output_g3_l1_biased.npz")
GOLDEN_OVERRIDES = {"physical__insulated": False, "physical__applied_bias": 0.1}


def test_output_golden_is_what_the_oracle_produces():
    """tests/golden/output_g3_l1_biased.npz (make_golden.py: oracle state after 5 steps -> oracle/output.py) guards
    the restatement of the output path and the I-V integrals against drift"""
    import sys
    sys.path.insert(0, os.path.dirname(GOLDEN_OUTPUT))
    import make_golden
    gold = np.load(GOLDEN_OUTPUT)
    fresh = make_golden.output_case(3, 1, int(gold["n_steps"]), **GOLDEN_OVERRIDES)
    for key in gold.files:
        assert np.allclose(fresh[key], gold[key], rtol=1e-12, atol=1e-300), key


@pytest.mark.gpu
def test_device_output_matches_golden_fixture():
    """the GPU box compares against numbers produced in the build container: patches of all five vectors after 5 steps
    (1e-9: the state tolerance) and the I-V point (1e-5: a small difference of nearly equal densities)"""
    gold = np.load(GOLDEN_OUTPUT)
    prob = pecs.SolarCellProblem(pecs.default_input_file(3, 1, **GOLDEN_OVERRIDES))
    prob.setup_full_system()
    prob.step(int(gold["n_steps"]))
    semi, elec, poisson = prob.output_snapshot()
    for s, (snap, k) in enumerate([(semi, 1), (semi, 2), (elec, 1), (elec, 2)]):
        assert np.abs(snap[f"density_{k}"] - gold[f"density_{s}"]).max() <= 1e-9 * np.abs(gold[f"density_{s}"]).max()
        assert np.abs(snap[f"current_{k}"] - gold[f"current_{s}"]).max() <= 1e-7 * np.abs(gold[f"current_{s}"]).max()
    assert np.abs(poisson["potential"] - gold["potential"]).max() <= 1e-9 * np.abs(gold["potential"]).max()
    assert np.abs(poisson["field"] - gold["field"]).max() <= 1e-9 * np.abs(gold["field"]).max()
    assert np.allclose(prob.interface_currents(), gold["interface_currents"], rtol=1e-5, atol=0)
    prob.close()


@pytest.mark.gpu
def test_restart_leg_continues_the_first_leg(tmp_path):
    """restart status = true (reference source/SolarCell.cpp:1980-1995): read_dofs takes the *.dofs files of the first
    leg as initial values, the time stamps run from `end time` to `end time 2`, and the files continue the numbering
    (time_step_number = number_outputs).  Two legs of 4 steps must end where one run of 8 steps ends: the restart files
    hold the complete state a step reads (the densities); the potential is solved again from them, from a zero initial
    guess instead of the previous potential (increment form), hence round-off-level differences, not bit identity."""
    first = pecs.default_input_file(2, 1, computational__end_time=0.2, computational__time_stamps=2)
    prob = pecs.SolarCellProblem(first)
    prob.set_output(str(tmp_path))
    prob.run_full_system()
    prob.close()
    before = set(os.listdir(tmp_path))
    second = pecs.default_input_file(2, 1, computational__end_time=0.2, computational__end_time_2=0.4,
                                     computational__time_stamps=2, computational__restart_status=True)
    prob = pecs.SolarCellProblem(second)
    prob.set_output(str(tmp_path))
    prob.run_full_system()
    after = set(os.listdir(tmp_path))
    # the second leg writes stamps 2 (its initial values), 3 and 4; stamps 0 and 1 of the first leg survive
    assert {"Poisson-003.vtu", "Poisson-004.vtu", "Semiconductor-004.vtu", "Electrolyte-004.vtu"} <= after - before
    assert {"Poisson-000.vtu", "Poisson-001.vtu"} <= after
    restarted = [prob.get_solution(s) for s in range(5)]
    prob.close()
    whole = pecs.SolarCellProblem(first)
    whole.setup_full_system()
    n_steps, t = 0, 0.0
    for stamp in (0.1, 0.2):          # the first leg's floating-point loop
        while t < stamp:
            t += 0.05
            n_steps += 1
    t = 0.2
    for stamp in (0.3, 0.4):          # the second leg's: time starts at `end time` exactly
        while t < stamp:
            t += 0.05
            n_steps += 1
    whole.step(n_steps)
    for s in range(5):
        got, want = restarted[s], whole.get_solution(s)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 1e-11 * scale, f"vector {s}"
    whole.close()


def test_truncated_restart_file_is_rejected(tmp_path):
    """block_read checks the byte count and the closing bracket (a truncated checkpoint must not be accepted silently)"""
    prm = pecs.default_input_file(2, 1, computational__restart_status=True)
    n = 12 * (4 ** 2 + 4 ** 3)
    good = f"{n}\n[".encode() + np.arange(n, dtype=np.float64).tobytes() + b"]"
    for name in ("Electrons", "Holes", "Reductants", "Oxidants"):
        (tmp_path / f"{name}.dofs").write_bytes(good)
    (tmp_path / "Holes.dofs").write_bytes(good[:-100])
    prob = pecs.SolarCellProblem(prm)
    prob.set_output(str(tmp_path), write_output=False)
    with pytest.raises(pecs.PecsError) as e:
        prob.setup_full_system()   # read_dofs comes before the (GPU-only) factorisations
    assert "truncated" in str(e.value)
    prob.close()
