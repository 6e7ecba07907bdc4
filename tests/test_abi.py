"""The C-ABI library loads on a CPU-only box, exports every symbol the headers declare, reads .prm files like the
reference's ParameterReader, and refuses to run the per-step path without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import pecs_b200 as pecs
from pecs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(headers=("pecs_b200.h", "pecs_b200_host.h")):
    names = set()
    for header in headers:
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(pecs_[a-z0-9_]+)\s*\(", text))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(_lib.SIGNATURES), "python binding and headers disagree"


def test_cpu_checkers_live_in_the_test_library_only():
    """the CPU evaluation of the device formulas and the host reference of the solve sweeps are test
    infrastructure -- exported by libpecs_b200_selftest.so (include/pecs_b200_selftest.h), absent from the product"""
    product = open(_lib.LIB_PATH, "rb").read()
    declared = _declared_symbols(("pecs_b200_selftest.h",))
    assert declared == set(_lib.SELFTEST_SIGNATURES) and len(declared) == 6
    test_lib = _lib.load_selftest()
    for name in declared:
        assert hasattr(test_lib, name)
        assert name.encode() not in product, f"{name} is still in the product library"
    assert b"solve_host" not in product and b"solve_system_host" not in product


def test_oracle_is_not_linked_into_the_product():
    """the product library must not depend on the oracle (checked on the dynamic symbol table and on the sources)"""
    data = open(_lib.LIB_PATH, "rb").read()
    assert b"oracle_create" not in data and b"libpecs_oracle" not in data
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pecs_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f


def test_parameter_scaling_matches_reference_defaults():
    """reference include/Parameters.hpp:181-242 applied to input_file.prm (values quoted in SURVEY App. D)"""
    prob = pecs.SolarCellProblem(pecs.default_input_file(2, 1))
    prob.setup_full_system_host()
    p = dict(zip(pecs.PARAM_NAMES, prob.params))
    assert p["mu_n"] == pytest.approx(3.48975e-3, rel=1e-12)
    assert p["mu_p"] == pytest.approx(1.2408e-3, rel=1e-12)
    assert p["mu_r"] == pytest.approx(2.585e-6, rel=1e-12)
    assert p["lambda2"] == pytest.approx(1.41218e-4, rel=1e-5)
    assert p["k_et"] == pytest.approx(1e-11, rel=1e-12) and p["k_ht"] == pytest.approx(1e-6, rel=1e-12)
    assert p["v_n"] == pytest.approx(3.0e-3, rel=1e-12) and p["v_p"] == pytest.approx(2.9e-3, rel=1e-12)
    assert p["gen_flux"] == pytest.approx(1.2e-11, rel=1e-12) and p["gen_alpha"] == pytest.approx(17.4974, rel=1e-12)
    assert p["phi_bi"] == pytest.approx(0.41 / 0.02585, rel=1e-12)
    assert p["delta_t"] == 0.05


def test_prm_file_round_trip(tmp_path):
    f = tmp_path / "input_file.prm"
    f.write_text(pecs.default_input_file(3, 0, physical__applied_bias=0.2, physical__insulated=False))
    prob = pecs.SolarCellProblem(str(f))
    prob.setup_full_system_host()
    assert prob.n_cells(0) == 64  # l = 0: no boundary layer, 4^3 cells
    assert dict(zip(pecs.PARAM_NAMES, prob.params))["phi_app"] == pytest.approx(0.2 / 0.02585)
    with pytest.raises(pecs.PecsError):
        pecs.SolarCellProblem("subsection mesh\n  set no such entry = 1\nend\n")


def test_no_cpu_fallback():
    if pecs.device_count() > 0:
        pytest.skip("a CUDA device is present")
    prob = pecs.SolarCellProblem(pecs.default_input_file(2, 1))
    with pytest.raises(pecs.PecsError) as e:
        prob.setup_full_system()
    assert e.value.status == "PECS_ERR_NO_DEVICE"
    with pytest.raises(pecs.PecsError):
        prob.run_full_system()
    with pytest.raises(pecs.PecsError):
        prob.run_test(pecs.KIND_TEST_TRANSIENT, 2)


def test_device_warmup_without_a_device_is_a_clean_error():
    """the optional warm-up entry reports the missing device like everything else (setup_full_system calls it on a
    second thread and ignores its status: set_solvers reports the failure)"""
    if pecs.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pecs.PecsError) as e:
        pecs.device_warmup(0)
    assert e.value.status == "PECS_ERR_NO_DEVICE"


def test_test_initial_condition_projection():
    """host projection (collocation form) == the oracle's mass-matrix projection"""
    from helpers import make_oracle
    prob = pecs.SolarCellProblem(None, test_defaults=True)
    prob.setup_test_host(pecs.KIND_TEST_TRANSIENT, 3)
    prob.project_test_initial_condition()
    o = make_oracle(prob, False, factor=False)
    o.project_test_initial_condition()
    assert np.abs(prob.host_solution(0) - o.solution(0)).max() <= 1e-13
