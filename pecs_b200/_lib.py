"""ctypes loader for libpecs_b200.so (the C ABI declared in include/pecs_b200.h and include/pecs_b200_host.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``pecs_b200/build.py`` into ``pecs_b200/lib``.
There is no Python or CPU fallback: if the shared library is missing this module raises on import of the
symbols, and without a CUDA device ``pecs_ctx_create`` fails with PECS_ERR_NO_DEVICE.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PECS_B200_LIB", os.path.join(_HERE, "lib", "libpecs_b200.so"))

c_int32_p = C.POINTER(C.c_int32)
c_double_p = C.POINTER(C.c_double)


class PecsError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"pecs status {status}: {message}")
        self.status = status


STATUS_NAMES = {0: "PECS_OK", 1: "PECS_ERR_INVALID", 2: "PECS_ERR_NO_DEVICE", 3: "PECS_ERR_CUDA",
                4: "PECS_ERR_SINGULAR", 5: "PECS_ERR_INTERNAL"}

# every exported symbol with (restype, argtypes); tests/test_abi.py checks this list against the headers
VOIDP = C.c_void_p
SIGNATURES = {
    # ---- include/pecs_b200.h ----
    "pecs_last_error": (C.c_char_p, []),
    "pecs_device_count": (C.c_int32, []),
    "pecs_device_warmup": (C.c_int, [C.c_int32]),
    "pecs_ctx_create": (C.c_int, [VOIDP, C.POINTER(VOIDP)]),
    "pecs_ctx_destroy": (None, [VOIDP]),
    "pecs_set_state": (C.c_int, [VOIDP, C.c_int32, c_double_p]),
    "pecs_get_state": (C.c_int, [VOIDP, C.c_int32, c_double_p]),
    "pecs_get_rhs": (C.c_int, [VOIDP, C.c_int32, c_double_p]),
    "pecs_set_rhs": (C.c_int, [VOIDP, C.c_int32, c_double_p]),
    "pecs_n_dofs": (C.c_int32, [VOIDP, C.c_int32]),
    "pecs_set_time": (C.c_int, [VOIDP, C.c_double]),
    "pecs_assemble_semiconductor_rhs": (C.c_int, [VOIDP]),
    "pecs_assemble_electrolyte_rhs": (C.c_int, [VOIDP]),
    "pecs_solve_full_system": (C.c_int, [VOIDP]),
    "pecs_solve_species": (C.c_int, [VOIDP, C.c_int32]),
    "pecs_assemble_poisson_rhs": (C.c_int, [VOIDP]),
    "pecs_solve_poisson": (C.c_int, [VOIDP]),
    "pecs_step": (C.c_int, [VOIDP, C.c_int32]),
    "pecs_step_local": (C.c_int, [VOIDP]),
    "pecs_step_finish": (C.c_int, [VOIDP]),
    "pecs_p2p_export": (C.c_int64, [VOIDP, VOIDP, C.c_int64]),
    "pecs_p2p_connect": (C.c_int, [VOIDP, C.c_int32, C.c_int32, VOIDP]),
    "pecs_density_block": (VOIDP, [VOIDP, C.c_int32, C.POINTER(C.c_int64)]),
    "pecs_stream": (VOIDP, [VOIDP]),
    "pecs_synchronize": (C.c_int, [VOIDP]),
    "pecs_step_host": (C.c_int, [VOIDP, C.c_int32, C.POINTER(c_double_p)]),
    "pecs_host_alloc": (VOIDP, [C.c_uint64]),
    "pecs_host_free": (None, [VOIDP]),
    "pecs_output_doubles": (C.c_int64, [VOIDP, C.c_int32]),
    "pecs_output_snapshot": (C.c_int, [VOIDP, c_double_p, C.POINTER(c_double_p)]),
    "pecs_output_wait": (C.c_int, [VOIDP]),
    "pecs_interface_currents": (C.c_int, [VOIDP, c_double_p]),
    "pecs_step_timed": (C.c_int, [VOIDP, C.c_int32, C.c_int32, c_double_p]),
    "pecs_time_kernel": (C.c_int, [VOIDP, C.c_int32, C.c_int32, c_double_p, c_int32_p]),
    "pecs_get_info": (C.c_int64, [VOIDP, C.c_int32]),
    # ---- include/pecs_b200_host.h ----
    "pecs_solarcell_create": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(VOIDP)]),
    "pecs_solarcell_destroy": (None, [VOIDP]),
    "pecs_solarcell_set_owned_species": (C.c_int, [VOIDP, C.c_int32]),
    "pecs_solarcell_setup_full_system_host": (C.c_int, [VOIDP]),
    "pecs_solarcell_setup_full_system": (C.c_int, [VOIDP]),
    "pecs_solarcell_setup_test_host": (C.c_int, [VOIDP, C.c_int32, C.c_int32]),
    "pecs_solarcell_setup_test": (C.c_int, [VOIDP, C.c_int32, C.c_int32]),
    "pecs_solarcell_run_full_system": (C.c_int, [VOIDP]),
    "pecs_solarcell_set_output": (C.c_int, [VOIDP, C.c_char_p, C.c_int32]),
    "pecs_solarcell_print_results": (C.c_int, [VOIDP, C.c_int32]),
    "pecs_solarcell_finish_output": (C.c_int, [VOIDP]),
    "pecs_solarcell_write_patches": (C.c_int, [VOIDP, C.c_int32, c_double_p, C.c_int32, C.c_char_p]),
    "pecs_solarcell_interface_currents": (C.c_int, [VOIDP, C.POINTER(c_double_p), c_double_p]),
    "pecs_solarcell_output_scales": (C.c_int, [VOIDP, c_double_p]),
    "pecs_solarcell_run_test": (C.c_int, [VOIDP, C.c_int32, C.c_int32, c_double_p]),
    "pecs_solarcell_ctx": (VOIDP, [VOIDP]),
    "pecs_solarcell_get_params": (C.c_int, [VOIDP, c_double_p]),
    "pecs_solarcell_delta_t": (C.c_double, [VOIDP]),
    "pecs_solarcell_n_cells": (C.c_int32, [VOIDP, C.c_int32]),
    "pecs_solarcell_get_mesh": (C.c_int, [VOIDP, C.c_int32, c_double_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                          c_int32_p, c_int32_p, c_double_p]),
    "pecs_solarcell_n_rt": (C.c_int32, [VOIDP]),
    "pecs_solarcell_get_poisson_face_dofs": (C.c_int, [VOIDP, c_int32_p]),
    "pecs_solarcell_n_constraints": (C.c_int32, [VOIDP]),
    "pecs_solarcell_get_constraints": (C.c_int, [VOIDP, c_int32_p, c_int32_p, c_double_p]),
    "pecs_solarcell_get_cell_map": (C.c_int, [VOIDP, C.c_int32, c_int32_p]),
    "pecs_solarcell_n_interface_pairs": (C.c_int32, [VOIDP]),
    "pecs_solarcell_get_interface_pairs": (C.c_int, [VOIDP, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    "pecs_solarcell_matrix_nnz": (C.c_int64, [VOIDP, C.c_int32]),
    "pecs_solarcell_get_matrix": (C.c_int, [VOIDP, C.c_int32, c_int32_p, c_int32_p, c_double_p]),
    "pecs_solarcell_project_initial_conditions": (C.c_int, [VOIDP]),
    "pecs_solarcell_project_test_initial_condition": (C.c_int, [VOIDP]),
    "pecs_solarcell_get_host_solution": (C.c_int, [VOIDP, C.c_int32, c_double_p]),
    "pecs_solarcell_ldg_errors": (C.c_int, [VOIDP, C.c_int32, C.c_double, c_double_p]),
    "pecs_solarcell_mixed_errors": (C.c_int, [VOIDP, c_double_p]),
    "pecs_solarcell_plan_stats": (C.c_int, [VOIDP, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "pecs_solarcell_plan_levels": (C.c_int32, [VOIDP, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_int32]),
    "pecs_solarcell_plan_fronts": (C.c_int64, [VOIDP, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int64]),
}

# ---- include/pecs_b200_selftest.h: the TEST library (CPU checkers; not part of the product, loaded on demand) ----
SELFTEST_SIGNATURES = {
    "pecs_solarcell_selftest_carrier_rhs": (C.c_int, [VOIDP, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p,
                                                     c_double_p, c_double_p, c_double_p]),
    "pecs_solarcell_selftest_poisson_rows": (C.c_int, [VOIDP, C.POINTER(c_double_p), c_double_p]),
    "pecs_solarcell_selftest_field_patches": (C.c_int, [VOIDP, c_double_p, C.c_double, c_double_p]),
    "pecs_solarcell_selftest_direct_solve": (C.c_int, [VOIDP, C.c_int32, C.c_int32, c_double_p, c_double_p]),
    "pecs_solarcell_selftest_prepared_hashes": (C.c_int, [VOIDP, C.c_int32, C.POINTER(C.c_uint64)]),
    "pecs_solarcell_selftest_ell_matvec": (C.c_int, [VOIDP, C.c_int32, C.c_int32, c_double_p, c_double_p, c_double_p,
                                                    C.POINTER(C.c_int32)]),
}
SELFTEST_LIB_PATH = os.path.join(_HERE, "lib", "libpecs_b200_selftest.so")

_lib = None
_selftest = None


def load_selftest():
    """libpecs_b200_selftest.so (tests only): CPU evaluation of the device formulas and of the setup tables"""
    global _selftest
    if _selftest is None:
        load()
        lib = C.CDLL(SELFTEST_LIB_PATH)
        for name, (restype, argtypes) in SELFTEST_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _selftest = lib
    return _selftest


def load():
    """Load libpecs_b200.so once and attach the signatures. Raises OSError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(pecs_b200 has no Python/CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)  # the test library resolves the host classes against it
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().pecs_last_error()
        raise PecsError(STATUS_NAMES.get(status, status), msg.decode() if msg else "")
