"""SolarCellProblem -- Python face of SOLARCELL::SolarCellProblem (reference include/SolarCell.hpp:105-757).

Method names follow the reference.  Every per-step method enqueues CUDA work through the C ABI; vectors come back
as numpy arrays in the reference's layout ([Jx | Jy | rho] per carrier, [RT0 | Phi] for Poisson).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import check

KIND_PRODUCTION, KIND_TEST_STEADY, KIND_TEST_TRANSIENT, KIND_TEST_DD_POISSON = 0, 1, 2, 3
ELECTRONS, HOLES, REDUCTANTS, OXIDANTS, POISSON = 0, 1, 2, 3, 4
SEMICONDUCTOR_MESH, ELECTROLYTE_MESH, POISSON_MESH = 0, 1, 2

# include/pecs_b200.h PECS_P_*
PARAM_NAMES = ["delta_t", "penalty", "mu_n", "mu_p", "mu_r", "mu_o", "eps_s", "eps_e", "lambda2", "k_et", "k_ht",
               "v_n", "v_p", "gen_flux", "gen_alpha", "gen_location", "rho_n_e", "rho_p_e", "rho_r_e", "rho_o_e",
               "phi_bi", "phi_app", "phi_sch", "sch_location", "transient", "srh", "n_intrinsic", "tau_n", "tau_p"]

(INFO_LAUNCHES_PER_STEP, INFO_FACTOR_BYTES, INFO_SOLVE_BYTES_PER_STEP, INFO_TREE_LEVELS_MAX, INFO_RHS_BYTES_PER_STEP,
 INFO_HOST_STEP_H2D_BYTES, INFO_HOST_STEP_D2H_BYTES, INFO_SOLVE_WAIT_ERRORS, INFO_SHARED_FACTOR_PAIRS) = range(9)


def device_count():
    return int(_lib.load().pecs_device_count())


def device_warmup(device=0):
    """pay a device's first-context costs now (CUDA context, kernel image, solver handles); optional, changes no result"""
    check(_lib.load().pecs_device_warmup(int(device)))


def default_input_file(global_refinements=4, local_refinements=1, **overrides):
    """Text of the reference's input_file.prm (reference input_file.prm:1-133) with optional overrides given as
    ``section__key=value`` (spaces in names written as single underscores), e.g. physical__applied_bias=0.1."""
    sections = {
        "computational": {"end time": "5e1", "end time 2": "1e5", "global refinements": str(global_refinements),
                          "local refinements": str(local_refinements), "restart status": "false",
                          "time stamps": "100", "time step size": "0.05"},
        "electrons": {"mobility": "1350.0", "recombination time": "5e-5", "recombination velocity": "3e5",
                      "transfer rate": "1e-19"},
        "holes": {"mobility": "480.0", "recombination time": "5e-5", "recombination velocity": "2.9e5",
                  "transfer rate": "1e-14"},
        "mesh": {"boundary layer": "0.1", "mesh height": "1", "mesh length": "1.0", "radius one": "0.3",
                 "radius two": "0.6"},
        "oxidants": {"mobility": "1.0"},
        "physical": {"absorption coefficient": "1.74974e5", "applied bias": "0.0", "built in bias": "0.41",
                     "characteristic density": "1.0e16", "characteristic length": "1.0e-4",
                     "characteristic time": "1.0e-12", "electrolyte permittivity": "1000",
                     "illumination status": "true", "insulated": "true", "intrinsic density": "2.564e9",
                     "photon flux": "1.2e17", "schottky bias": "0.0", "schottky status": "true",
                     "semiconductor permittivity": "11.9",
                     # not in the reference's file (its SRH_Recombination returns 0.0): switches the formula on
                     "srh recombination": "false"},
        "reductants": {"mobility": "1.0"},
    }
    for key, value in overrides.items():
        sec, name = key.split("__", 1)
        name = name.replace("_", " ")
        if sec not in sections or name not in sections[sec]:
            raise KeyError(f"unknown parameter {sec}/{name}")
        sections[sec][name] = str(value).lower() if isinstance(value, bool) else str(value)
    lines = []
    for sec, entries in sections.items():
        lines.append(f"subsection {sec}")
        lines += [f"  set {k} = {v}" for k, v in entries.items()]
        lines.append("end\n")
    return "\n".join(lines)


def _dp(a):
    return a.ctypes.data_as(_lib.c_double_p)


def _ip(a):
    return a.ctypes.data_as(_lib.c_int32_p)


class SolarCellProblem:
    def __init__(self, prm_text=None, test_defaults=False, device=0):
        """main() of the reference: read the parameter file, construct SolarCellProblem<2>(degree=1, prm).
        prm_text may be the text of a .prm file or a path to one."""
        self._lib = _lib.load()
        if prm_text is not None and "\n" not in prm_text and os.path.exists(prm_text):
            with open(prm_text) as f:
                prm_text = f.read()
        self._h = C.c_void_p()
        check(self._lib.pecs_solarcell_create((prm_text or "").encode(), int(bool(test_defaults)), int(device),
                                              C.byref(self._h)))

    def close(self):
        for p in getattr(self, "_pinned", []):
            self._lib.pecs_host_free(p)
        self._pinned = []
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pecs_solarcell_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup ----
    def setup_full_system_host(self):
        check(self._lib.pecs_solarcell_setup_full_system_host(self._h))

    def setup_full_system(self):
        check(self._lib.pecs_solarcell_setup_full_system(self._h))

    def setup_test_host(self, kind, n_refine):
        check(self._lib.pecs_solarcell_setup_test_host(self._h, kind, n_refine))

    def setup_test(self, kind, n_refine):
        check(self._lib.pecs_solarcell_setup_test(self._h, kind, n_refine))

    def run_full_system(self):
        check(self._lib.pecs_solarcell_run_full_system(self._h))

    # ---- output path (reference SolarCell.cpp:1826-1858) ----
    def set_output(self, directory=None, write_output=True):
        check(self._lib.pecs_solarcell_set_output(self._h, (directory or "").encode(), int(bool(write_output))))

    def print_results(self, time_step_number):
        """Poisson-NNN.vtu, Semiconductor-NNN.vtu, Electrolyte-NNN.vtu of the current state; asynchronous."""
        check(self._lib.pecs_solarcell_print_results(self._h, int(time_step_number)))

    def finish_output(self):
        check(self._lib.pecs_solarcell_finish_output(self._h))

    def write_patches(self, which, patches, time_step_number, directory="."):
        """host half of the output path alone: the .vtu file of mesh `which` from given patch values"""
        patches = np.ascontiguousarray(patches, dtype=np.float64)
        check(self._lib.pecs_solarcell_write_patches(self._h, which, _dp(patches), int(time_step_number),
                                                     directory.encode()))

    @property
    def output_scales(self):
        """PostProcessor scales {potential, field, density, current}"""
        s = np.zeros(4)
        check(self._lib.pecs_solarcell_output_scales(self._h, _dp(s)))
        return s

    def output_snapshot(self):
        """The rescaled patch values straight from pecs_output_snapshot: list of three dicts of arrays
        (pairs: current_1 [4n,3], density_1 [4n], current_2, density_2; Poisson: field [4n,3], potential [4n])."""
        n = [int(self._lib.pecs_output_doubles(self.ctx, w)) for w in range(3)]
        bufs = [np.zeros(max(k, 0)) for k in n]
        ptrs = (_lib.c_double_p * 3)(*[_dp(b) if b.size else None for b in bufs])
        check(self._lib.pecs_output_snapshot(self.ctx, _dp(self.output_scales), ptrs))
        check(self._lib.pecs_output_wait(self.ctx))
        out = []
        for w, b in enumerate(bufs):
            if b.size == 0:
                out.append(None)
            elif w < 2:
                c = b.size // 32
                out.append({"current_1": b[:12 * c].reshape(-1, 3), "density_1": b[12 * c:16 * c],
                            "current_2": b[16 * c:28 * c].reshape(-1, 3), "density_2": b[28 * c:]})
            else:
                c = b.size // 16
                out.append({"field": b[:12 * c].reshape(-1, 3), "potential": b[12 * c:]})
        return out

    def interface_currents(self, states=None):
        """(electron-transfer, hole-transfer) current through the semiconductor-electrolyte interface, scaled units;
        states = four carrier vectors, or None for the current device state (one point of an I-V curve)"""
        out = np.zeros(2)
        if states is None:
            check(self._lib.pecs_solarcell_interface_currents(self._h, None, _dp(out)))
        else:
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in states]
            ptrs = (_lib.c_double_p * 4)(*[_dp(a) for a in arrs])
            check(self._lib.pecs_solarcell_interface_currents(self._h, ptrs, _dp(out)))
        return out

    def selftest_carrier_rhs(self, which, u1, u2, X, o1=None, o2=None):
        """CPU evaluation of the production kernels' arithmetic (csrc/rhs_math.hpp): carrier right-hand sides of
        subdomain `which`; without o1/o2 (the other subdomain's carriers) only the cell terms -> (rhs1, rhs2)"""
        u1, u2, X = (np.ascontiguousarray(a, dtype=np.float64) for a in (u1, u2, X))
        if o1 is not None:
            o1, o2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (o1, o2))
        r1, r2 = np.zeros_like(u1), np.zeros_like(u2)
        check(_lib.load_selftest().pecs_solarcell_selftest_carrier_rhs(
            self._h, which, _dp(u1), _dp(u2), _dp(o1) if o1 is not None else None, _dp(o2) if o2 is not None else None,
            _dp(X), _dp(r1), _dp(r2)))
        return r1, r2

    def selftest_poisson_rows(self, densities):
        """CPU evaluation of the Poisson charge rows the device computes (csrc/rhs_math.hpp) -> [n_poisson_cells]"""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in densities]
        ptrs = (_lib.c_double_p * 4)(*[_dp(a) for a in arrs])
        out = np.zeros(self.n_cells(POISSON_MESH))
        check(_lib.load_selftest().pecs_solarcell_selftest_poisson_rows(self._h, ptrs, _dp(out)))
        return out

    def selftest_field_patches(self, X, scale):
        """CPU evaluation of the RT0 field at the patch vertices as the output kernel computes it -> [4n, 2]"""
        X = np.ascontiguousarray(X, dtype=np.float64)
        out = np.zeros((4 * self.n_cells(POISSON_MESH), 2))
        check(_lib.load_selftest().pecs_solarcell_selftest_field_patches(self._h, _dp(X), float(scale), _dp(out)))
        return out

    def run_test(self, kind, n_refine):
        """test_steady_state / test_transient / test_DD_Poisson at one level -> dict of L2 errors."""
        e = np.zeros(4)
        check(self._lib.pecs_solarcell_run_test(self._h, kind, n_refine, _dp(e)))
        return {"u": e[0], "J": e[1], "Phi": e[2], "D": e[3]}

    def project_initial_conditions(self):
        check(self._lib.pecs_solarcell_project_initial_conditions(self._h))

    def project_test_initial_condition(self):
        check(self._lib.pecs_solarcell_project_test_initial_condition(self._h))

    # ---- tables ----
    @property
    def ctx(self):
        return C.c_void_p(self._lib.pecs_solarcell_ctx(self._h))

    @property
    def params(self):
        p = np.zeros(32)
        check(self._lib.pecs_solarcell_get_params(self._h, _dp(p)))
        return p

    @property
    def delta_t(self):
        return float(self._lib.pecs_solarcell_delta_t(self._h))

    def n_cells(self, which):
        return int(self._lib.pecs_solarcell_n_cells(self._h, which))

    def mesh(self, which):
        n = self.n_cells(which)
        m = {"n_cells": n, "vertices": np.zeros((n, 4, 2)), "material_id": np.zeros(n, np.int32),
             "level": np.zeros(n, np.int32), "face_kind": np.zeros((n, 4), np.int32),
             "neighbor": np.zeros((n, 4), np.int32), "neighbor2": np.zeros((n, 4), np.int32),
             "boundary_id": np.zeros((n, 4), np.int32), "nb_parent_diameter": np.zeros((n, 4))}
        check(self._lib.pecs_solarcell_get_mesh(self._h, which, _dp(m["vertices"]), _ip(m["material_id"]),
                                                _ip(m["level"]), _ip(m["face_kind"]), _ip(m["neighbor"]),
                                                _ip(m["neighbor2"]), _ip(m["boundary_id"]),
                                                _dp(m["nb_parent_diameter"])))
        return m

    @property
    def n_rt(self):
        return int(self._lib.pecs_solarcell_n_rt(self._h))

    def poisson_face_dofs(self):
        a = np.zeros((self.n_cells(POISSON_MESH), 4), np.int32)
        check(self._lib.pecs_solarcell_get_poisson_face_dofs(self._h, _ip(a)))
        return a

    def constraints(self):
        n = int(self._lib.pecs_solarcell_n_constraints(self._h))
        dof, master, w = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n)
        check(self._lib.pecs_solarcell_get_constraints(self._h, _ip(dof), _ip(master), _dp(w)))
        return dof, master, w

    def cell_map(self, which):
        a = np.zeros(self.n_cells(which), np.int32)
        check(self._lib.pecs_solarcell_get_cell_map(self._h, which, _ip(a)))
        return a

    def interface_pairs(self):
        n = int(self._lib.pecs_solarcell_n_interface_pairs(self._h))
        arrs = [np.zeros(n, np.int32) for _ in range(4)]
        check(self._lib.pecs_solarcell_get_interface_pairs(self._h, *[_ip(a) for a in arrs]))
        return arrs

    def matrix(self, which):
        """scipy.sparse.csr_matrix of a constant matrix: 0..3 species, 4 Poisson, 5/6 mass matrices."""
        import scipy.sparse as sp
        nnz = int(self._lib.pecs_solarcell_matrix_nnz(self._h, which))
        if which == POISSON:
            n = self.n_rt + self.n_cells(POISSON_MESH)
        else:
            n = 12 * self.n_cells(SEMICONDUCTOR_MESH if which in (0, 1, 5) else ELECTROLYTE_MESH)
        rp, col, val = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        check(self._lib.pecs_solarcell_get_matrix(self._h, which, _ip(rp), _ip(col), _dp(val)))
        return sp.csr_matrix((val, col, rp), shape=(n, n))

    def host_solution(self, which):
        n = 12 * self.n_cells(SEMICONDUCTOR_MESH if which in (0, 1) else ELECTROLYTE_MESH)
        a = np.zeros(n)
        check(self._lib.pecs_solarcell_get_host_solution(self._h, which, _dp(a)))
        return a

    # ---- device state ----
    def n_dofs(self, which):
        return int(self._lib.pecs_n_dofs(self.ctx, which))

    def get_solution(self, which):
        a = np.zeros(self.n_dofs(which))
        check(self._lib.pecs_get_state(self.ctx, which, _dp(a)))
        return a

    def set_solution(self, which, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        assert a.size == self.n_dofs(which)
        check(self._lib.pecs_set_state(self.ctx, which, _dp(a)))

    def get_rhs(self, which):
        a = np.zeros(self.n_dofs(which))
        check(self._lib.pecs_get_rhs(self.ctx, which, _dp(a)))
        return a

    def set_rhs(self, which, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        assert a.size == self.n_dofs(which)
        check(self._lib.pecs_set_rhs(self.ctx, which, _dp(a)))

    # ---- the hot path (reference source/SolarCell.cpp:2057-2075) ----
    def set_time(self, t):
        check(self._lib.pecs_set_time(self.ctx, float(t)))

    def assemble_semiconductor_rhs(self):
        check(self._lib.pecs_assemble_semiconductor_rhs(self.ctx))

    def assemble_electrolyte_rhs(self):
        check(self._lib.pecs_assemble_electrolyte_rhs(self.ctx))

    def solve_full_system(self):
        check(self._lib.pecs_solve_full_system(self.ctx))

    def solve_species(self, which):
        check(self._lib.pecs_solve_species(self.ctx, which))

    def assemble_Poisson_rhs(self):
        check(self._lib.pecs_assemble_poisson_rhs(self.ctx))

    def solve_Poisson(self):
        check(self._lib.pecs_solve_poisson(self.ctx))

    def step(self, n_steps=1):
        check(self._lib.pecs_step(self.ctx, int(n_steps)))

    # ---- one step cut at its exchange point (pecs_b200/shard.py drives these over several GPUs) ----
    def set_owned_species(self, mask):
        """before setup_*: bit k set = carrier k is factorised and solved by this process"""
        check(self._lib.pecs_solarcell_set_owned_species(self._h, int(mask)))

    def step_local(self):
        check(self._lib.pecs_step_local(self.ctx))

    def step_finish(self):
        check(self._lib.pecs_step_finish(self.ctx))

    def p2p_export(self):
        """CUDA IPC handles of this context's state vectors and flag block (bytes), for the other ranks"""
        buf = C.create_string_buffer(1024)
        n = self._lib.pecs_p2p_export(self.ctx, buf, 1024)
        if n < 0:
            raise RuntimeError("pecs_p2p_export failed")
        return buf.raw[:n]

    def p2p_connect(self, rank, world, blobs):
        """blobs: the p2p_export() bytes of all ranks in rank order; fuses the density exchange into the solves"""
        joined = b"".join(blobs)
        check(self._lib.pecs_p2p_connect(self.ctx, int(rank), int(world), C.create_string_buffer(joined, len(joined))))

    def density_block(self, which):
        """(device pointer, number of doubles) of the density block of carrier `which`"""
        n = C.c_int64(0)
        ptr = self._lib.pecs_density_block(self.ctx, int(which), C.byref(n))
        if not ptr:
            raise ValueError("no such carrier in this context")
        return int(ptr), int(n.value)

    @property
    def stream(self):
        """the context's main cudaStream_t as an integer handle"""
        return int(self._lib.pecs_stream(self.ctx) or 0)

    def pinned_states(self):
        """five page-locked numpy arrays (electrons, holes, reductants, oxidants, Poisson) for step_host()"""
        out = []
        for w in range(5):
            n = self.n_dofs(w)
            p = self._lib.pecs_host_alloc(8 * max(n, 1))
            if not p:
                raise MemoryError("pecs_host_alloc failed")
            arr = np.ctypeslib.as_array(C.cast(p, _lib.c_double_p), shape=(max(n, 1),))[:n]
            self._pinned = getattr(self, "_pinned", []) + [p]
            out.append(arr)
        return out

    def step_host(self, n_steps, states):
        """n steps with host-resident state: H2D of the five solutions, the steps, D2H of the five solutions."""
        ptrs = (_lib.c_double_p * 5)(*[_dp(a) if a.size else None for a in states])
        check(self._lib.pecs_step_host(self.ctx, int(n_steps), ptrs))

    def synchronize(self):
        check(self._lib.pecs_synchronize(self.ctx))

    def step_timed(self, n_steps, sectioned=False):
        ms = np.zeros(6)
        check(self._lib.pecs_step_timed(self.ctx, int(n_steps), int(sectioned), _dp(ms)))
        return ms

    def time_kernel(self, which, repeats):
        ms, launches = C.c_double(0), C.c_int32(0)
        check(self._lib.pecs_time_kernel(self.ctx, which, repeats, C.byref(ms), C.byref(launches)))
        return ms.value, launches.value

    def info(self, what):
        return int(self._lib.pecs_get_info(self.ctx, what))

    # ---- verification of the setup tables (CPU tests) ----
    def plan_stats(self, which, leaf_nodes=0):
        st = np.zeros(8, np.int64)
        check(self._lib.pecs_solarcell_plan_stats(self._h, which, leaf_nodes, st.ctypes.data_as(C.POINTER(C.c_int64))))
        return dict(zip(["fronts", "levels", "max_np", "max_nb", "fwd_entries", "bwd_entries", "upd_entries"], st))

    def plan_levels(self, which, leaf_nodes=0):
        out = np.zeros((64, 6), np.int64)
        n = self._lib.pecs_solarcell_plan_levels(self._h, which, leaf_nodes, out.ctypes.data_as(C.POINTER(C.c_int64)), 64)
        return out[:n]

    def plan_fronts(self, which, leaf_nodes=0):
        """per front: depth, np, nb, log2P forward, log2P backward, small forward, small backward, parent"""
        cap = 1 << 20
        out = np.zeros((cap, 8), np.int32)
        n = self._lib.pecs_solarcell_plan_fronts(self._h, which, leaf_nodes, _ip(out), cap)
        if n < 0:
            raise RuntimeError("plan_fronts failed")
        return out[:n]

    def selftest_direct_solve(self, which, b, leaf_nodes=0):
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros_like(b)
        check(_lib.load_selftest().pecs_solarcell_selftest_direct_solve(self._h, which, leaf_nodes, _dp(b), _dp(x)))
        return x

    def selftest_prepared_hashes(self, which):
        """hashes of everything the host preparation of system `which` produces (tests: independent of the thread count)"""
        import ctypes
        h = (ctypes.c_uint64 * 8)()
        check(_lib.load_selftest().pecs_solarcell_selftest_prepared_hashes(self._h, which, h))
        return tuple(int(v) for v in h)

    def selftest_ell_matvec(self, which, table, x):
        """(y from the host ELL table with the device kernel's arithmetic, y from the CSR matrix, (rows, slots, block))"""
        import ctypes
        n = self.n_cells(which // 2)
        rows = (4 * n, 4 * n, 8 * n, 8 * n)[table]
        x = np.ascontiguousarray(x, np.float64)
        y_ell, y_csr = np.zeros(rows), np.zeros(rows)
        shape = (ctypes.c_int32 * 3)()
        check(_lib.load_selftest().pecs_solarcell_selftest_ell_matvec(self._h, which, table, _dp(x), _dp(y_ell), _dp(y_csr), shape))
        return y_ell, y_csr, tuple(shape)

    # ---- post-processing ----
    def ldg_errors(self, which, time):
        e = np.zeros(2)
        check(self._lib.pecs_solarcell_ldg_errors(self._h, which, float(time), _dp(e)))
        return e

    def mixed_errors(self):
        e = np.zeros(2)
        check(self._lib.pecs_solarcell_mixed_errors(self._h, _dp(e)))
        return e
