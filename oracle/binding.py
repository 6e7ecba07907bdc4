"""ctypes binding of oracle/_build/libpecs_oracle.so (see oracle/capi.cpp). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libpecs_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ -O3 -fopenmp)."""
    srcs = [os.path.join(_HERE, f) for f in ("pecs_oracle.cpp", "capi.cpp", "pecs_oracle.hpp", "fe_values.hpp",
                                              "sparse_lu.hpp", "Makefile")]
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_last_error.restype = C.c_char_p
        _lib.oracle_create.restype = C.c_void_p
        _lib.oracle_matrix_nnz.restype = C.c_long
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Oracle:
    """oracle::SolarCellProblem. Meshes are DATA (the tables a deal.II Triangulation would provide); everything
    else -- dof numbering, mappings, matrices, right-hand sides, LU solves, errors -- is restated independently."""

    def __init__(self, params, full_system=True):
        self.lib = _load()
        p = np.ascontiguousarray(params, dtype=np.float64)
        self.h = C.c_void_p(self.lib.oracle_create(_d(p), int(p.size), int(bool(full_system))))
        self.full_system = bool(full_system)

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status):
        if status != 0:
            raise RuntimeError("oracle: " + self.lib.oracle_last_error().decode())

    def set_mesh(self, which, m):
        arrs = (np.ascontiguousarray(m["vertices"], np.float64), np.ascontiguousarray(m["material_id"], np.int32),
                np.ascontiguousarray(m["face_kind"], np.int32), np.ascontiguousarray(m["neighbor"], np.int32),
                np.ascontiguousarray(m["neighbor2"], np.int32), np.ascontiguousarray(m["boundary_id"], np.int32),
                np.ascontiguousarray(m["nb_parent_diameter"], np.float64))
        self._check(self.lib.oracle_set_mesh(self.h, which, int(m["n_cells"]), _d(arrs[0]), _i(arrs[1]), _i(arrs[2]),
                                             _i(arrs[3]), _i(arrs[4]), _i(arrs[5]), _d(arrs[6])))

    def setup(self, transient_or_steady=1.0, factor=True):
        self._check(self.lib.oracle_setup(self.h, C.c_double(transient_or_steady), int(bool(factor))))

    def n_dofs(self, which):
        return int(self.lib.oracle_n_dofs(self.h, which))

    @property
    def n_rt(self):
        return int(self.lib.oracle_n_rt(self.h))

    def get_vector(self, which, kind):
        a = np.zeros(self.n_dofs(which))
        self._check(self.lib.oracle_get_vector(self.h, which, kind, _d(a)))
        return a

    def set_vector(self, which, kind, values):
        a = np.ascontiguousarray(values, np.float64)
        assert a.size == self.n_dofs(which)
        self._check(self.lib.oracle_set_vector(self.h, which, kind, _d(a)))

    def solution(self, which):
        return self.get_vector(which, 0)

    def rhs(self, which):
        return self.get_vector(which, 1)

    def matrix(self, which):
        import scipy.sparse as sp
        nnz = int(self.lib.oracle_matrix_nnz(self.h, which))
        n = self.n_dofs({5: 0, 6: 2}.get(which, which))
        rp, col, val = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        self._check(self.lib.oracle_get_matrix(self.h, which, _i(rp), _i(col), _d(val)))
        return sp.csr_matrix((val, col, rp), shape=(n, n))

    def poisson_face_dofs(self, n_cells):
        a = np.zeros((n_cells, 4), np.int32)
        self.lib.oracle_get_poisson_face_dofs(self.h, _i(a))
        return a

    def cell_map(self, which, n_cells):
        a = np.zeros(n_cells, np.int32)
        self.lib.oracle_get_cell_map(self.h, which, _i(a))
        return a

    def call(self, name, *args):
        self._check(getattr(self.lib, "oracle_" + name)(self.h, *args))

    def project_initial_conditions(self):
        self.call("project_initial_conditions")

    def assemble_semiconductor_rhs(self):
        self.call("assemble_semiconductor_rhs")

    def assemble_electrolyte_rhs(self):
        self.call("assemble_electrolyte_rhs")

    def solve_full_system(self):
        self.call("solve_full_system")

    def solve_species(self, s):
        self.call("solve_species", int(s))

    def assemble_Poisson_rhs(self):
        self.call("assemble_Poisson_rhs")

    def solve_Poisson(self):
        self.call("solve_Poisson")

    def distribute_Poisson(self):
        self.call("distribute_Poisson")

    def set_solvers(self):
        self.call("set_solvers")

    def step(self, n=1):
        t = np.zeros(5)
        self._check(self.lib.oracle_step(self.h, int(n), _d(t)))
        return t

    def project_test_initial_condition(self):
        self.call("project_test_initial_condition")

    def assemble_test_steady_rhs(self):
        self.call("assemble_test_steady_rhs")

    def assemble_test_transient_rhs(self, t):
        self.call("assemble_test_transient_rhs", C.c_double(t))

    def assemble_coupled_Poisson_test_rhs(self, t):
        self.call("assemble_coupled_Poisson_test_rhs", C.c_double(t))

    def assemble_coupled_DD_test_rhs(self, t):
        self.call("assemble_coupled_DD_test_rhs", C.c_double(t))

    def ldg_errors(self, which, t):
        e = np.zeros(2)
        self._check(self.lib.oracle_ldg_errors(self.h, int(which), C.c_double(t), _d(e)))
        return e

    def mixed_errors(self):
        e = np.zeros(2)
        self._check(self.lib.oracle_mixed_errors(self.h, _d(e)))
        return e
