// oracle/capi.cpp -- TEST INFRASTRUCTURE ONLY. extern "C" surface of the CPU oracle for ctypes (tests/, bench.py
// cpu_baseline leg, __graft_entry__.smoke()).  Nothing under pecs_b200/ may link or load this library.
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <string>

#include "pecs_oracle.hpp"

using oracle::SolarCellProblem;

namespace {
thread_local std::string g_error;
template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}
oracle::Carrier& species(SolarCellProblem* p, int s) {
  switch (s) {
    case 0: return p->electron_hole_pair.carrier_1;
    case 1: return p->electron_hole_pair.carrier_2;
    case 2: return p->redox_pair.carrier_1;
    case 3: return p->redox_pair.carrier_2;
  }
  throw std::runtime_error("species must be 0..3");
}
} // namespace

extern "C" {

const char* oracle_last_error() { return g_error.c_str(); }

void* oracle_create(const double* params, int n_params, int full_system) {
  SolarCellProblem* p = new SolarCellProblem();
  for (int i = 0; i < n_params && i < oracle::P_COUNT; ++i) p->prm[i] = params[i];
  p->full_system = full_system != 0;
  p->delta_t = p->prm[oracle::P_DELTA_T];
  p->electron_hole_pair.carrier_1.charge_number = -1.0; // reference SolarCell.cpp:52-61
  p->electron_hole_pair.carrier_2.charge_number = 1.0;
  p->electron_hole_pair.carrier_1.scaled_mobility = p->prm[oracle::P_MU_N];
  p->electron_hole_pair.carrier_2.scaled_mobility = p->prm[oracle::P_MU_P];
  p->electron_hole_pair.material_permittivity = p->prm[oracle::P_EPS_S];
  p->redox_pair.carrier_1.charge_number = -1.0; // reference SolarCell.cpp:69-78
  p->redox_pair.carrier_2.charge_number = 1.0;
  p->redox_pair.carrier_1.scaled_mobility = p->prm[oracle::P_MU_R];
  p->redox_pair.carrier_2.scaled_mobility = p->prm[oracle::P_MU_O];
  p->redox_pair.material_permittivity = p->prm[oracle::P_EPS_E];
  p->electron_hole_pair.penalty = p->redox_pair.penalty = p->prm[oracle::P_PENALTY];
  return p;
}
void oracle_destroy(void* h) { delete static_cast<SolarCellProblem*>(h); }

// which: 0 semiconductor, 1 electrolyte, 2 Poisson
int oracle_set_mesh(void* h, int which, int n_cells, const double* vertices, const int* material, const int* face_kind,
                    const int* neighbor, const int* neighbor2, const int* boundary_id, const double* nb_parent_diameter) {
  return guarded([&] {
    SolarCellProblem* p = static_cast<SolarCellProblem*>(h);
    oracle::Mesh& m = which == 0 ? p->semiconductor_mesh : (which == 1 ? p->electrolyte_mesh : p->Poisson_mesh);
    m.n_cells = n_cells;
    m.vtx.assign(vertices, vertices + 8 * (size_t)n_cells);
    m.material.assign(material, material + n_cells);
    m.face_kind.assign(face_kind, face_kind + 4 * (size_t)n_cells);
    m.neighbor.assign(neighbor, neighbor + 4 * (size_t)n_cells);
    m.neighbor2.assign(neighbor2, neighbor2 + 4 * (size_t)n_cells);
    m.boundary_id.assign(boundary_id, boundary_id + 4 * (size_t)n_cells);
    m.nb_parent_diameter.assign(nb_parent_diameter, nb_parent_diameter + 4 * (size_t)n_cells);
  });
}

// dofs + mappings + matrices (+ factorisation if factor != 0)
int oracle_setup(void* h, double transient_or_steady, int factor) {
  return guarded([&] {
    SolarCellProblem* p = static_cast<SolarCellProblem*>(h);
    p->setup_dofs();
    p->setup_mappings();
    p->assemble_Poisson_matrix();
    p->assemble_LDG_system(transient_or_steady);
    if (factor) p->set_solvers();
  });
}

int oracle_n_dofs(void* h, int which) { // 0..3 species, 4 Poisson
  SolarCellProblem* p = static_cast<SolarCellProblem*>(h);
  if (which == 4) return (int)p->Poisson_object.solution.size();
  return (int)species(p, which).solution.size();
}
int oracle_n_rt(void* h) { return static_cast<SolarCellProblem*>(h)->Poisson_object.n_rt; }

// kind: 0 solution, 1 system_rhs ; which: 0..3 species, 4 Poisson
int oracle_get_vector(void* h, int which, int kind, double* out) {
  return guarded([&] {
    SolarCellProblem* p = static_cast<SolarCellProblem*>(h);
    const std::vector<double>& v = which == 4 ? (kind == 0 ? p->Poisson_object.solution : p->Poisson_object.system_rhs)
                                              : (kind == 0 ? species(p, which).solution : species(p, which).system_rhs);
    std::memcpy(out, v.data(), v.size() * sizeof(double));
  });
}
int oracle_set_vector(void* h, int which, int kind, const double* in) {
  return guarded([&] {
    SolarCellProblem* p = static_cast<SolarCellProblem*>(h);
    std::vector<double>& v = which == 4 ? (kind == 0 ? p->Poisson_object.solution : p->Poisson_object.system_rhs)
                                        : (kind == 0 ? species(p, which).solution : species(p, which).system_rhs);
    std::memcpy(v.data(), in, v.size() * sizeof(double));
  });
}

// matrices: which 0..3 species system matrix, 4 Poisson, 5 semiconductor mass, 6 electrolyte mass
static const oracle::SparseMatrix& matrix_of(SolarCellProblem* p, int which) {
  if (which == 4) return p->Poisson_object.system_matrix;
  if (which == 5) return p->electron_hole_pair.mass_matrix;
  if (which == 6) return p->redox_pair.mass_matrix;
  return species(p, which).system_matrix;
}
long oracle_matrix_nnz(void* h, int which) { return (long)matrix_of(static_cast<SolarCellProblem*>(h), which).nnz(); }
int oracle_get_matrix(void* h, int which, int* row_ptr, int* col, double* val) {
  return guarded([&] {
    std::vector<int> rp, c;
    std::vector<double> v;
    matrix_of(static_cast<SolarCellProblem*>(h), which).to_csr(rp, c, v);
    std::memcpy(row_ptr, rp.data(), rp.size() * sizeof(int));
    std::memcpy(col, c.data(), c.size() * sizeof(int));
    std::memcpy(val, v.data(), v.size() * sizeof(double));
  });
}
int oracle_get_poisson_face_dofs(void* h, int* out) {
  SolarCellProblem* p = static_cast<SolarCellProblem*>(h);
  std::memcpy(out, p->Poisson_object.face_dof.data(), p->Poisson_object.face_dof.size() * sizeof(int));
  return 0;
}
int oracle_get_cell_map(void* h, int which, int* out) { // 0: s_2_p, 1: e_2_p
  SolarCellProblem* p = static_cast<SolarCellProblem*>(h);
  for (const auto& kv : which == 0 ? p->s_2_p_map : p->e_2_p_map) out[kv.first] = kv.second;
  return 0;
}

#define ORACLE_CALL(name, expr) \
  int name(void* h) { return guarded([&] { SolarCellProblem* p = static_cast<SolarCellProblem*>(h); expr; }); }
ORACLE_CALL(oracle_project_initial_conditions, p->project_initial_conditions())
ORACLE_CALL(oracle_assemble_semiconductor_rhs, p->assemble_semiconductor_rhs())
ORACLE_CALL(oracle_assemble_electrolyte_rhs, p->assemble_electrolyte_rhs())
ORACLE_CALL(oracle_solve_full_system, p->solve_full_system())
ORACLE_CALL(oracle_assemble_Poisson_rhs, p->assemble_Poisson_rhs())
ORACLE_CALL(oracle_solve_Poisson, p->solve_Poisson())
ORACLE_CALL(oracle_set_solvers, p->set_solvers())
ORACLE_CALL(oracle_project_test_initial_condition, p->project_test_initial_condition())
ORACLE_CALL(oracle_assemble_test_steady_rhs, p->assemble_test_steady_rhs())

// constraints.distribute(solution) alone (reference source/Poisson.cpp:104): for callers that bring their own direct
// solver for the condensed Poisson matrix (bench.py's CPU arm uses SuperLU as the UMFPACK stand-in)
ORACLE_CALL(oracle_distribute_Poisson, p->Poisson_object.constraints.distribute(p->Poisson_object.solution))

int oracle_solve_species(void* h, int s) {
  return guarded([&] { species(static_cast<SolarCellProblem*>(h), s).solve(); });
}

// n IMEX steps in the production order (reference SolarCell.cpp:2057-2075); section wall times (seconds) are
// accumulated into times[5] under the reference's TimerOutput section names.
int oracle_step(void* h, int n_steps, double* times) {
  return guarded([&] {
    SolarCellProblem* p = static_cast<SolarCellProblem*>(h);
    typedef std::chrono::steady_clock clk;
    auto lap = [&](int k, clk::time_point& t0) {
      const clk::time_point t1 = clk::now();
      if (times) times[k] += std::chrono::duration<double>(t1 - t0).count();
      t0 = t1;
    };
    for (int s = 0; s < n_steps; ++s) {
      clk::time_point t0 = clk::now();
      p->assemble_semiconductor_rhs(); lap(0, t0);
      p->assemble_electrolyte_rhs();   lap(1, t0);
      p->solve_full_system();          lap(2, t0);
      p->assemble_Poisson_rhs();       lap(3, t0);
      p->solve_Poisson();              lap(4, t0);
    }
  });
}

int oracle_assemble_test_transient_rhs(void* h, double time) {
  return guarded([&] { static_cast<SolarCellProblem*>(h)->assemble_test_transient_rhs(time); });
}
int oracle_assemble_coupled_Poisson_test_rhs(void* h, double time) {
  return guarded([&] { static_cast<SolarCellProblem*>(h)->assemble_coupled_Poisson_test_rhs(time); });
}
int oracle_assemble_coupled_DD_test_rhs(void* h, double time) {
  return guarded([&] { static_cast<SolarCellProblem*>(h)->assemble_coupled_DD_test_rhs(time); });
}
int oracle_ldg_errors(void* h, int which, double time, double* out2) {
  return guarded([&] { static_cast<SolarCellProblem*>(h)->ldg_errors(which, time, out2[0], out2[1]); });
}
int oracle_mixed_errors(void* h, double* out2) {
  return guarded([&] { static_cast<SolarCellProblem*>(h)->mixed_errors(out2[0], out2[1]); });
}

} // extern "C"
