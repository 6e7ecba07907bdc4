// oracle/pecs_oracle.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// CPU restatement of the reference's per-step IMEX path and of the one-time assembly it depends on, written
// the way the reference is written (per-cell local assemblers with FEValues-style scratch, ordered scatter,
// a direct sparse LU per system).  Every function cites the reference file:line it follows.
// PARITY UNPINNED w.r.t. raw reference vectors: the reference needs deal.II/UMFPACK/TBB, none of which exist in
// this container, and it ships no golden vectors.  What pins this oracle: the reference's own manufactured
// solution tests (analytic solutions, expected L2 orders k+1) -- see tests/test_oracle_convergence.py.
#pragma once
#include <map>
#include <utility>
#include <vector>

#include "fe_values.hpp"
#include "sparse_lu.hpp"

namespace oracle {

enum { Interface = 0, Dirichlet = 1, Neumann = 2, Schottky = 3 };                       // reference Grid.hpp:141-147
enum { FACE_SAME_LEVEL = 0, FACE_BOUNDARY = 1, FACE_HAS_CHILDREN = 2, FACE_COARSER = 3 }; // mesh-table encoding

// indices into the flat parameter array (scaled values, reference include/Parameters.hpp:181-242)
enum ParamIndex {
  P_DELTA_T = 0, P_PENALTY, P_MU_N, P_MU_P, P_MU_R, P_MU_O, P_EPS_S, P_EPS_E, P_LAMBDA2, P_K_ET, P_K_HT, P_V_N, P_V_P,
  P_GEN_FLUX, P_GEN_ALPHA, P_GEN_LOCATION, P_RHO_N_E, P_RHO_P_E, P_RHO_R_E, P_RHO_O_E, P_PHI_BI, P_PHI_APP, P_PHI_SCH,
  P_SCH_LOCATION, P_TRANSIENT, P_SRH, P_N_INTRINSIC, P_TAU_N, P_TAU_P, P_COUNT
};

struct Mesh {
  int n_cells = 0;
  std::vector<double> vtx;  // [n][4][2]
  std::vector<int> material, face_kind, neighbor, neighbor2, boundary_id;
  std::vector<double> nb_parent_diameter;
  const double* v(int c) const { return &vtx[8 * (size_t)c]; }
  double diameter(int c) const;
  void center(int c, double& x, double& y) const;
  void face_center(int c, int f, double& x, double& y) const;
};

// row-wise dynamic sparse matrix (dealii::SparseMatrix stand-in)
struct SparseMatrix {
  int n = 0;
  std::vector<std::vector<std::pair<int, double>>> rows;
  void reinit(int n_) { n = n_; rows.assign(n_, {}); }
  void add(int i, int j, double v);
  void vmult(std::vector<double>& y, const std::vector<double>& x) const;
  size_t nnz() const;
  void to_csr(std::vector<int>& rp, std::vector<int>& col, std::vector<double>& val) const;
  CscMatrix to_csc() const;
};

// ConstraintMatrix stand-in: x[dof] = weight * x[master] (master < 0: x[dof] = 0)
struct Constraints {
  struct Line { int master; double weight; };
  std::map<int, Line> lines;
  bool is_constrained(int dof) const { return lines.count(dof) != 0; }
  void distribute_local_to_global(const std::vector<double>& local, const std::vector<int>& dofs,
                                  std::vector<double>& global) const;
  void distribute_local_to_global(const std::vector<std::vector<double>>& local, const std::vector<int>& dofs,
                                  SparseMatrix& global) const;
  void distribute(std::vector<double>& x) const;
};

struct Carrier { // reference include/Carrier.hpp:57-94
  SparseMatrix system_matrix;
  std::vector<double> system_rhs, solution;
  SparseLU solver;
  double scaled_mobility = 1.0, charge_number = 0.0;
  void set_solver(const std::vector<int>& order) { solver.initialize(system_matrix.to_csc(), order); }
  void solve() { solver.vmult(solution.data(), system_rhs.data()); }
};

struct CarrierPair { // reference include/CarrierPair.hpp:72-108
  Carrier carrier_1, carrier_2;
  SparseMatrix mass_matrix;
  double penalty = 1.0, material_permittivity = 1.0;
  int n_cells = 0;
  std::vector<int> elimination_order;
  void setup_dofs(const Mesh& mesh);
  // component-wise renumbering: [Jx | Jy | rho], 4 per cell (reference CarrierPair.cpp:29-33)
  void get_dof_indices(int cell, std::vector<int>& idx) const {
    for (int i = 0; i < 12; ++i) idx[i] = (i / 4) * 4 * n_cells + 4 * cell + (i % 4);
  }
};

struct PoissonData { // reference include/Poisson.hpp:44-83
  SparseMatrix system_matrix;
  std::vector<double> system_rhs, solution;
  Constraints constraints;
  SparseLU solver;
  int n_cells = 0, n_rt = 0;
  std::vector<int> face_dof; // [n][4]
  std::vector<int> elimination_order;
  void setup_dofs(const Mesh& mesh);
  void get_dof_indices(int cell, std::vector<int>& idx) const {
    for (int f = 0; f < 4; ++f) idx[f] = face_dof[4 * cell + f];
    idx[4] = n_rt + cell;
  }
  void set_solver() { solver.initialize(system_matrix.to_csc(), elimination_order); }
  void solve() {
    solver.vmult(solution.data(), system_rhs.data());
    constraints.distribute(solution);
  }
};

class SolarCellProblem {
public:
  double prm[P_COUNT] = {};
  bool full_system = true;
  Mesh semiconductor_mesh, electrolyte_mesh, Poisson_mesh;
  CarrierPair electron_hole_pair, redox_pair;
  PoissonData Poisson_object;
  double delta_t = 0.0;

  // cell maps built by centre matching (reference source/SolarCell.cpp:156-372)
  std::map<int, int> s_2_p_map, e_2_p_map;
  std::vector<int> semi_interface_cells, semi_interface_faces, elec_interface_cells, elec_interface_faces;
  std::map<int, int> semi_interface_map, elec_interface_map;

  void setup_dofs();
  void setup_mappings();
  void assemble_Poisson_matrix();
  void assemble_LDG_system(double transient_or_steady);
  void set_solvers();
  void project_initial_conditions(); // constants rho_e in the density dofs (reference SolarCell.cpp:1999-2021)

  // the hot path (reference source/SolarCell.cpp:2057-2075)
  void assemble_semiconductor_rhs();
  void assemble_electrolyte_rhs();
  void solve_full_system();
  void assemble_Poisson_rhs();
  void solve_Poisson();

  // manufactured-solution variants (reference SolarCell.cpp:2749-3117, LDG.cpp:681-982, MixedFEM.cpp:166-254)
  void project_test_initial_condition();
  void assemble_test_steady_rhs();
  void assemble_test_transient_rhs(double time);
  void assemble_coupled_Poisson_test_rhs(double time);
  void assemble_coupled_DD_test_rhs(double time);
  void ldg_errors(int which_solution, double time, double& density_error, double& current_error) const;
  void mixed_errors(double& potential_error, double& field_error) const;

private:
  void assemble_local_LDG(CarrierPair& pair, const Mesh& mesh, double transient_or_steady);
  void assemble_flux_terms(CarrierPair& pair, const Mesh& mesh);
  void assemble_local_semiconductor_rhs(int cell, std::vector<double>& rhs1, std::vector<double>& rhs2) const;
  void assemble_local_electrolyte_rhs(int cell, std::vector<double>& rhs1, std::vector<double>& rhs2) const;
  void assemble_local_Poisson_rhs(int cell, bool semiconductor, std::vector<int>& dofs, std::vector<double>& rhs) const;
  double generation(const Tensor1& p) const;
};

} // namespace oracle
