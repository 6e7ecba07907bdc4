// SolarCell.hpp -- SOLARCELL::SolarCellProblem, the problem orchestrator, host side.
//
// Mirror of reference include/SolarCell.hpp:105-757 / source/SolarCell.cpp.  Same public entry points
// (constructor, run_full_system, test_steady_state, test_transient, test_DD_Poisson) and the same hot-path
// method names; what differs is what runs underneath:
//   * one-time setup (grids, dofs, mappings, the constant matrices) is own host code producing flat tables;
//   * set_solvers() hands those tables to pecs_ctx_create (include/pecs_b200.h), which factorises on the device;
//   * the five methods the time loop calls are thin forwards to the C ABI -- all arithmetic is in CUDA kernels.
// The five hot-path methods are private in the reference (include/SolarCell.hpp:530-703); they are public here so
// that the Python binding, the parity tests and bench.py can drive them one by one.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/pecs_b200.h"
#include "Carrier.hpp"
#include "Grid.hpp"
#include "LDG.hpp"
#include "MixedFEM.hpp"
#include "Parameters.hpp"

namespace SOLARCELL {

// stand-in for dealii::ConvergenceTable: named columns of numbers, printed with reduction rates
class ConvergenceTable {
public:
  void add_value(const std::string& key, double v);
  const std::vector<double>& column(const std::string& key) const;
  void write_text(std::ostream& out, const std::vector<std::string>& rate_columns) const;

private:
  std::vector<std::string> order_;
  std::map<std::string, std::vector<double>> cols_;
};

class SolarCellProblem {
public:
  SolarCellProblem(const unsigned int degree, ParameterSpace::ParameterHandler& param);
  ~SolarCellProblem();
  SolarCellProblem(const SolarCellProblem&) = delete;
  SolarCellProblem& operator=(const SolarCellProblem&) = delete;

  // ---- reference public interface (include/SolarCell.hpp:111-301) ----
  void run_full_system();
  void test_steady_state(const unsigned int& n_refine, ConvergenceTable& Mixed_table, ConvergenceTable& LDG_table);
  void test_transient(const unsigned int& n_refine, ConvergenceTable& LDG_table);
  void test_DD_Poisson(const unsigned int& n_refine, ConvergenceTable& Mixed_table, ConvergenceTable& LDG_table);

  // ---- staged setup (what run_full_system / the tests do before their time loops) ----
  void setup_full_system();                             // reference SolarCell.cpp:1898-2034
  void setup_test(int kind, unsigned int n_refine);     // kind = PECS_KIND_TEST_*
  void setup_full_system_host();                        // ... the part of it that needs no device (tables + matrices)
  void setup_test_host(int kind, unsigned int n_refine);
  void setup_dofs();                                    // reference SolarCell.cpp:105-121
  void setup_mappings();                                // reference SolarCell.cpp:156-372 (hashed, not O(N^2))
  void assemble_Poisson_matrix();                       // reference SolarCell.cpp:377-406
  void assemble_LDG_system(const double& transient_or_steady); // reference SolarCell.cpp:822-932
  void set_solvers();                                   // reference SolarCell.cpp:1733-1747 -> pecs_ctx_create
  void project_initial_conditions();                    // reference SolarCell.cpp:1999-2021
  void project_test_initial_condition();                // reference SolarCell.cpp:2898-2902

  // ---- the hot path (reference SolarCell.cpp:2057-2075); all forward to the device ----
  void assemble_semiconductor_rhs();
  void assemble_electrolyte_rhs();
  void solve_full_system();
  void assemble_Poisson_rhs();
  void solve_Poisson();
  void step(int n_steps); // n iterations of the loop body through the captured CUDA graph
  // reference SolarCell.cpp:1826-1858: the three rescaled .vtu files of one time stamp.  The patch values are computed
  // and rescaled on the device and leave on the context's output stream; the files are written by a background thread.
  // Returns at once: the time loop goes on while the output drains (finish_output() waits for it).
  void print_results(unsigned int time_step_number);
  void finish_output();
  void set_time(double time);
  void synchronize();

  // ---- I-V post-processing (the reference has no counterpart: `applied bias` is one scalar) ----
  // Charge-transfer currents through the semiconductor-electrolyte interface, the integrals of the two interface terms
  // the step assembles (reference SolarCell.cpp:1265-1347, 1692-1712), in scaled units:
  //   out[0] = int_Sigma k_et (rho_n - rho_n^e) rho_o ds  (electron transfer),
  //   out[1] = int_Sigma k_ht (rho_p - rho_p^e) rho_r ds  (hole transfer)
  // from host state vectors (electrons, holes, reductants, oxidants); NULL: the current device state is downloaded.
  void interface_currents(const double* const states[4], double out[2]);

  // ---- post-processing of the manufactured tests (host, after download) ----
  void ldg_errors(int which, double time, double& density_error, double& current_error);
  void mixed_errors(double& potential_error, double& field_error);

  // scaled parameters in PECS_P_* order
  void fill_params(double params[32]) const;

  // ---- state (public like the reference's L2 structs) ----
  const unsigned int degree;
  ParameterSpace::ParameterHandler& prm;
  ParameterSpace::Parameters sim_params;
  bool full_system = true;
  int kind = PECS_KIND_PRODUCTION;
  int device = 0;
  int owned_species = 0; // bit k: carrier k is solved by this process' context; 0 = all (pecs_problem_desc::owned_species)
  double delta_t = 0.0;
  bool verbose = false;
  bool write_output = true;           // run_full_system writes the reference's .vtu files
  std::string output_directory = "."; // where print_results / print_dofs put them

  pecs::Triangulation Poisson_triangulation, semiconductor_triangulation, electrolyte_triangulation;
  Poisson::PoissonData Poisson_object;
  ChargeCarrierSpace::CarrierPair electron_hole_pair, redox_pair;
  MixedPoisson::MixedFEM Mixed_Assembler;
  LDG_System::LDG LDG_Assembler;

  // mappings as flat tables (reference keeps std::maps keyed by (level,index), SolarCell.hpp:385-406)
  std::vector<int> s_2_p_map, e_2_p_map;
  std::vector<int> semi_interface_cells, semi_interface_faces, elec_interface_cells, elec_interface_faces;

  pecs_ctx* ctx = nullptr;

private:
  // output path state: geometry of the three patch meshes, a ring of two page-locked host slots, the writer thread
  struct OutputState;
  std::unique_ptr<OutputState> output_;
  struct BoundaryFaces {
    std::vector<int> cell, face, id;
  };
  static BoundaryFaces boundary_faces(const pecs::MeshTables& mesh);
  void release_ctx();
};

} // namespace SOLARCELL
