// Carrier.hpp / CarrierPair / PoissonData -- the data holders of the reference, host side.
//
// Mirrors reference include/Carrier.hpp:57-94, include/CarrierPair.hpp:72-108 and include/Poisson.hpp:44-83:
// plain structs with public members named as in the reference (system_matrix, system_rhs, solution,
// mass_matrix, constraints, penalty, ...).  The difference is where the per-step state lives: `solution` and
// `system_rhs` are host MIRRORS; the live copies are in HBM inside the pecs_ctx, and set_solver()/solve() forward
// to the C ABI (include/pecs_b200.h).  pull()/push() move a vector between the mirror and the device.
#pragma once
#include <string>
#include <vector>

#include "../../../include/pecs_b200.h"
#include "Csr.hpp"
#include "DoFTables.hpp"
#include "Triangulation.hpp"

namespace ChargeCarrierSpace {

struct Carrier {
  pecs::CsrMatrix system_matrix;
  std::vector<double> system_rhs; // host mirror
  std::vector<double> solution;   // host mirror
  double scaled_mobility = 1.0;
  double charge_number = 0.0;
  std::string name;

  // device binding
  pecs_ctx* ctx = nullptr;
  int species = -1;

  void set_name(const std::string& str_name) { name = str_name; }
  // reference Carrier.cpp:26-32 factorises here; on the device all factorisations happen together inside
  // pecs_ctx_create (SolarCellProblem::set_solvers), so this only checks that the binding exists.
  void set_solver() const;
  // reference Carrier.cpp:34-40: solution = A^-1 system_rhs (on the device, asynchronous)
  void solve();
  void pull_solution();       // device -> solution
  void pull_rhs();            // device -> system_rhs
  void push_solution() const; // solution -> device
};

struct CarrierPair {
  Carrier carrier_1, carrier_2;
  pecs::CsrMatrix mass_matrix;
  pecs::CarrierDofs dofs; // DG: no constraints (reference CarrierPair.cpp:62-63)
  double penalty = 1.0;
  double material_permittivity = 1.0;
  std::string material_name;

  void set_name(const std::string& str_name) { material_name = str_name; }
  // reference CarrierPair.cpp:23-64
  void setup_dofs(const pecs::MeshTables& mesh);
  void print_info() const;
  // restart files, reference CarrierPair.cpp:89-121 (Vector::block_write / block_read layout)
  void print_dofs(const std::string& directory = ".");
  void read_dofs(const std::string& directory = ".");
  // reference CarrierPair.cpp:124-142
  void set_semiconductor_for_testing(double mobility_1, double mobility_2);
};

} // namespace ChargeCarrierSpace

namespace Poisson {

struct PoissonData {
  pecs::CsrMatrix system_matrix;
  std::vector<double> system_rhs; // host mirror
  std::vector<double> solution;   // host mirror
  pecs::PoissonDofs dofs;         // numbering + constraints (hanging edges, Neumann edges)
  pecs_ctx* ctx = nullptr;

  // reference Poisson.cpp:18-69
  void setup_dofs(const pecs::MeshTables& mesh, int neumann_id);
  void print_info() const;
  void set_solver() const;
  // reference Poisson.cpp:98-105: solver.vmult + constraints.distribute, on the device
  void solve();
  void pull_solution();
  void pull_rhs();
};

} // namespace Poisson
