// Carrier.cpp -- see Carrier.hpp.
#include "Carrier.hpp"

#include <fstream>
#include <iostream>
#include <stdexcept>

#include "../error.hpp"

namespace {
void check(pecs_status s, const char* what) {
  if (s != PECS_OK) throw pecs::StatusError(s, std::string(what) + ": " + pecs_last_error());
}
void require_ctx(const pecs_ctx* ctx, const char* who) {
  if (!ctx)
    throw std::runtime_error(std::string(who) +
                             ": no device context (SolarCellProblem::set_solvers() must run on a CUDA device first; "
                             "there is no CPU fallback)");
}

// dealii::Vector<double>::block_write / block_read layout: "<size>\n[" + raw doubles + "]"
void block_write(const std::string& file, const std::vector<double>& v) {
  std::ofstream out(file.c_str(), std::ios::binary);
  out << v.size() << "\n[";
  out.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(double)));
  out << "]";
  out.flush();
  if (!out) throw std::runtime_error("cannot write restart file " + file); // a failed checkpoint must not go unnoticed
}
void block_read(const std::string& file, std::vector<double>& v) {
  std::ifstream in(file.c_str(), std::ios::binary);
  if (!in) throw std::runtime_error("cannot open restart file " + file);
  size_t n = 0;
  in >> n;
  char c = 0;
  in.get(c); // '\n'
  in.get(c); // '['
  if (c != '[' || n != v.size()) throw std::runtime_error("restart file " + file + " does not match this mesh");
  in.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(double)));
  // a truncated file must not be accepted silently: all bytes there, and the closing bracket behind them
  if ((size_t)in.gcount() != n * sizeof(double) || !in.get(c) || c != ']')
    throw std::runtime_error("restart file " + file + " is truncated or corrupt");
}
} // namespace

namespace ChargeCarrierSpace {

void Carrier::set_solver() const { require_ctx(ctx, "Carrier::set_solver"); }
void Carrier::solve() {
  require_ctx(ctx, "Carrier::solve");
  check(pecs_solve_species(ctx, species), "Carrier::solve");
}
void Carrier::pull_solution() {
  require_ctx(ctx, "Carrier::pull_solution");
  check(pecs_get_state(ctx, species, solution.data()), "Carrier::pull_solution");
}
void Carrier::pull_rhs() {
  require_ctx(ctx, "Carrier::pull_rhs");
  check(pecs_get_rhs(ctx, species, system_rhs.data()), "Carrier::pull_rhs");
}
void Carrier::push_solution() const {
  require_ctx(ctx, "Carrier::push_solution");
  check(pecs_set_state(ctx, species, solution.data()), "Carrier::push_solution");
}

void CarrierPair::setup_dofs(const pecs::MeshTables& mesh) {
  dofs.n_cells = mesh.n_cells;
  const size_t n = (size_t)dofs.n_dofs();
  for (Carrier* c : {&carrier_1, &carrier_2}) {
    c->solution.assign(n, 0.0);
    c->system_rhs.assign(n, 0.0);
  }
}
void CarrierPair::print_info() const {
  // reference CarrierPair.cpp:66-87
  std::cout << "Number of DOFS " << material_name << ": " << 2 * dofs.n_dofs() << " = 2 x (" << 8 * dofs.n_cells << " + "
            << 4 * dofs.n_cells << ")" << std::endl;
}
void CarrierPair::print_dofs(const std::string& directory) {
  carrier_1.pull_solution();
  carrier_2.pull_solution();
  block_write(directory + "/" + carrier_1.name + ".dofs", carrier_1.solution);
  block_write(directory + "/" + carrier_2.name + ".dofs", carrier_2.solution);
}
void CarrierPair::read_dofs(const std::string& directory) {
  block_read(directory + "/" + carrier_1.name + ".dofs", carrier_1.solution);
  block_read(directory + "/" + carrier_2.name + ".dofs", carrier_2.solution);
}
void CarrierPair::set_semiconductor_for_testing(double mobility_1, double mobility_2) {
  penalty = 1.0;
  carrier_1.scaled_mobility = mobility_1;
  carrier_2.scaled_mobility = mobility_2;
}

} // namespace ChargeCarrierSpace

namespace Poisson {

void PoissonData::setup_dofs(const pecs::MeshTables& mesh, int neumann_id) {
  dofs = pecs::build_poisson_dofs(mesh, neumann_id);
  solution.assign((size_t)dofs.n_dofs(), 0.0);
  system_rhs.assign((size_t)dofs.n_dofs(), 0.0);
}
void PoissonData::print_info() const {
  std::cout << "Number of DOFS Poisson: " << dofs.n_dofs() << " (" << dofs.n_rt << " + " << dofs.n_cells << ")" << std::endl;
}
void PoissonData::set_solver() const { require_ctx(ctx, "PoissonData::set_solver"); }
void PoissonData::solve() {
  require_ctx(ctx, "PoissonData::solve");
  check(pecs_solve_poisson(ctx), "PoissonData::solve");
}
void PoissonData::pull_solution() {
  require_ctx(ctx, "PoissonData::pull_solution");
  check(pecs_get_state(ctx, PECS_POISSON, solution.data()), "PoissonData::pull_solution");
}
void PoissonData::pull_rhs() {
  require_ctx(ctx, "PoissonData::pull_rhs");
  check(pecs_get_rhs(ctx, PECS_POISSON, system_rhs.data()), "PoissonData::pull_rhs");
}

} // namespace Poisson
