// SolarCell.cpp -- see SolarCell.hpp.
#include "SolarCell.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <future>
#include <iomanip>
#include <iostream>
#include <stdexcept>

#include "../error.hpp"
#include <unordered_map>

#include "../fe.hpp"
#include "../rhs_math.hpp"
#include "../test_functions.hpp"
#include "SolverSetup.hpp"

namespace SOLARCELL {

namespace {
void check(pecs_status s, const char* what) {
  if (s != PECS_OK) throw pecs::StatusError(s, std::string(what) + ": " + pecs_last_error());
}
void require_ctx(const pecs_ctx* ctx, const char* who) {
  if (!ctx) throw std::runtime_error(std::string(who) + ": set_solvers() has not created a device context");
}
pecs_csr view(const pecs::CsrMatrix& A) { return pecs_csr{A.n, A.row_ptr.data(), A.col.data(), A.val.data()}; }

// exact-coordinate hash of a point (cell / face centres of matching cells are computed from identical vertex
// coordinates with identical arithmetic, so they are bit-identical; the reference's 1e-13 tolerance,
// SolarCell.cpp:216, 315, is checked afterwards)
struct PointKey {
  double x, y;
  bool operator==(const PointKey& o) const { return x == o.x && y == o.y; }
};
struct PointHash {
  size_t operator()(const PointKey& k) const {
    std::uint64_t a, b;
    std::memcpy(&a, &k.x, 8);
    std::memcpy(&b, &k.y, 8);
    return (size_t)(a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull + (a << 6) + (a >> 2)));
  }
};
pecs::fe::CellVerts verts_of(const pecs::MeshTables& mesh, int c) {
  pecs::fe::CellVerts v;
  for (int a = 0; a < 4; ++a) {
    v.x[a] = mesh.vtx(c)[2 * a];
    v.y[a] = mesh.vtx(c)[2 * a + 1];
  }
  return v;
}
} // namespace

// ------------------------------------------------------------------------------------------- ConvergenceTable
void ConvergenceTable::add_value(const std::string& key, double v) {
  if (!cols_.count(key)) order_.push_back(key);
  cols_[key].push_back(v);
}
const std::vector<double>& ConvergenceTable::column(const std::string& key) const { return cols_.at(key); }
void ConvergenceTable::write_text(std::ostream& out, const std::vector<std::string>& rate_columns) const {
  for (const std::string& k : order_) {
    out << std::setw(14) << k;
    if (std::find(rate_columns.begin(), rate_columns.end(), k) != rate_columns.end()) out << std::setw(8) << "rate";
  }
  out << "\n";
  const size_t rows = order_.empty() ? 0 : cols_.at(order_[0]).size();
  for (size_t r = 0; r < rows; ++r) {
    for (const std::string& k : order_) {
      const std::vector<double>& c = cols_.at(k);
      out << std::setw(14) << std::setprecision(4) << std::scientific << c[r];
      if (std::find(rate_columns.begin(), rate_columns.end(), k) != rate_columns.end()) {
        if (r == 0)
          out << std::setw(8) << "-";
        else
          out << std::setw(8) << std::fixed << std::setprecision(2) << std::log2(c[r - 1] / c[r]);
      }
    }
    out << "\n";
  }
}

// ------------------------------------------------------------------------------------------- construction
struct SolarCellProblem::OutputState {
  std::unique_ptr<pecs::VtuMesh> mesh[3];
  double* slot[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  unsigned long ticket[2] = {0, 0}; // writer job that still reads the slot
  int next = 0;
  pecs::OutputQueue queue; // declared last: destroyed (drained) first
  ~OutputState() {
    queue.wait_idle();
    for (auto& s : slot)
      for (double* p : s)
        if (p) pecs_host_free(p);
  }
};

SolarCellProblem::SolarCellProblem(const unsigned int degree_, ParameterSpace::ParameterHandler& param)
    : degree(degree_), prm(param) {
  if (degree != 1) throw std::runtime_error("SolarCellProblem: only degree 1 is built (reference main.cpp:11 uses 1)");
  sim_params.parse_and_scale_parameters(prm);
  // names, charge signs, mobilities, permittivities: reference SolarCell.cpp:52-85
  electron_hole_pair.carrier_1.set_name("Electrons");
  electron_hole_pair.carrier_1.charge_number = -1.0;
  electron_hole_pair.carrier_1.scaled_mobility = sim_params.scaled_electron_mobility;
  electron_hole_pair.carrier_2.set_name("Holes");
  electron_hole_pair.carrier_2.charge_number = 1.0;
  electron_hole_pair.carrier_2.scaled_mobility = sim_params.scaled_hole_mobility;
  electron_hole_pair.set_name("Semiconductor-");
  electron_hole_pair.material_permittivity = sim_params.semiconductor_permittivity;
  redox_pair.carrier_1.set_name("Reductants");
  redox_pair.carrier_1.charge_number = -1.0;
  redox_pair.carrier_1.scaled_mobility = sim_params.scaled_reductant_mobility;
  redox_pair.carrier_2.set_name("Oxidants");
  redox_pair.carrier_2.charge_number = 1.0;
  redox_pair.carrier_2.scaled_mobility = sim_params.scaled_oxidant_mobility;
  redox_pair.set_name("Electrolyte-");
  redox_pair.material_permittivity = sim_params.electrolyte_permittivity;
}

SolarCellProblem::~SolarCellProblem() { release_ctx(); }

void SolarCellProblem::release_ctx() {
  output_.reset(); // drains the writer thread, which still uses the context
  if (ctx) pecs_ctx_destroy(ctx);
  ctx = nullptr;
  for (ChargeCarrierSpace::Carrier* c : {&electron_hole_pair.carrier_1, &electron_hole_pair.carrier_2,
                                          &redox_pair.carrier_1, &redox_pair.carrier_2})
    c->ctx = nullptr;
  Poisson_object.ctx = nullptr;
}

void SolarCellProblem::fill_params(double p[32]) const {
  std::fill(p, p + 32, 0.0);
  p[PECS_P_DELTA_T] = delta_t;
  p[PECS_P_PENALTY] = electron_hole_pair.penalty;
  p[PECS_P_MU_N] = electron_hole_pair.carrier_1.scaled_mobility;
  p[PECS_P_MU_P] = electron_hole_pair.carrier_2.scaled_mobility;
  p[PECS_P_MU_R] = redox_pair.carrier_1.scaled_mobility;
  p[PECS_P_MU_O] = redox_pair.carrier_2.scaled_mobility;
  p[PECS_P_EPS_S] = sim_params.semiconductor_permittivity;
  p[PECS_P_EPS_E] = sim_params.electrolyte_permittivity;
  p[PECS_P_LAMBDA2] = sim_params.scaled_debeye_length;
  p[PECS_P_K_ET] = sim_params.scaled_k_et;
  p[PECS_P_K_HT] = sim_params.scaled_k_ht;
  p[PECS_P_V_N] = sim_params.scaled_electron_recombo_v;
  p[PECS_P_V_P] = sim_params.scaled_hole_recombo_v;
  // Generation::set_illuminated_params / set_dark_params, reference Generation.cpp:5-24
  const bool lit = sim_params.illum_or_dark;
  p[PECS_P_GEN_FLUX] = lit ? sim_params.scaled_photon_flux : 0.0;
  p[PECS_P_GEN_ALPHA] = lit ? sim_params.scaled_absorption_coeff : 0.0;
  p[PECS_P_GEN_LOCATION] = lit ? sim_params.scaled_domain_height : 0.0;
  // Electrons_/Holes_/Reductants_/Oxidants_Equilibrium, reference InitialConditions.cpp:20,43,61,79
  p[PECS_P_RHO_N_E] = 2.0;
  p[PECS_P_RHO_P_E] = 0.0;
  p[PECS_P_RHO_R_E] = 30.0;
  p[PECS_P_RHO_O_E] = 29.0;
  p[PECS_P_PHI_BI] = sim_params.scaled_built_in_bias;
  p[PECS_P_PHI_APP] = sim_params.scaled_applied_bias;
  p[PECS_P_PHI_SCH] = sim_params.scaled_schottky_bias;
  // the reference never calls Schottky_Bias::set_location; the intended location is the top
  p[PECS_P_SCH_LOCATION] = sim_params.scaled_domain_height;
  p[PECS_P_TRANSIENT] = (kind == PECS_KIND_TEST_STEADY) ? 0.0 : 1.0;
  p[PECS_P_SRH] = (kind == PECS_KIND_PRODUCTION && sim_params.srh_recombination) ? 1.0 : 0.0;
  p[PECS_P_N_INTRINSIC] = sim_params.scaled_intrinsic_density;
  p[PECS_P_TAU_N] = sim_params.scaled_electron_recombo_t;
  p[PECS_P_TAU_P] = sim_params.scaled_hole_recombo_t;
}

// ------------------------------------------------------------------------------------------- setup pieces
void SolarCellProblem::setup_dofs() {
  Poisson_object.setup_dofs(Poisson_triangulation.tables(), Grid_Maker::Neumann);
  electron_hole_pair.setup_dofs(semiconductor_triangulation.tables());
  if (full_system) redox_pair.setup_dofs(electrolyte_triangulation.tables());
}

void SolarCellProblem::setup_mappings() {
  const pecs::MeshTables& P = Poisson_triangulation.tables();
  const pecs::MeshTables& S = semiconductor_triangulation.tables();
  std::unordered_map<PointKey, int, PointHash> centres;
  centres.reserve(2 * (size_t)P.n_cells);
  for (int c = 0; c < P.n_cells; ++c) {
    const pecs::Point2 p = P.center(c);
    centres[PointKey{p.x, p.y}] = c;
  }
  auto match_cells = [&](const pecs::MeshTables& M, std::vector<int>& out) {
    out.assign(M.n_cells, -1);
    for (int c = 0; c < M.n_cells; ++c) {
      const pecs::Point2 p = M.center(c);
      auto it = centres.find(PointKey{p.x, p.y});
      if (it == centres.end()) throw std::runtime_error("setup_mappings: a carrier cell has no Poisson cell");
      out[c] = it->second;
    }
  };
  match_cells(S, s_2_p_map);
  semi_interface_cells.clear();
  semi_interface_faces.clear();
  elec_interface_cells.clear();
  elec_interface_faces.clear();
  if (!full_system) return;
  const pecs::MeshTables& E = electrolyte_triangulation.tables();
  match_cells(E, e_2_p_map);
  std::unordered_map<PointKey, std::pair<int, int>, PointHash> elec_faces;
  for (int c = 0; c < E.n_cells; ++c)
    for (int f = 0; f < 4; ++f)
      if (E.face_kind[4 * c + f] == pecs::FACE_BOUNDARY && E.boundary_id[4 * c + f] == Grid_Maker::Interface) {
        const pecs::Point2 p = E.face_center(c, f);
        elec_faces[PointKey{p.x, p.y}] = std::make_pair(c, f);
      }
  for (int c = 0; c < S.n_cells; ++c)
    for (int f = 0; f < 4; ++f)
      if (S.face_kind[4 * c + f] == pecs::FACE_BOUNDARY && S.boundary_id[4 * c + f] == Grid_Maker::Interface) {
        const pecs::Point2 p = S.face_center(c, f);
        auto it = elec_faces.find(PointKey{p.x, p.y});
        if (it == elec_faces.end()) throw std::runtime_error("setup_mappings: unmatched interface face");
        semi_interface_cells.push_back(c);
        semi_interface_faces.push_back(f);
        elec_interface_cells.push_back(it->second.first);
        elec_interface_faces.push_back(it->second.second);
      }
}

void SolarCellProblem::assemble_Poisson_matrix() {
  Poisson_object.system_matrix = Mixed_Assembler.assemble_Poisson_matrix(
      Poisson_triangulation.tables(), Poisson_object.dofs, sim_params.semiconductor_permittivity,
      sim_params.electrolyte_permittivity, sim_params.scaled_debeye_length);
}

void SolarCellProblem::assemble_LDG_system(const double& transient_or_steady) {
  auto do_pair = [&](ChargeCarrierSpace::CarrierPair& pair, const pecs::MeshTables& mesh) {
    pair.mass_matrix = LDG_Assembler.assemble_mass_matrix(mesh, delta_t);
    LDG_Assembler.assemble_system_matrices(mesh, Grid_Maker::Dirichlet, pair.carrier_1.scaled_mobility,
                                           pair.carrier_2.scaled_mobility, delta_t, transient_or_steady, pair.penalty,
                                           pair.carrier_1.system_matrix, pair.carrier_2.system_matrix);
  };
  // the two subdomains share nothing: assembled side by side (each with its own team of threads)
  std::future<void> electrolyte;
  if (full_system)
    electrolyte = std::async(std::launch::async, [&] { do_pair(redox_pair, electrolyte_triangulation.tables()); });
  do_pair(electron_hole_pair, semiconductor_triangulation.tables());
  if (full_system) electrolyte.get();
}

SolarCellProblem::BoundaryFaces SolarCellProblem::boundary_faces(const pecs::MeshTables& mesh) {
  BoundaryFaces b;
  for (int c = 0; c < mesh.n_cells; ++c)
    for (int f = 0; f < 4; ++f)
      if (mesh.face_kind[4 * c + f] == pecs::FACE_BOUNDARY) {
        b.cell.push_back(c);
        b.face.push_back(f);
        b.id.push_back(mesh.boundary_id[4 * c + f]);
      }
  return b;
}

void SolarCellProblem::set_solvers() {
  release_ctx();
  pecs_problem_desc d;
  std::memset(&d, 0, sizeof(d));
  d.kind = kind;
  d.full_system = full_system ? 1 : 0;
  d.device = device;
  d.owned_species = owned_species;
  fill_params(d.params);

  const pecs::MeshTables& S = semiconductor_triangulation.tables();
  const pecs::MeshTables& P = Poisson_triangulation.tables();
  const BoundaryFaces sb = boundary_faces(S), pb = boundary_faces(P);
  BoundaryFaces eb;
  auto fill_domain = [](pecs_domain_desc& dd, const pecs::MeshTables& M, const std::vector<int>& map,
                        const BoundaryFaces& b, const ChargeCarrierSpace::CarrierPair& pair) {
    dd.n_cells = M.n_cells;
    dd.vertices = M.vertices.data();
    dd.poisson_cell = map.data();
    dd.n_boundary_faces = (int)b.cell.size();
    dd.bface_cell = b.cell.data();
    dd.bface_face = b.face.data();
    dd.bface_id = b.id.data();
    dd.system_matrix[0] = view(pair.carrier_1.system_matrix);
    dd.system_matrix[1] = view(pair.carrier_2.system_matrix);
  };
  fill_domain(d.semiconductor, S, s_2_p_map, sb, electron_hole_pair);
  if (full_system) {
    const pecs::MeshTables& E = electrolyte_triangulation.tables();
    eb = boundary_faces(E);
    fill_domain(d.electrolyte, E, e_2_p_map, eb, redox_pair);
  }
  d.poisson.n_cells = P.n_cells;
  d.poisson.vertices = P.vertices.data();
  d.poisson.n_rt = Poisson_object.dofs.n_rt;
  d.poisson.face_dof = Poisson_object.dofs.face_dof.data();
  d.poisson.n_boundary_faces = (int)pb.cell.size();
  d.poisson.bface_cell = pb.cell.data();
  d.poisson.bface_face = pb.face.data();
  d.poisson.bface_id = pb.id.data();
  d.poisson.system_matrix = view(Poisson_object.system_matrix);
  std::vector<int> c_dof, c_master;
  std::vector<double> c_weight;
  for (const pecs::ConstraintLine& l : Poisson_object.dofs.constraints) {
    c_dof.push_back(l.dof);
    c_master.push_back(l.master);
    c_weight.push_back(l.weight);
  }
  d.poisson.n_constraints = (int)c_dof.size();
  d.poisson.constraint_dof = c_dof.data();
  d.poisson.constraint_master = c_master.data();
  d.poisson.constraint_weight = c_weight.data();
  d.interface_pairs.n_pairs = (int)semi_interface_cells.size();
  d.interface_pairs.semi_cell = semi_interface_cells.data();
  d.interface_pairs.semi_face = semi_interface_faces.data();
  d.interface_pairs.elec_cell = elec_interface_cells.data();
  d.interface_pairs.elec_face = elec_interface_faces.data();

  check(pecs_ctx_create(&d, &ctx), "SolarCellProblem::set_solvers");
  electron_hole_pair.carrier_1.ctx = ctx;
  electron_hole_pair.carrier_1.species = PECS_ELECTRONS;
  electron_hole_pair.carrier_2.ctx = ctx;
  electron_hole_pair.carrier_2.species = PECS_HOLES;
  redox_pair.carrier_1.ctx = ctx;
  redox_pair.carrier_1.species = PECS_REDUCTANTS;
  redox_pair.carrier_2.ctx = ctx;
  redox_pair.carrier_2.species = PECS_OXIDANTS;
  Poisson_object.ctx = ctx;
  // the reference factorises per object (SolarCell.cpp:1738-1746); the binding check stands in for that
  Poisson_object.set_solver();
  electron_hole_pair.carrier_1.set_solver();
  electron_hole_pair.carrier_2.set_solver();
  if (full_system) {
    redox_pair.carrier_1.set_solver();
    redox_pair.carrier_2.set_solver();
  }
}

void SolarCellProblem::project_initial_conditions() {
  // VectorTools::project of a constant onto DG = the constant in the density dofs, 0 in the currents
  auto fill = [](ChargeCarrierSpace::Carrier& c, int n_cells, double v) {
    std::fill(c.solution.begin(), c.solution.end(), 0.0);
    std::fill(c.solution.begin() + 8 * (size_t)n_cells, c.solution.end(), v);
  };
  fill(electron_hole_pair.carrier_1, electron_hole_pair.dofs.n_cells, 2.0);
  fill(electron_hole_pair.carrier_2, electron_hole_pair.dofs.n_cells, 0.0);
  if (full_system) {
    fill(redox_pair.carrier_1, redox_pair.dofs.n_cells, 30.0);
    fill(redox_pair.carrier_2, redox_pair.dofs.n_cells, 29.0);
  }
}

void SolarCellProblem::project_test_initial_condition() {
  // L2 projection with QGauss(2) onto Q1 per cell.  With 4 Gauss points and 4 basis functions the projection is
  // collocation at the Gauss points: u = B^-1 f(x_q), B_qa = N_a(x_q) = b (x) b with b the 2x2 1-D table, and
  // b^-1 = sqrt(3) [[1-g, -g], [-g, 1-g]], g = (1 - 1/sqrt 3)/2 -- independent of the cell geometry.
  const pecs::MeshTables& S = semiconductor_triangulation.tables();
  std::vector<double>& u = electron_hole_pair.carrier_1.solution;
  std::fill(u.begin(), u.end(), 0.0);
  const double g = 0.5 - 0.5 / std::sqrt(3.0), r3 = std::sqrt(3.0);
  const double gp[2] = {g, 1.0 - g};
  const double binv[2][2] = {{r3 * (1 - g), -r3 * g}, {-r3 * g, r3 * (1 - g)}};
  for (int c = 0; c < S.n_cells; ++c) {
    const pecs::fe::CellVerts v = verts_of(S, c);
    double f[2][2]; // [qy][qx]
    for (int qy = 0; qy < 2; ++qy)
      for (int qx = 0; qx < 2; ++qx) {
        double x, y;
        pecs::fe::map_point(v, gp[qx], gp[qy], x, y);
        f[qy][qx] = pecs::testfn::initial_condition(x, y);
      }
    for (int ay = 0; ay < 2; ++ay)
      for (int ax = 0; ax < 2; ++ax) {
        double s = 0;
        for (int qy = 0; qy < 2; ++qy)
          for (int qx = 0; qx < 2; ++qx) s += binv[ax][qx] * binv[ay][qy] * f[qy][qx];
        u[8 * (size_t)S.n_cells + 4 * c + ax + 2 * ay] = s;
      }
  }
}

// ------------------------------------------------------------------------------------------- hot path forwards
void SolarCellProblem::assemble_semiconductor_rhs() {
  require_ctx(ctx, "assemble_semiconductor_rhs");
  check(pecs_assemble_semiconductor_rhs(ctx), "assemble_semiconductor_rhs");
}
void SolarCellProblem::assemble_electrolyte_rhs() {
  require_ctx(ctx, "assemble_electrolyte_rhs");
  check(pecs_assemble_electrolyte_rhs(ctx), "assemble_electrolyte_rhs");
}
void SolarCellProblem::solve_full_system() {
  require_ctx(ctx, "solve_full_system");
  check(pecs_solve_full_system(ctx), "solve_full_system");
}
void SolarCellProblem::assemble_Poisson_rhs() {
  require_ctx(ctx, "assemble_Poisson_rhs");
  check(pecs_assemble_poisson_rhs(ctx), "assemble_Poisson_rhs");
}
void SolarCellProblem::solve_Poisson() { Poisson_object.solve(); }
void SolarCellProblem::step(int n_steps) {
  require_ctx(ctx, "step");
  check(pecs_step(ctx, n_steps), "step");
}
void SolarCellProblem::set_time(double time) {
  require_ctx(ctx, "set_time");
  check(pecs_set_time(ctx, time), "set_time");
}
void SolarCellProblem::synchronize() {
  require_ctx(ctx, "synchronize");
  check(pecs_synchronize(ctx), "synchronize");
}

// ------------------------------------------------------------------------------------------- production run
void SolarCellProblem::setup_full_system_host() {
  full_system = true;
  kind = PECS_KIND_PRODUCTION;
  pecs::PhaseTimer timer("setup_full_system (host)");
  Grid_Maker::Grid grid_maker(sim_params);
  grid_maker.make_grids(semiconductor_triangulation, electrolyte_triangulation, Poisson_triangulation, full_system);
  timer.lap("make_grids");
  setup_dofs();
  timer.lap("setup_dofs");
  setup_mappings();
  timer.lap("setup_mappings");
  if (verbose) {
    Poisson_object.print_info();
    electron_hole_pair.print_info();
    redox_pair.print_info();
  }
  electron_hole_pair.penalty = 1.0; // "dont remove", reference SolarCell.cpp:1944-1945
  redox_pair.penalty = 1.0;
  assemble_Poisson_matrix();
  timer.lap("assemble_Poisson_matrix");
  delta_t = sim_params.delta_t;
  assemble_LDG_system(1.0);
  timer.lap("assemble_LDG_system");
}

void SolarCellProblem::setup_full_system() {
  // the device's one-time costs (context, kernel image, solver handles) are paid while the host builds its tables; a
  // failure here is reported by set_solvers, which does the same work itself
  const int warm_device = device;
  std::future<void> warm = std::async(std::launch::async, [warm_device] { (void)pecs_device_warmup(warm_device); });
  setup_full_system_host();
  pecs::PhaseTimer timer("setup_full_system");
  // initial values first (reference SolarCell.cpp:1980-2021): a missing or truncated restart file is reported before
  // the factorisations are paid for
  if (sim_params.restart_status) {
    electron_hole_pair.read_dofs(output_directory);
    redox_pair.read_dofs(output_directory);
  } else {
    project_initial_conditions();
  }
  timer.lap("initial conditions");
  set_solvers(); // waits for what is left of the warm-up where it first needs the device, its host preparations running
  timer.lap("set_solvers (tables + pecs_ctx_create)");
  electron_hole_pair.carrier_1.push_solution();
  electron_hole_pair.carrier_2.push_solution();
  redox_pair.carrier_1.push_solution();
  redox_pair.carrier_2.push_solution();
  // initial potential and field, reference SolarCell.cpp:2033-2034
  assemble_Poisson_rhs();
  solve_Poisson();
  timer.lap("initial states, first Poisson solve");
}

// ------------------------------------------------------------------------------------------------- output path
void SolarCellProblem::print_results(unsigned int time_step_number) {
  require_ctx(ctx, "SolarCellProblem::print_results");
  if (!output_) {
    output_.reset(new OutputState());
    output_->mesh[0].reset(new pecs::VtuMesh(semiconductor_triangulation.tables()));
    if (full_system) output_->mesh[1].reset(new pecs::VtuMesh(electrolyte_triangulation.tables()));
    output_->mesh[2].reset(new pecs::VtuMesh(Poisson_triangulation.tables()));
    for (auto& s : output_->slot)
      for (int w = 0; w < 3; ++w) {
        const int64_t n = pecs_output_doubles(ctx, w);
        if (n > 0 && output_->mesh[w]) {
          s[w] = static_cast<double*>(pecs_host_alloc((uint64_t)n * sizeof(double)));
          if (!s[w]) throw std::runtime_error("print_results: cannot allocate page-locked output buffers");
        }
      }
  }
  OutputState& out = *output_;
  const int k = out.next;
  out.next ^= 1;
  out.queue.wait_for(out.ticket[k]); // the job that wrote the files of two stamps ago has released this slot
  const PostProcessor scales(sim_params, true, "");
  double sc[4];
  scales.get_scales(sc);
  check(pecs_output_snapshot(ctx, sc, out.slot[k]), "print_results");
  // the reference runs these three as tasks of a TaskGroup (SolarCell.cpp:1829-1857); here they are one job of the
  // writer thread, which first waits for the snapshot's copies to land
  pecs_ctx* c = ctx;
  double* const* slot = out.slot[k];
  const std::string dir = output_directory;
  out.ticket[k] = out.queue.submit([this, c, slot, dir, time_step_number] {
    check(pecs_output_wait(c), "print_results (writer)");
    // three files, three tasks -- as the reference's TaskGroup; a cfg3 stamp is 110 MB of .vtu
    std::future<void> poisson = std::async(std::launch::async, [&] {
      Mixed_Assembler.output_rescaled_results(*output_->mesh[2], slot[2], sim_params, time_step_number, dir);
    });
    std::future<void> electrolyte;
    if (full_system)
      electrolyte = std::async(std::launch::async, [&] {
        LDG_Assembler.output_rescaled_results(*output_->mesh[1], redox_pair, sim_params, slot[1], time_step_number, dir);
      });
    std::exception_ptr failure;
    try {
      LDG_Assembler.output_rescaled_results(*output_->mesh[0], electron_hole_pair, sim_params, slot[0], time_step_number, dir);
    } catch (...) {
      failure = std::current_exception();
    }
    try {
      poisson.get();
    } catch (...) {
      if (!failure) failure = std::current_exception();
    }
    if (electrolyte.valid()) {
      try {
        electrolyte.get();
      } catch (...) {
        if (!failure) failure = std::current_exception();
      }
    }
    if (failure) std::rethrow_exception(failure);
  });
}

void SolarCellProblem::finish_output() {
  if (!output_) return;
  output_->queue.wait_idle();
  const std::string e = output_->queue.error();
  if (!e.empty()) throw std::runtime_error("output path: " + e);
}

void SolarCellProblem::run_full_system() {
  setup_full_system();
  const unsigned int number_outputs = sim_params.time_stamps;
  std::vector<double> timeStamps(number_outputs);
  double time;
  if (sim_params.restart_status) {
    for (unsigned int i = 0; i < number_outputs; i++)
      timeStamps[i] = sim_params.t_end + ((i + 1) * (sim_params.t_end_2 - sim_params.t_end) / number_outputs);
    time = sim_params.t_end;
  } else {
    for (unsigned int i = 0; i < number_outputs; i++) timeStamps[i] = (i + 1) * sim_params.t_end / number_outputs;
    time = 0.0;
  }
  // a restarted run continues the numbering of the first leg's files (reference SolarCell.cpp:1993: time_step_number =
  // number_outputs), so it does not overwrite Poisson-000...NNN.vtu, Semiconductor-..., Electrolyte-...
  unsigned int time_step_number = sim_params.restart_status ? number_outputs : 0;
  if (write_output) print_results(time_step_number); // the initial values, reference SolarCell.cpp:2037-2039
  time_step_number++;
  for (unsigned int k = 0; k < number_outputs; k++) {
    // same floating-point loop condition as the reference (SolarCell.cpp:2055-2080); the steps between two time
    // stamps are counted first and then replayed as one graph launch sequence
    int n = 0;
    while (time < timeStamps[k]) {
      time += delta_t;
      ++n;
    }
    step(n);
    // reference SolarCell.cpp:2082-2087; returns at once, the next steps run while the files are written
    if (write_output) print_results(time_step_number);
    time_step_number++;
  }
  synchronize();
  finish_output();
  electron_hole_pair.print_dofs(output_directory);
  redox_pair.print_dofs(output_directory);
}

// ------------------------------------------------------------------------------------------- I-V post-processing
void SolarCellProblem::interface_currents(const double* const states[4], double out[2]) {
  if (!full_system) throw std::runtime_error("interface_currents: needs the full (two-subdomain) system");
  const double* u[4];
  if (states) {
    for (int k = 0; k < 4; ++k) {
      if (!states[k]) throw std::runtime_error("interface_currents: missing state vector");
      u[k] = states[k];
    }
  } else {
    // from the device state: integrated ON the device (pecs_interface_currents), nothing is downloaded but the result
    require_ctx(ctx, "SolarCellProblem::interface_currents");
    check(pecs_interface_currents(ctx, out), "SolarCellProblem::interface_currents");
    return;
  }
  double prm[32];
  fill_params(prm);
  const pecs::MeshTables& S = semiconductor_triangulation.tables();
  const size_t ns = (size_t)S.n_cells, ne = (size_t)electrolyte_triangulation.tables().n_cells;
  double i_et = 0.0, i_ht = 0.0;
  for (size_t k = 0; k < semi_interface_cells.size(); ++k) {
    const int cs = semi_interface_cells[k], fs = semi_interface_faces[k];
    const int ce = elec_interface_cells[k], fe = elec_interface_faces[k];
    const pecs::fe::CellVerts v = verts_of(S, cs);
    double geom[4][4];
    pecs::rhsmath::boundary_geometry(v, 1.0, &geom[0][0]);
    const double* rn = u[0] + 8 * ns + 4 * (size_t)cs;
    const double* rp = u[1] + 8 * ns + 4 * (size_t)cs;
    const double* rr = u[2] + 8 * ne + 4 * (size_t)ce;
    const double* ro = u[3] + 8 * ne + 4 * (size_t)ce;
    for (int q = 0; q < 3; ++q) {
      const double t = pecs::fe::gauss_x(q), W = geom[fs][2] * pecs::fe::gauss_w(q);
      double xi, eta, N[4], Nn[4];
      pecs::fe::face_point(fs, t, xi, eta);
      pecs::fe::shape(xi, eta, N);
      pecs::fe::face_point(fe, t, xi, eta); // same quadrature index on both sides
      pecs::fe::shape(xi, eta, Nn);
      using pecs::rhsmath::trace;
      i_et += prm[PECS_P_K_ET] * (trace(N, rn) - prm[PECS_P_RHO_N_E]) * trace(Nn, ro) * W;
      i_ht += prm[PECS_P_K_HT] * (trace(N, rp) - prm[PECS_P_RHO_P_E]) * trace(Nn, rr) * W;
    }
  }
  out[0] = i_et;
  out[1] = i_ht;
}

// ------------------------------------------------------------------------------------------- manufactured tests
void SolarCellProblem::setup_test(int test_kind, unsigned int n_refine) {
  setup_test_host(test_kind, n_refine);
  set_solvers();
}

void SolarCellProblem::setup_test_host(int test_kind, unsigned int n_refine) {
  full_system = false;
  kind = test_kind;
  sim_params.set_params_for_testing(n_refine);
  Grid_Maker::Grid grid_maker(sim_params);
  const int n = (int)n_refine;
  if (kind == PECS_KIND_TEST_STEADY) {
    grid_maker.make_test_grid(Poisson_triangulation, n);
    grid_maker.make_test_grid(semiconductor_triangulation, n);
  } else if (kind == PECS_KIND_TEST_TRANSIENT) {
    grid_maker.make_test_grid(Poisson_triangulation, n);
    grid_maker.make_test_tran_grid(semiconductor_triangulation, n);
  } else if (kind == PECS_KIND_TEST_DD_POISSON) {
    grid_maker.make_DD_Poisson_grid(Poisson_triangulation, n);
    grid_maker.make_DD_Poisson_grid(semiconductor_triangulation, n);
  } else {
    throw std::runtime_error("setup_test: unknown kind");
  }
  setup_dofs();
  setup_mappings();
  electron_hole_pair.set_semiconductor_for_testing(sim_params.scaled_electron_mobility, sim_params.scaled_hole_mobility);
  if (kind == PECS_KIND_TEST_STEADY) {
    delta_t = 1.0; // unused: transient_or_steady = 0 (the reference leaves it uninitialised, SolarCell.cpp:2749-2770)
    assemble_Poisson_matrix();
    assemble_LDG_system(0.0);
  } else {
    // delta_t = h^(k+1), reference SolarCell.cpp:2886-2889, 2996-2998
    const pecs::MeshTables& S = semiconductor_triangulation.tables();
    double h = 0.0;
    for (int c = 0; c < S.n_cells; ++c) h = std::max(h, S.diameter(c));
    delta_t = 1.0;
    for (unsigned int i = 0; i < degree + 1; i++) delta_t *= h;
    assemble_Poisson_matrix();
    assemble_LDG_system(1.0);
  }
}

void SolarCellProblem::test_steady_state(const unsigned int& n_refine, ConvergenceTable& Mixed_table,
                                         ConvergenceTable& LDG_table) {
  setup_test(PECS_KIND_TEST_STEADY, n_refine);
  const pecs::MeshTables& P = Poisson_triangulation.tables();
  double h = 0.0;
  for (int c = 0; c < P.n_cells; ++c) h = std::max(h, P.diameter(c));
  assemble_Poisson_rhs();
  assemble_semiconductor_rhs();
  solve_Poisson();
  electron_hole_pair.carrier_1.solve();
  double primary_error, flux_error;
  mixed_errors(primary_error, flux_error);
  Mixed_table.add_value("h", h);
  Mixed_table.add_value("cells", P.n_cells);
  Mixed_table.add_value("dofs", Poisson_object.dofs.n_dofs());
  Mixed_table.add_value("Phi", primary_error);
  Mixed_table.add_value("D", flux_error);
  ldg_errors(0, 0.0, primary_error, flux_error);
  LDG_table.add_value("h", h);
  LDG_table.add_value("cells", semiconductor_triangulation.n_active_cells());
  LDG_table.add_value("dofs", electron_hole_pair.dofs.n_dofs());
  LDG_table.add_value("u", primary_error);
  LDG_table.add_value("J", flux_error);
}

void SolarCellProblem::test_transient(const unsigned int& n_refine, ConvergenceTable& LDG_table) {
  setup_test(PECS_KIND_TEST_TRANSIENT, n_refine);
  const pecs::MeshTables& S = semiconductor_triangulation.tables();
  double h = 0.0;
  for (int c = 0; c < S.n_cells; ++c) h = std::max(h, S.diameter(c));
  project_test_initial_condition();
  electron_hole_pair.carrier_1.push_solution();
  const double t_end = 1.0;
  double time = 0.0;
  while (time < t_end) {
    set_time(time);
    assemble_semiconductor_rhs();
    electron_hole_pair.carrier_1.solve();
    time += delta_t;
  }
  double primary_error, flux_error;
  ldg_errors(1, time, primary_error, flux_error);
  LDG_table.add_value("h", h);
  LDG_table.add_value("cells", S.n_cells);
  LDG_table.add_value("dofs", electron_hole_pair.dofs.n_dofs());
  LDG_table.add_value("u", primary_error);
  LDG_table.add_value("J", flux_error);
}

void SolarCellProblem::test_DD_Poisson(const unsigned int& n_refine, ConvergenceTable& Mixed_table,
                                       ConvergenceTable& LDG_table) {
  setup_test(PECS_KIND_TEST_DD_POISSON, n_refine);
  const pecs::MeshTables& S = semiconductor_triangulation.tables();
  double h = 0.0;
  for (int c = 0; c < S.n_cells; ++c) h = std::max(h, S.diameter(c));
  project_test_initial_condition();
  electron_hole_pair.carrier_1.push_solution();
  const double t_end = 1.0;
  double time = 0.0;
  while (time < t_end) { // order: Poisson first, then the carrier (reference SolarCell.cpp:3019-3081)
    set_time(time);
    assemble_Poisson_rhs();
    solve_Poisson();
    assemble_semiconductor_rhs();
    electron_hole_pair.carrier_1.solve();
    time += delta_t;
  }
  double primary_error, flux_error;
  ldg_errors(2, time, primary_error, flux_error);
  LDG_table.add_value("h", h);
  LDG_table.add_value("cells", S.n_cells);
  LDG_table.add_value("dofs", electron_hole_pair.dofs.n_dofs());
  LDG_table.add_value("u", primary_error);
  LDG_table.add_value("J", flux_error);
  mixed_errors(primary_error, flux_error);
  Mixed_table.add_value("h", h);
  Mixed_table.add_value("cells", Poisson_triangulation.n_active_cells());
  Mixed_table.add_value("dofs", Poisson_object.dofs.n_dofs());
  Mixed_table.add_value("Phi", primary_error);
  Mixed_table.add_value("D", flux_error);
}

// L2 errors with QIterated(QTrapez, degree+2) per direction, reference LDG.cpp:984-1133 / MixedFEM.cpp:256-295
void SolarCellProblem::ldg_errors(int which, double time, double& density_error, double& current_error) {
  electron_hole_pair.carrier_1.pull_solution();
  const pecs::MeshTables& S = semiconductor_triangulation.tables();
  const std::vector<double>& u = electron_hole_pair.carrier_1.solution;
  const size_t n = (size_t)S.n_cells;
  const double T[4] = {0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0}, W[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
  double eu = 0.0, eq = 0.0;
  for (int c = 0; c < S.n_cells; ++c) {
    const pecs::fe::CellVerts v = verts_of(S, c);
    for (int qy = 0; qy < 4; ++qy)
      for (int qx = 0; qx < 4; ++qx) {
        const pecs::fe::Jac j = pecs::fe::jacobian(v, T[qx], T[qy]);
        double N[4], x, y, uh[3] = {0, 0, 0}, ex[3];
        pecs::fe::shape(T[qx], T[qy], N);
        pecs::fe::map_point(v, T[qx], T[qy], x, y);
        for (int a = 0; a < 4; ++a)
          for (int comp = 0; comp < 3; ++comp) uh[comp] += u[comp * 4 * n + 4 * c + a] * N[a];
        if (which == 0)
          pecs::testfn::poisson_solution(x, y, ex);
        else if (which == 1)
          pecs::testfn::ldg_solution(x, y, time, ex);
        else
          pecs::testfn::dd_solution(x, y, time, ex);
        const double w = j.det * W[qx] * W[qy];
        eu += (uh[2] - ex[2]) * (uh[2] - ex[2]) * w;
        eq += ((uh[0] - ex[0]) * (uh[0] - ex[0]) + (uh[1] - ex[1]) * (uh[1] - ex[1])) * w;
      }
  }
  density_error = std::sqrt(eu);
  current_error = std::sqrt(eq);
}

void SolarCellProblem::mixed_errors(double& potential_error, double& field_error) {
  Poisson_object.pull_solution();
  const pecs::MeshTables& P = Poisson_triangulation.tables();
  const std::vector<double>& X = Poisson_object.solution;
  const double T[4] = {0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0}, W[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
  double ep = 0.0, ed = 0.0;
  for (int c = 0; c < P.n_cells; ++c) {
    const pecs::fe::CellVerts v = verts_of(P, c);
    const double phi = X[Poisson_object.dofs.phi_dof(c)];
    for (int qy = 0; qy < 4; ++qy)
      for (int qx = 0; qx < 4; ++qx) {
        const pecs::fe::Jac j = pecs::fe::jacobian(v, T[qx], T[qy]);
        double px[4], py[4], x, y, ex[3], D[2] = {0, 0};
        pecs::fe::rt0_times_det(j, T[qx], T[qy], px, py);
        pecs::fe::map_point(v, T[qx], T[qy], x, y);
        for (int f = 0; f < 4; ++f) {
          const double Xf = X[Poisson_object.dofs.face_dof[4 * c + f]] / j.det;
          D[0] += Xf * px[f];
          D[1] += Xf * py[f];
        }
        pecs::testfn::poisson_solution(x, y, ex);
        const double w = j.det * W[qx] * W[qy];
        ep += (phi - ex[2]) * (phi - ex[2]) * w;
        ed += ((D[0] - ex[0]) * (D[0] - ex[0]) + (D[1] - ex[1]) * (D[1] - ex[1])) * w;
      }
  }
  potential_error = std::sqrt(ep);
  field_error = std::sqrt(ed);
}

} // namespace SOLARCELL
