// SolverSetup.cpp -- see SolverSetup.hpp.
#include "SolverSetup.hpp"

#include <cstdlib>

#include "../error.hpp"
#include "SchurReduction.hpp"
#include "SolarCell.hpp"

namespace pecs {

NodeLayout carrier_nodes(const pecs_domain_desc& d) {
  NodeLayout L;
  const int n = d.n_cells;
  L.node_of_dof.resize(12 * (size_t)n);
  L.x.resize(n);
  L.y.resize(n);
  for (int c = 0; c < n; ++c) {
    const double* v = d.vertices + 8 * (size_t)c;
    L.x[c] = 0.25 * (v[0] + v[2] + v[4] + v[6]);
    L.y[c] = 0.25 * (v[1] + v[3] + v[5] + v[7]);
    for (int i = 0; i < 12; ++i) L.node_of_dof[(size_t)(i / 4) * 4 * n + 4 * c + (i % 4)] = c;
  }
  return L;
}

NodeLayout poisson_nodes(const pecs_poisson_desc& d) {
  NodeLayout L;
  const int n = d.n_cells;
  L.node_of_dof.assign((size_t)d.n_rt + n, -1);
  L.x.resize(n);
  L.y.resize(n);
  for (int c = 0; c < n; ++c) {
    const double* v = d.vertices + 8 * (size_t)c;
    L.x[c] = 0.25 * (v[0] + v[2] + v[4] + v[6]);
    L.y[c] = 0.25 * (v[1] + v[3] + v[5] + v[7]);
    L.node_of_dof[(size_t)d.n_rt + c] = c;
  }
  // pass 1: the cell that sees an edge as its face 1 or 3 owns it; pass 2: leftovers go to their first cell
  for (int pass = 0; pass < 2; ++pass)
    for (int c = 0; c < n; ++c)
      for (int f = 0; f < 4; ++f) {
        if (pass == 0 && !(f & 1)) continue;
        int& slot = L.node_of_dof[d.face_dof[4 * c + f]];
        if (slot < 0) slot = c;
      }
  return L;
}

NodeLayout carrier_density_nodes(const pecs_domain_desc& d) {
  NodeLayout L;
  const int n = d.n_cells;
  L.node_of_dof.resize(4 * (size_t)n);
  L.x.resize(n);
  L.y.resize(n);
  for (int c = 0; c < n; ++c) {
    const double* v = d.vertices + 8 * (size_t)c;
    L.x[c] = 0.25 * (v[0] + v[2] + v[4] + v[6]);
    L.y[c] = 0.25 * (v[1] + v[3] + v[5] + v[7]);
    for (int a = 0; a < 4; ++a) L.node_of_dof[4 * (size_t)c + a] = c;
  }
  return L;
}

bool schur_reduction_enabled() {
  const char* e = std::getenv("PECS_B200_NO_SCHUR");
  return !(e && e[0] == '1');
}

int default_leaf_nodes(bool poisson) {
  if (const char* e = std::getenv("PECS_B200_LEAF_NODES")) {
    const int v = std::atoi(e);
    if (v > 0) return v;
  }
  return poisson ? 16 : 8;
}

namespace {
struct CarrierRef {
  const MeshTables* mesh;
  const CsrMatrix* A;
};
CarrierRef carrier_ref(SOLARCELL::SolarCellProblem& s, int which) {
  if (which < 0 || which > 3) throw StatusError(PECS_ERR_INVALID, "system selector must be 0..4");
  const bool semi = which <= 1;
  const ChargeCarrierSpace::CarrierPair& pair = semi ? s.electron_hole_pair : s.redox_pair;
  return {semi ? &s.semiconductor_triangulation.tables() : &s.electrolyte_triangulation.tables(),
          (which % 2 == 0) ? &pair.carrier_1.system_matrix : &pair.carrier_2.system_matrix};
}
} // namespace

void solve_system_host(SOLARCELL::SolarCellProblem& s, int which, int leaf_nodes, const double* b, double* x) {
  std::vector<double> fwd, bwd;
  if (which == PECS_POISSON || !schur_reduction_enabled()) {
    const SolvePlan plan = plan_for_system(s, which, leaf_nodes);
    const CsrMatrix& A = which == PECS_POISSON ? s.Poisson_object.system_matrix : *carrier_ref(s, which).A;
    factorize_host(plan, A, fwd, bwd);
    solve_host(plan, fwd, bwd, b, x);
    return;
  }
  const CarrierRef ref = carrier_ref(s, which);
  const int n = ref.mesh->n_cells, nq = 8 * n, nu = 4 * n;
  SchurReduction R;
  if (!build_schur_reduction(*ref.A, n, R)) throw StatusError(PECS_ERR_INTERNAL, "carrier (q,q) block couples cells");
  pecs_domain_desc d{};
  d.n_cells = n;
  d.vertices = ref.mesh->vertices.data();
  const NodeLayout L = carrier_density_nodes(d);
  const SolvePlan plan = build_solve_plan(R.S, L.node_of_dof, L.x, L.y, leaf_nodes > 0 ? leaf_nodes : default_leaf_nodes(false));
  factorize_host(plan, R.S, fwd, bwd);
  std::vector<double> t(nu), rt(nu), q1(nq), q2(nq);
  R.T1.vmult(t.data(), b);                       // T1 r_q
  for (int i = 0; i < nu; ++i) rt[i] = b[nq + i] - t[i];
  solve_host(plan, fwd, bwd, rt.data(), x + nq); // u
  R.Ainv.vmult(q1.data(), b);
  R.T2.vmult(q2.data(), x + nq);
  for (int i = 0; i < nq; ++i) x[i] = q1[i] - q2[i];
}

SolvePlan plan_for_system(SOLARCELL::SolarCellProblem& s, int which, int leaf_nodes) {
  if (which == PECS_POISSON) {
    pecs_poisson_desc d{};
    const MeshTables& P = s.Poisson_triangulation.tables();
    d.n_cells = P.n_cells;
    d.vertices = P.vertices.data();
    d.n_rt = s.Poisson_object.dofs.n_rt;
    d.face_dof = s.Poisson_object.dofs.face_dof.data();
    const NodeLayout L = poisson_nodes(d);
    return build_solve_plan(s.Poisson_object.system_matrix, L.node_of_dof, L.x, L.y,
                            leaf_nodes > 0 ? leaf_nodes : default_leaf_nodes(true));
  }
  if (which < 0 || which > 3) throw StatusError(PECS_ERR_INVALID, "plan_for_system: which must be 0..4");
  const bool semi = which <= 1;
  const MeshTables& M = semi ? s.semiconductor_triangulation.tables() : s.electrolyte_triangulation.tables();
  pecs_domain_desc d{};
  d.n_cells = M.n_cells;
  d.vertices = M.vertices.data();
  const ChargeCarrierSpace::CarrierPair& pair = semi ? s.electron_hole_pair : s.redox_pair;
  const CsrMatrix& A = (which % 2 == 0) ? pair.carrier_1.system_matrix : pair.carrier_2.system_matrix;
  if (schur_reduction_enabled()) {
    SchurReduction R;
    if (build_schur_reduction(A, M.n_cells, R)) {
      const NodeLayout L = carrier_density_nodes(d);
      return build_solve_plan(R.S, L.node_of_dof, L.x, L.y, leaf_nodes > 0 ? leaf_nodes : default_leaf_nodes(false));
    }
  }
  const NodeLayout L = carrier_nodes(d);
  return build_solve_plan(A, L.node_of_dof, L.x, L.y, leaf_nodes > 0 ? leaf_nodes : default_leaf_nodes(false));
}

} // namespace pecs
