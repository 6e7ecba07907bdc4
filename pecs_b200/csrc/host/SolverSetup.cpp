// SolverSetup.cpp -- see SolverSetup.hpp.
#include "SolverSetup.hpp"

#include "../error.hpp"
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

int default_leaf_nodes(bool poisson) { return poisson ? 16 : 4; }

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
  const NodeLayout L = carrier_nodes(d);
  const ChargeCarrierSpace::CarrierPair& pair = semi ? s.electron_hole_pair : s.redox_pair;
  const CsrMatrix& A = (which % 2 == 0) ? pair.carrier_1.system_matrix : pair.carrier_2.system_matrix;
  return build_solve_plan(A, L.node_of_dof, L.x, L.y, leaf_nodes > 0 ? leaf_nodes : default_leaf_nodes(false));
}

} // namespace pecs
