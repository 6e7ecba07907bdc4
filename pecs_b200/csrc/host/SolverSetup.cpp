// SolverSetup.cpp -- see SolverSetup.hpp.
#include "SolverSetup.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <future>
#include <thread>
#ifdef __linux__
#include <sched.h>
#endif

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

namespace {

struct PoissonNd {
  const std::vector<std::vector<int>>& adj; // on unknowns
  const pecs_poisson_desc& d;
  std::vector<double> cx, cy;
  int leaf_cells;
  std::vector<char> assigned;  // per unknown
  std::vector<int> side;       // per cell: 0 outside the current region, 1 / 2
  std::vector<int> comp;       // per cell: component id inside the current region (valid while stamp matches)
  std::vector<int> comp_stamp; // per cell
  int stamp = 0;
  EliminationTree out;

  int phi(int cell) const { return d.n_rt + cell; }

  // connected components of the region's cells through edge fluxes no ancestor has taken; returns their number
  int components(const std::vector<int>& cells) {
    ++stamp;
    for (int c : cells) comp_stamp[c] = -stamp; // in region, not visited
    int n_comp = 0;
    std::vector<int> stack;
    for (int c0 : cells) {
      if (comp_stamp[c0] == stamp) continue;
      comp_stamp[c0] = stamp;
      comp[c0] = n_comp;
      stack.push_back(c0);
      while (!stack.empty()) {
        const int c = stack.back();
        stack.pop_back();
        for (int e : adj[phi(c)]) {
          if (e >= d.n_rt || assigned[e]) continue;
          for (int w : adj[e]) {
            if (w < d.n_rt) continue;
            const int c2 = w - d.n_rt;
            if (comp_stamp[c2] == -stamp) {
              comp_stamp[c2] = stamp;
              comp[c2] = n_comp;
              stack.push_back(c2);
            }
          }
        }
      }
      ++n_comp;
    }
    return n_comp;
  }

  // returns the tree index; `delayed` receives the potentials (as cells) this region passes up, one per component
  int build(std::vector<int>& cells, std::vector<int>& delayed) {
    if ((int)cells.size() <= leaf_cells) {
      const int n_comp = components(cells);
      std::vector<char> has(n_comp, 0);
      EliminationTree::Node t;
      for (int c : cells) {
        if (!has[comp[c]]) {
          has[comp[c]] = 1;
          delayed.push_back(c);
        } else {
          t.nodes.push_back(phi(c));
        }
        for (int e : adj[phi(c)])
          if (e < d.n_rt && !assigned[e]) {
            assigned[e] = 1;
            t.nodes.push_back(e);
          }
        for (int f = 0; f < 4; ++f) { // decoupled rows (hanging children, Neumann edges) go with their first cell
          const int e = d.face_dof[4 * (size_t)c + f];
          if (!assigned[e] && adj[e].empty()) {
            assigned[e] = 1;
            t.nodes.push_back(e);
          }
        }
      }
      std::sort(t.nodes.begin(), t.nodes.end());
      out.tree.push_back(std::move(t));
      return (int)out.tree.size() - 1;
    }
    std::vector<int> bestS, bestA, bestB;
    for (int dir = 0; dir < 2; ++dir) {
      const std::vector<double>& c = dir == 0 ? cx : cy;
      std::vector<int> sorted = cells;
      const size_t half = sorted.size() / 2;
      std::nth_element(sorted.begin(), sorted.begin() + half, sorted.end(),
                       [&](int a, int b) { return c[a] < c[b] || (c[a] == c[b] && a < b); });
      for (size_t k = 0; k < sorted.size(); ++k) side[sorted[k]] = k < half ? 1 : 2;
      std::vector<int> S;
      for (size_t k = 0; k < half; ++k)
        for (int e : adj[phi(sorted[k])]) {
          if (e >= d.n_rt || assigned[e]) continue;
          bool other = false;
          for (int w : adj[e])
            if (w >= d.n_rt && side[w - d.n_rt] == 2) other = true;
          if (other) S.push_back(e);
        }
      std::sort(S.begin(), S.end());
      S.erase(std::unique(S.begin(), S.end()), S.end());
      for (int v : sorted) side[v] = 0;
      if (dir == 0 || S.size() < bestS.size()) {
        bestS.swap(S);
        bestA.assign(sorted.begin(), sorted.begin() + half);
        bestB.assign(sorted.begin() + half, sorted.end());
      }
    }
    // components of this region are defined BEFORE its own separator is taken out
    const int n_comp = components(cells);
    std::vector<int> comp_of_cell_local;
    std::vector<int> region_cells = cells; // keep for the component lookup of the delayed potentials
    std::vector<int> region_comp(region_cells.size());
    for (size_t k = 0; k < region_cells.size(); ++k) region_comp[k] = comp[region_cells[k]];
    for (int e : bestS) assigned[e] = 1;
    cells.clear();
    cells.shrink_to_fit();
    EliminationTree::Node t;
    std::vector<int> from_children;
    t.child[0] = build(bestA, from_children);
    t.child[1] = build(bestB, from_children);
    // the recursion overwrote comp[]: restore this region's numbering
    for (size_t k = 0; k < region_cells.size(); ++k) comp[region_cells[k]] = region_comp[k];
    std::vector<char> has(n_comp, 0);
    t.nodes = bestS;
    for (int c : from_children) {
      if (!has[comp[c]]) {
        has[comp[c]] = 1;
        delayed.push_back(c);
      } else {
        t.nodes.push_back(phi(c));
      }
    }
    std::sort(t.nodes.begin(), t.nodes.end());
    out.tree.push_back(std::move(t));
    return (int)out.tree.size() - 1;
  }
};

} // namespace

SolvePlan poisson_plan(const CsrMatrix& A, const pecs_poisson_desc& d, int leaf_cells) {
  if (std::getenv("PECS_B200_POISSON_CELL_NODES") != nullptr) return plan_from_layout(A, poisson_nodes(d), leaf_cells);
  const int n = d.n_rt + d.n_cells;
  if (A.n != n) throw StatusError(PECS_ERR_INVALID, "poisson_plan: matrix size");
  std::vector<int> identity(n);
  for (int i = 0; i < n; ++i) identity[i] = i;
  const int threads = preparation_threads();
  const std::vector<std::vector<int>> adj = node_adjacency(A, identity, n, threads);
  PoissonNd nd{adj, d, {}, {}, std::max(1, leaf_cells), std::vector<char>(n, 0), std::vector<int>(d.n_cells, 0),
               std::vector<int>(d.n_cells, 0), std::vector<int>(d.n_cells, 0), 0, {}};
  nd.cx.resize(d.n_cells);
  nd.cy.resize(d.n_cells);
  for (int c = 0; c < d.n_cells; ++c) {
    const double* v = d.vertices + 8 * (size_t)c;
    nd.cx[c] = 0.25 * (v[0] + v[2] + v[4] + v[6]);
    nd.cy[c] = 0.25 * (v[1] + v[3] + v[5] + v[7]);
  }
  std::vector<int> all(d.n_cells), rest;
  for (int c = 0; c < d.n_cells; ++c) all[c] = c;
  nd.build(all, rest);
  // the root eliminates the potentials that are still delayed, and anything the tables never mentioned
  EliminationTree::Node& root = nd.out.tree.back();
  for (int c : rest) root.nodes.push_back(nd.phi(c));
  for (int e = 0; e < d.n_rt; ++e)
    if (!nd.assigned[e]) root.nodes.push_back(e);
  std::sort(root.nodes.begin(), root.nodes.end());
  return build_solve_plan(A, identity, n, adj, nd.out, threads);
}

NodeLayout carrier_density_nodes(const pecs_domain_desc& d) {
  NodeLayout L;
  const int n = d.n_cells;
  L.node_of_dof.resize(4 * (size_t)n);
  L.group_of_node.resize(4 * (size_t)n);
  L.x.resize(n);
  L.y.resize(n);
  const bool per_cell = std::getenv("PECS_B200_CELL_SEPARATORS") != nullptr; // comparison: whole cells as graph nodes
  for (int c = 0; c < n; ++c) {
    const double* v = d.vertices + 8 * (size_t)c;
    L.x[c] = 0.25 * (v[0] + v[2] + v[4] + v[6]);
    L.y[c] = 0.25 * (v[1] + v[3] + v[5] + v[7]);
    for (int a = 0; a < 4; ++a) {
      L.node_of_dof[4 * (size_t)c + a] = per_cell ? c : 4 * c + a;
      L.group_of_node[4 * (size_t)c + a] = c;
    }
  }
  if (per_cell) L.group_of_node.clear();
  return L;
}

SolvePlan plan_from_layout(const CsrMatrix& A, const NodeLayout& L, int leaf_groups, int threads) {
  return build_solve_plan(A, L.node_of_dof, L.group_of_node, L.x, L.y, leaf_groups, threads);
}

bool schur_reduction_enabled() {
  const char* e = std::getenv("PECS_B200_NO_SCHUR");
  return !(e && e[0] == '1');
}

int default_leaf_nodes(bool poisson) {
  if (poisson)
    if (const char* e = std::getenv("PECS_B200_POISSON_LEAF_NODES")) {
      const int v = std::atoi(e);
      if (v > 0) return v;
    }
  if (const char* e = std::getenv("PECS_B200_LEAF_NODES")) {
    const int v = std::atoi(e);
    if (v > 0) return v;
  }
  return 16;
}

namespace {
double wall_seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace

PhaseTimer::PhaseTimer(const char* scope_) : on(std::getenv("PECS_B200_SETUP_TIMING") != nullptr), scope(scope_), t(wall_seconds()) {}
void PhaseTimer::lap(const char* what) {
  if (!on) return;
  const double now = wall_seconds();
  std::fprintf(stderr, "%s: %-44s %.2f s\n", scope, what, now - t);
  t = now;
}

CsrMatrix copy_csr(const pecs_csr& a, int expected_n, const char* what) {
  if (!(a.n == expected_n && a.row_ptr && a.col && a.val)) throw StatusError(PECS_ERR_INVALID, what);
  CsrMatrix A;
  A.n = a.n;
  A.row_ptr.assign(a.row_ptr, a.row_ptr + a.n + 1);
  const size_t nnz = (size_t)A.row_ptr[a.n];
  A.col.assign(a.col, a.col + nnz);
  A.val.assign(a.val, a.val + nnz);
  return A;
}

int preparation_threads() {
  if (const char* e = std::getenv("PECS_B200_SETUP_THREADS"))
    if (std::atoi(e) > 0) return std::atoi(e);
  int hw = (int)std::thread::hardware_concurrency();
#ifdef __linux__
  cpu_set_t allowed; // a rank of a multi-GPU job is pinned to its share of the cores (sweep.pin_to_gpu_numa_node)
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) hw = std::max(1, std::min(hw, CPU_COUNT(&allowed)));
#endif
  return std::max(1, std::min(8, hw / 4));
}

namespace {
// the two permuted copies are independent of each other: the transposed one on a second thread
void permuted_copies_of(PreparedSystem& ps) {
  const int threads = std::max(1, preparation_threads() / 2);
  std::future<CsrMatrix> transposed =
      std::async(std::launch::async, [&ps, threads] { return permute_csr(ps.A, ps.plan.perm, true, threads); });
  ps.Ap = permute_csr(ps.A, ps.plan.perm, false, threads);
  ps.Apt = transposed.get();
}
} // namespace

namespace {
void ell_tables_of(PreparedSystem& ps) {
  const int threads = preparation_threads();
  ps.ell_A = build_ell(ps.A, &ps.plan.iperm, threads);
  if (ps.reduced) {
    ps.ell_T1 = build_ell(ps.R.T1, &ps.plan.iperm, threads); // r~ is produced directly in elimination order
    ps.ell_Ainv = build_ell(ps.R.Ainv, nullptr, threads);
    ps.ell_T2 = build_ell(ps.R.T2, nullptr, threads);
  }
}
} // namespace

PreparedSystem prepare_carrier(const pecs_domain_desc& d, int k, bool for_device) {
  PhaseTimer timer(k == 0 ? "prepare carrier_1" : "prepare carrier_2");
  PreparedSystem ps;
  ps.present = true;
  const int n = d.n_cells;
  const pecs_csr& a = d.system_matrix[k];
  if (!(a.n == 12 * n && a.row_ptr && a.col && a.val)) throw StatusError(PECS_ERR_INVALID, "domain: system matrix size");
  if (schur_reduction_enabled() &&
      build_schur_reduction(CsrView(a.n, a.row_ptr, a.col, a.val), n, ps.R, preparation_threads())) {
    timer.lap("Schur reduction of the currents");
    ps.reduced = true;
    ps.A = std::move(ps.R.S); // R keeps T1, Ainv, T2
    ps.plan = plan_from_layout(ps.A, carrier_density_nodes(d), default_leaf_nodes(false), preparation_threads());
  } else {
    ps.A = copy_csr(a, 12 * n, "domain: system matrix size");
    ps.plan = plan_from_layout(ps.A, carrier_nodes(d), default_leaf_nodes(false), preparation_threads());
  }
  timer.lap("dissection and symbolic plan");
  if (for_device) permuted_copies_of(ps);
  timer.lap("matrix in elimination order (+ transpose)");
  if (for_device) ell_tables_of(ps);
  timer.lap("ELL tables");
  return ps;
}

PreparedSystem prepare_poisson(const pecs_poisson_desc& P, int n_dofs, bool for_device) {
  PhaseTimer timer("prepare Poisson");
  PreparedSystem ps;
  ps.present = true;
  ps.A = copy_csr(P.system_matrix, n_dofs, "poisson: system matrix size");
  ps.plan = poisson_plan(ps.A, P, default_leaf_nodes(true));
  timer.lap("dissection and symbolic plan");
  if (for_device) permuted_copies_of(ps);
  timer.lap("matrix in elimination order (+ transpose)");
  if (for_device) ell_tables_of(ps);
  timer.lap("ELL tables");
  return ps;
}

SolvePlan plan_for_system(SOLARCELL::SolarCellProblem& s, int which, int leaf_nodes) {
  if (which == PECS_POISSON) {
    pecs_poisson_desc d{};
    const MeshTables& P = s.Poisson_triangulation.tables();
    d.n_cells = P.n_cells;
    d.vertices = P.vertices.data();
    d.n_rt = s.Poisson_object.dofs.n_rt;
    d.face_dof = s.Poisson_object.dofs.face_dof.data();
    return poisson_plan(s.Poisson_object.system_matrix, d, leaf_nodes > 0 ? leaf_nodes : default_leaf_nodes(true));
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
      return plan_from_layout(R.S, L, leaf_nodes > 0 ? leaf_nodes : default_leaf_nodes(false));
    }
  }
  const NodeLayout L = carrier_nodes(d);
  return plan_from_layout(A, L, leaf_nodes > 0 ? leaf_nodes : default_leaf_nodes(false));
}

} // namespace pecs
