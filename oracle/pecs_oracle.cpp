// oracle/pecs_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle). See pecs_oracle.hpp.
#include "pecs_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <stdexcept>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

// ------------------------------------------------------------------------------------------------ mesh helpers
double Mesh::diameter(int c) const { // cell->diameter(): longest diagonal
  const double* p = v(c);
  return std::max(std::hypot(p[6] - p[0], p[7] - p[1]), std::hypot(p[4] - p[2], p[5] - p[3]));
}
void Mesh::center(int c, double& x, double& y) const {
  const double* p = v(c);
  x = 0.25 * (p[0] + p[2] + p[4] + p[6]);
  y = 0.25 * (p[1] + p[3] + p[5] + p[7]);
}
void Mesh::face_center(int c, int f, double& x, double& y) const {
  static const int fv[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
  const double* p = v(c);
  x = 0.5 * (p[2 * fv[f][0]] + p[2 * fv[f][1]]);
  y = 0.5 * (p[2 * fv[f][0] + 1] + p[2 * fv[f][1] + 1]);
}

// ------------------------------------------------------------------------------------------------ sparse matrix
void SparseMatrix::add(int i, int j, double v) {
  auto& r = rows[i];
  for (auto& e : r)
    if (e.first == j) {
      e.second += v;
      return;
    }
  r.emplace_back(j, v);
}
void SparseMatrix::vmult(std::vector<double>& y, const std::vector<double>& x) const {
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (const auto& e : rows[i]) s += e.second * x[e.first];
    y[i] = s;
  }
}
size_t SparseMatrix::nnz() const {
  size_t s = 0;
  for (const auto& r : rows) s += r.size();
  return s;
}
void SparseMatrix::to_csr(std::vector<int>& rp, std::vector<int>& col, std::vector<double>& val) const {
  rp.assign(n + 1, 0);
  col.clear();
  val.clear();
  for (int i = 0; i < n; ++i) {
    std::vector<std::pair<int, double>> r = rows[i];
    std::sort(r.begin(), r.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    for (const auto& e : r) {
      col.push_back(e.first);
      val.push_back(e.second);
    }
    rp[i + 1] = (int)col.size();
  }
}
CscMatrix SparseMatrix::to_csc() const {
  CscMatrix A;
  A.n = n;
  A.colptr.assign(n + 1, 0);
  for (const auto& r : rows)
    for (const auto& e : r) ++A.colptr[e.first + 1];
  std::partial_sum(A.colptr.begin(), A.colptr.end(), A.colptr.begin());
  A.rowind.resize(A.colptr[n]);
  A.val.resize(A.colptr[n]);
  std::vector<int> pos(A.colptr.begin(), A.colptr.end() - 1);
  for (int i = 0; i < n; ++i)
    for (const auto& e : rows[i]) {
      const int p = pos[e.first]++;
      A.rowind[p] = i;
      A.val[p] = e.second;
    }
  return A;
}

// ------------------------------------------------------------------------------------------------ constraints
void Constraints::distribute_local_to_global(const std::vector<double>& local, const std::vector<int>& dofs,
                                             std::vector<double>& global) const {
  for (size_t i = 0; i < dofs.size(); ++i) {
    auto it = lines.find(dofs[i]);
    if (it == lines.end())
      global[dofs[i]] += local[i];
    else if (it->second.master >= 0)
      global[it->second.master] += it->second.weight * local[i];
  }
}
void Constraints::distribute_local_to_global(const std::vector<std::vector<double>>& local, const std::vector<int>& dofs,
                                             SparseMatrix& global) const {
  const size_t n = dofs.size();
  for (size_t i = 0; i < n; ++i) {
    auto ci = lines.find(dofs[i]);
    const bool i_con = ci != lines.end();
    for (size_t j = 0; j < n; ++j) {
      if (local[i][j] == 0.0) continue;
      auto cj = lines.find(dofs[j]);
      const bool j_con = cj != lines.end();
      int gi = dofs[i], gj = dofs[j];
      double w = 1.0;
      if (i_con) { gi = ci->second.master; w *= ci->second.weight; }
      if (j_con) { gj = cj->second.master; w *= cj->second.weight; }
      if (gi >= 0 && gj >= 0) global.add(gi, gj, w * local[i][j]);
    }
    // constrained rows keep a non-zero diagonal so that the matrix stays regular
    if (i_con) global.add(dofs[i], dofs[i], local[i][i] != 0.0 ? local[i][i] : 1.0);
  }
}
void Constraints::distribute(std::vector<double>& x) const {
  for (const auto& l : lines) x[l.first] = l.second.master >= 0 ? l.second.weight * x[l.second.master] : 0.0;
}

// ------------------------------------------------------------------------------------------------ ordering
// Fill-reducing elimination order by geometric nested dissection on the matrix graph (stands in for UMFPACK's
// COLAMD/AMD pre-ordering; any order gives the same solution up to round-off).
namespace {
struct NdContext {
  const std::vector<std::vector<int>>* adj;
  const std::vector<double>*cx, *cy;
  std::vector<int> part; // scratch: 0 none, 1 = A, 2 = B
  std::vector<int> order;
};

void nd_recurse(NdContext& ctx, std::vector<int>& nodes) {
  if (nodes.size() <= 48) {
    for (int v : nodes) ctx.order.push_back(v);
    return;
  }
  std::vector<int> bestA, bestB, bestS;
  for (int dir = 0; dir < 2; ++dir) {
    const std::vector<double>& coord = dir == 0 ? *ctx.cx : *ctx.cy;
    std::vector<int> sorted = nodes;
    const size_t half = sorted.size() / 2;
    std::nth_element(sorted.begin(), sorted.begin() + half, sorted.end(), [&](int a, int b) {
      return coord[a] < coord[b] || (coord[a] == coord[b] && a < b);
    });
    for (size_t k = 0; k < sorted.size(); ++k) ctx.part[sorted[k]] = k < half ? 1 : 2;
    std::vector<int> A, B, S;
    for (size_t k = 0; k < sorted.size(); ++k) {
      const int v = sorted[k];
      if (k >= half) {
        B.push_back(v);
        continue;
      }
      bool touches = false;
      for (int w : (*ctx.adj)[v])
        if (ctx.part[w] == 2) {
          touches = true;
          break;
        }
      (touches ? S : A).push_back(v);
    }
    for (int v : sorted) ctx.part[v] = 0;
    if (dir == 0 || S.size() < bestS.size()) {
      bestA.swap(A);
      bestB.swap(B);
      bestS.swap(S);
    }
  }
  if (bestA.empty() || bestB.empty()) { // degenerate split: stop here
    for (int v : nodes) ctx.order.push_back(v);
    return;
  }
  nodes.clear();
  nodes.shrink_to_fit();
  nd_recurse(ctx, bestA);
  nd_recurse(ctx, bestB);
  for (int v : bestS) ctx.order.push_back(v);
}

std::vector<int> nested_dissection_order(const SparseMatrix& A, const std::vector<double>& cx,
                                         const std::vector<double>& cy) {
  std::vector<std::vector<int>> adj(A.n);
  for (int i = 0; i < A.n; ++i)
    for (const auto& e : A.rows[i])
      if (e.first != i) {
        adj[i].push_back(e.first);
        adj[e.first].push_back(i);
      }
  for (auto& a : adj) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  NdContext ctx;
  ctx.adj = &adj;
  ctx.cx = &cx;
  ctx.cy = &cy;
  ctx.part.assign(A.n, 0);
  std::vector<int> all(A.n);
  std::iota(all.begin(), all.end(), 0);
  nd_recurse(ctx, all);
  return ctx.order;
}
} // namespace

// ------------------------------------------------------------------------------------------------ dof setup
void CarrierPair::setup_dofs(const Mesh& mesh) {
  n_cells = mesh.n_cells;
  const int n = 12 * n_cells;
  for (Carrier* c : {&carrier_1, &carrier_2}) {
    c->system_matrix.reinit(n);
    c->system_rhs.assign(n, 0.0);
    c->solution.assign(n, 0.0);
  }
  mass_matrix.reinit(n);
}

void PoissonData::setup_dofs(const Mesh& mesh) {
  // distribute_dofs: cell by cell, the four lines then the cell interior; component_wise keeps the relative
  // order of the flux dofs and moves the potentials behind them (reference Poisson.cpp:25-28)
  n_cells = mesh.n_cells;
  face_dof.assign(4 * (size_t)n_cells, -1);
  n_rt = 0;
  for (int c = 0; c < n_cells; ++c)
    for (int f = 0; f < 4; ++f) {
      if (face_dof[4 * c + f] >= 0) continue;
      face_dof[4 * c + f] = n_rt;
      if (mesh.face_kind[4 * c + f] == FACE_SAME_LEVEL) face_dof[4 * mesh.neighbor[4 * c + f] + (f ^ 1)] = n_rt;
      ++n_rt;
    }
  constraints.lines.clear();
  for (int c = 0; c < n_cells; ++c)
    for (int f = 0; f < 4; ++f) {
      const int kind = mesh.face_kind[4 * c + f];
      if (kind == FACE_COARSER) // make_hanging_node_constraints: child edge = 1/2 parent edge
        constraints.lines[face_dof[4 * c + f]] = {face_dof[4 * mesh.neighbor[4 * c + f] + (f ^ 1)], 0.5};
      else if (kind == FACE_BOUNDARY && mesh.boundary_id[4 * c + f] == Neumann) // make_zero_boundary_constraints
        constraints.lines[face_dof[4 * c + f]] = {-1, 0.0};
    }
  const int n = n_rt + n_cells;
  system_matrix.reinit(n);
  system_rhs.assign(n, 0.0);
  solution.assign(n, 0.0);
}

void SolarCellProblem::setup_dofs() { // reference SolarCell.cpp:105-121
  Poisson_object.setup_dofs(Poisson_mesh);
  electron_hole_pair.setup_dofs(semiconductor_mesh);
  if (full_system) redox_pair.setup_dofs(electrolyte_mesh);
}

// ------------------------------------------------------------------------------------------------ mappings
void SolarCellProblem::setup_mappings() {
  // reference SolarCell.cpp:156-372 matches centres with |d| < 1e-13 in O(N^2); the same pairs are found here
  // through an ordered map on the (identically computed) centre coordinates.
  typedef std::pair<double, double> Key;
  auto center_map = [](const Mesh& m) {
    std::map<Key, int> mp;
    for (int c = 0; c < m.n_cells; ++c) {
      double x, y;
      m.center(c, x, y);
      mp[Key(x, y)] = c;
    }
    return mp;
  };
  const std::map<Key, int> pmap = center_map(Poisson_mesh);
  auto match = [&](const Mesh& m, std::map<int, int>& out) {
    out.clear();
    for (int c = 0; c < m.n_cells; ++c) {
      double x, y;
      m.center(c, x, y);
      auto it = pmap.find(Key(x, y));
      if (it == pmap.end()) throw std::runtime_error("setup_mappings: carrier cell without Poisson cell");
      out[c] = it->second;
    }
  };
  match(semiconductor_mesh, s_2_p_map);
  if (!full_system) return;
  match(electrolyte_mesh, e_2_p_map);

  semi_interface_cells.clear();
  semi_interface_faces.clear();
  std::map<Key, std::pair<int, int>> elec_faces;
  for (int c = 0; c < electrolyte_mesh.n_cells; ++c)
    for (int f = 0; f < 4; ++f)
      if (electrolyte_mesh.face_kind[4 * c + f] == FACE_BOUNDARY && electrolyte_mesh.boundary_id[4 * c + f] == Interface) {
        double x, y;
        electrolyte_mesh.face_center(c, f, x, y);
        elec_faces[Key(x, y)] = std::make_pair(c, f);
      }
  for (int c = 0; c < semiconductor_mesh.n_cells; ++c)
    for (int f = 0; f < 4; ++f)
      if (semiconductor_mesh.face_kind[4 * c + f] == FACE_BOUNDARY &&
          semiconductor_mesh.boundary_id[4 * c + f] == Interface) {
        double x, y;
        semiconductor_mesh.face_center(c, f, x, y);
        auto it = elec_faces.find(Key(x, y));
        if (it == elec_faces.end()) throw std::runtime_error("setup_mappings: unmatched interface face");
        semi_interface_cells.push_back(c);
        semi_interface_faces.push_back(f);
        elec_interface_cells.push_back(it->second.first);
        elec_interface_faces.push_back(it->second.second);
      }
  semi_interface_map.clear();
  elec_interface_map.clear();
  for (size_t i = 0; i < semi_interface_cells.size(); ++i) {
    semi_interface_map[semi_interface_cells[i]] = (int)i;
    elec_interface_map[elec_interface_cells[i]] = (int)i;
  }
}

// ------------------------------------------------------------------------------------------------ Poisson matrix
void SolarCellProblem::assemble_Poisson_matrix() {
  // reference MixedFEM.cpp:29-164 (local) + SolarCell.cpp:408-417 (constrained scatter)
  FEValues fe;
  std::vector<int> dofs(5);
  std::vector<std::vector<double>> local(5, std::vector<double>(5));
  for (int cell = 0; cell < Poisson_mesh.n_cells; ++cell) {
    fe.reinit(Poisson_mesh.v(cell));
    Poisson_object.get_dof_indices(cell, dofs);
    for (auto& r : local) std::fill(r.begin(), r.end(), 0.0);
    const int mat = Poisson_mesh.material[cell];
    double permittivity;
    if (mat == 0 || mat == 1)
      permittivity = prm[P_EPS_S];
    else if (mat == 2 || mat == 3)
      permittivity = prm[P_EPS_E];
    else
      throw std::runtime_error("CELL TYPE NOT SEMICONDUCTOR OR ELECTROLYTE");
    for (int q = 0; q < 9; ++q)
      for (int i = 0; i < 5; ++i) {
        const Tensor1 psi_i_field = fe.field_value(i, q);
        const double div_psi_i_field = fe.field_divergence(i, q);
        const double psi_i_potential = fe.potential_value(i, q);
        for (int j = 0; j < 5; ++j) {
          const Tensor1 psi_j_field = fe.field_value(j, q);
          const double div_psi_j_field = fe.field_divergence(j, q);
          const double psi_j_potential = fe.potential_value(j, q);
          local[i][j] += ((psi_i_field * psi_j_field) * (1.0 / permittivity) - div_psi_i_field * psi_j_potential -
                          psi_i_potential * prm[P_LAMBDA2] * div_psi_j_field) *
                         fe.JxW(q);
        }
      }
    Poisson_object.constraints.distribute_local_to_global(local, dofs, Poisson_object.system_matrix);
  }
}

// ------------------------------------------------------------------------------------------------ LDG matrices
void SolarCellProblem::assemble_local_LDG(CarrierPair& pair, const Mesh& mesh, double transient_or_steady) {
  // mass matrix: reference LDG.cpp:40-83 ; cell + boundary terms: LDG.cpp:85-281
  FEValues fe;
  FEFaceValues ffe;
  std::vector<int> dofs(12);
  const double mu[2] = {pair.carrier_1.scaled_mobility, pair.carrier_2.scaled_mobility};
  SparseMatrix* mats[2] = {&pair.carrier_1.system_matrix, &pair.carrier_2.system_matrix};
  for (int cell = 0; cell < mesh.n_cells; ++cell) {
    fe.reinit(mesh.v(cell));
    pair.get_dof_indices(cell, dofs);
    const double h = mesh.diameter(cell);
    double local_mass[12][12] = {}, local_1[12][12] = {}, local_2[12][12] = {};
    for (int q = 0; q < 9; ++q)
      for (int i = 0; i < 12; ++i) {
        const Tensor1 psi_i_field = fe.current_value(i, q);
        const double div_psi_i_field = fe.current_divergence(i, q);
        const Tensor1 grad_psi_i_density = fe.density_gradient(i, q);
        const double psi_i_density = fe.density_value(i, q);
        for (int j = 0; j < 12; ++j) {
          const Tensor1 psi_j_field = fe.current_value(j, q);
          const double psi_j_density = fe.density_value(j, q);
          local_mass[i][j] += (1.0 / delta_t) * psi_i_density * psi_j_density * fe.JxW(q);
          const double common = (transient_or_steady / delta_t) * psi_i_density * psi_j_density -
                                div_psi_i_field * psi_j_density - grad_psi_i_density * psi_j_field;
          local_1[i][j] += (common + (psi_i_field * psi_j_field) * (1.0 / mu[0])) * fe.JxW(q);
          local_2[i][j] += (common + (psi_i_field * psi_j_field) * (1.0 / mu[1])) * fe.JxW(q);
        }
      }
    for (int face_no = 0; face_no < 4; ++face_no) {
      if (mesh.face_kind[4 * cell + face_no] != FACE_BOUNDARY) continue;
      ffe.reinit(mesh.v(cell), face_no);
      const int bid = mesh.boundary_id[4 * cell + face_no];
      if (bid == Dirichlet) {
        for (int q = 0; q < 3; ++q)
          for (int i = 0; i < 12; ++i) {
            const double psi_i_density = ffe.density_value(i, q);
            for (int j = 0; j < 12; ++j) {
              const Tensor1 psi_j_field = ffe.current_value(j, q);
              const double psi_j_density = ffe.density_value(j, q);
              const double v = psi_i_density * (ffe.normal_vector(q) * psi_j_field + (pair.penalty / h) * psi_j_density) *
                               ffe.JxW(q);
              local_1[i][j] += v;
              local_2[i][j] += v;
            }
          }
      } else if (bid == Interface || bid == Neumann || bid == Schottky) {
        for (int q = 0; q < 3; ++q)
          for (int i = 0; i < 12; ++i) {
            const Tensor1 psi_i_field = ffe.current_value(i, q);
            for (int j = 0; j < 12; ++j) {
              const double v = (psi_i_field * ffe.normal_vector(q)) * ffe.density_value(j, q) * ffe.JxW(q);
              local_1[i][j] += v;
              local_2[i][j] += v;
            }
          }
      } else {
        throw std::runtime_error("LDG: no other boundary terms");
      }
    }
    for (int i = 0; i < 12; ++i)
      for (int j = 0; j < 12; ++j) {
        if (local_mass[i][j] != 0.0) pair.mass_matrix.add(dofs[i], dofs[j], local_mass[i][j]);
        if (local_1[i][j] != 0.0) mats[0]->add(dofs[i], dofs[j], local_1[i][j]);
        if (local_2[i][j] != 0.0) mats[1]->add(dofs[i], dofs[j], local_2[i][j]);
      }
  }
}

void SolarCellProblem::assemble_flux_terms(CarrierPair& pair, const Mesh& mesh) {
  // sequential face loop: reference LDG.cpp:283-426; local flux matrices: LDG.cpp:429-622;
  // scatter into both carriers: LDG.cpp:624-678.
  // Deviation: on a face with children the reference re-initialises the SUB-face evaluator but
  // then reads the (stale) face evaluator.  The intended sub-face integral is implemented here.
  FEFaceValues face_values, neighbor_face_values;
  FESubfaceValues subface_values;
  std::vector<int> dofs(12), ndofs(12);
  Tensor1 beta{{1.0, 1.0}};
  const double bn = std::sqrt(beta * beta);
  beta.c[0] /= bn;
  beta.c[1] /= bn;
  SparseMatrix* mats[2] = {&pair.carrier_1.system_matrix, &pair.carrier_2.system_matrix};

  auto local_flux = [&](const FEValuesBase& minus, const FEValuesBase& plus, double penalty) {
    double vi_ui[12][12] = {}, vi_ue[12][12] = {}, ve_ui[12][12] = {}, ve_ue[12][12] = {};
    for (int q = 0; q < 3; ++q) {
      const Tensor1 n = minus.normal_vector(q);
      const double JxW = minus.JxW(q);
      for (int i = 0; i < 12; ++i) {
        const Tensor1 pi_m = minus.current_value(i, q), pi_p = plus.current_value(i, q);
        const double vi_m = minus.density_value(i, q), vi_p = plus.density_value(i, q);
        for (int j = 0; j < 12; ++j) {
          const Tensor1 pj_m = minus.current_value(j, q), pj_p = plus.current_value(j, q);
          const double vj_m = minus.density_value(j, q), vj_p = plus.density_value(j, q);
          vi_ui[i][j] += (0.5 * ((pi_m * n) * vj_m + vi_m * (n * pj_m)) + (beta * pi_m) * vj_m - (beta * pj_m) * vi_m +
                          penalty * vi_m * vj_m) * JxW;
          vi_ue[i][j] += (0.5 * ((pi_m * n) * vj_p + vi_m * (n * pj_p)) - (beta * pi_m) * vj_p + (beta * pj_p) * vi_m -
                          penalty * vi_m * vj_p) * JxW;
          ve_ui[i][j] += (-0.5 * ((pi_p * n) * vj_m + vi_p * (n * pj_m)) - (beta * pi_p) * vj_m + (beta * pj_m) * vi_p -
                          penalty * vi_p * vj_m) * JxW;
          ve_ue[i][j] += (-0.5 * ((pi_p * n) * vj_p + vi_p * (n * pj_p)) + (beta * pi_p) * vj_p - (beta * pj_p) * vi_p +
                          penalty * vi_p * vj_p) * JxW;
        }
      }
    }
    for (SparseMatrix* M : mats)
      for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) {
          if (vi_ui[i][j] != 0.0) M->add(dofs[i], dofs[j], vi_ui[i][j]);
          if (vi_ue[i][j] != 0.0) M->add(dofs[i], ndofs[j], vi_ue[i][j]);
          if (ve_ui[i][j] != 0.0) M->add(ndofs[i], dofs[j], ve_ui[i][j]);
          if (ve_ue[i][j] != 0.0) M->add(ndofs[i], ndofs[j], ve_ue[i][j]);
        }
  };

  for (int cell = 0; cell < mesh.n_cells; ++cell) {
    pair.get_dof_indices(cell, dofs);
    for (int face_no = 0; face_no < 4; ++face_no) {
      const int kind = mesh.face_kind[4 * cell + face_no];
      if (kind == FACE_BOUNDARY) continue;
      const int neighbor_face_no = face_no ^ 1; // neighbor_of_neighbor on these uniformly oriented meshes
      if (kind == FACE_HAS_CHILDREN) {
        for (int subface_no = 0; subface_no < 2; ++subface_no) {
          const int neighbor_child = subface_no == 0 ? mesh.neighbor[4 * cell + face_no] : mesh.neighbor2[4 * cell + face_no];
          subface_values.reinit(mesh.v(cell), face_no, subface_no);
          neighbor_face_values.reinit(mesh.v(neighbor_child), neighbor_face_no);
          pair.get_dof_indices(neighbor_child, ndofs);
          // h of the (refined, inactive) neighbour itself, as in reference LDG.cpp:369-370
          const double h = std::min(mesh.diameter(cell), mesh.nb_parent_diameter[4 * cell + face_no]);
          local_flux(subface_values, neighbor_face_values, pair.penalty / h);
        }
      } else if (kind == FACE_SAME_LEVEL) {
        const int neighbor = mesh.neighbor[4 * cell + face_no];
        if (neighbor > cell) { // the cell with the lower index does the work
          face_values.reinit(mesh.v(cell), face_no);
          neighbor_face_values.reinit(mesh.v(neighbor), neighbor_face_no);
          pair.get_dof_indices(neighbor, ndofs);
          const double h = std::min(mesh.diameter(cell), mesh.diameter(neighbor));
          local_flux(face_values, neighbor_face_values, pair.penalty / h);
        }
      }
      // FACE_COARSER: the coarse neighbour assembles this face
    }
  }
}

void SolarCellProblem::assemble_LDG_system(double transient_or_steady) { // reference SolarCell.cpp:822-932
  assemble_local_LDG(electron_hole_pair, semiconductor_mesh, transient_or_steady);
  assemble_flux_terms(electron_hole_pair, semiconductor_mesh);
  if (full_system) {
    assemble_local_LDG(redox_pair, electrolyte_mesh, transient_or_steady);
    assemble_flux_terms(redox_pair, electrolyte_mesh);
  }
}

void SolarCellProblem::set_solvers() { // reference SolarCell.cpp:1733-1747
  {
    const int n = Poisson_object.n_rt + Poisson_object.n_cells;
    std::vector<double> cx(n), cy(n);
    for (int c = 0; c < Poisson_mesh.n_cells; ++c) {
      Poisson_mesh.center(c, cx[Poisson_object.n_rt + c], cy[Poisson_object.n_rt + c]);
      for (int f = 0; f < 4; ++f) {
        const int d = Poisson_object.face_dof[4 * c + f];
        Poisson_mesh.face_center(c, f, cx[d], cy[d]);
      }
    }
    Poisson_object.elimination_order = nested_dissection_order(Poisson_object.system_matrix, cx, cy);
    Poisson_object.set_solver();
  }
  auto do_pair = [&](CarrierPair& pair, const Mesh& mesh) {
    const int n = 12 * mesh.n_cells;
    std::vector<double> cx(n), cy(n);
    std::vector<int> dofs(12);
    for (int c = 0; c < mesh.n_cells; ++c) {
      double x, y;
      mesh.center(c, x, y);
      pair.get_dof_indices(c, dofs);
      for (int i = 0; i < 12; ++i) {
        cx[dofs[i]] = x;
        cy[dofs[i]] = y;
      }
    }
    pair.elimination_order = nested_dissection_order(pair.carrier_1.system_matrix, cx, cy);
#pragma omp parallel sections
    {
#pragma omp section
      pair.carrier_1.set_solver(pair.elimination_order);
#pragma omp section
      pair.carrier_2.set_solver(pair.elimination_order);
    }
  };
  do_pair(electron_hole_pair, semiconductor_mesh);
  if (full_system) do_pair(redox_pair, electrolyte_mesh);
}

void SolarCellProblem::project_initial_conditions() {
  // VectorTools::project of a constant onto DG = that constant in the density dofs
  auto fill = [](CarrierPair& pair, double v1, double v2) {
    const int n = pair.n_cells;
    std::fill(pair.carrier_1.solution.begin(), pair.carrier_1.solution.end(), 0.0);
    std::fill(pair.carrier_2.solution.begin(), pair.carrier_2.solution.end(), 0.0);
    for (int k = 8 * n; k < 12 * n; ++k) {
      pair.carrier_1.solution[k] = v1;
      pair.carrier_2.solution[k] = v2;
    }
  };
  fill(electron_hole_pair, prm[P_RHO_N_E], prm[P_RHO_P_E]);
  if (full_system) fill(redox_pair, prm[P_RHO_R_E], prm[P_RHO_O_E]);
}

// ------------------------------------------------------------------------------------------------ carrier rhs
double SolarCellProblem::generation(const Tensor1& p) const { // reference Generation.cpp:29-44
  return prm[P_GEN_ALPHA] * prm[P_GEN_FLUX] * std::exp(prm[P_GEN_ALPHA] * (p.c[1] - prm[P_GEN_LOCATION]));
}
// reference include/SolarCell.hpp:86-98.  The reference's function returns 0.0 and carries the Shockley-Read-Hall formula
// as a comment; P_SRH != 0 switches that commented formula on, exactly as written there.
static double SRH_Recombination(double electron_density, double hole_density, const double* prm) {
  if (prm[P_SRH] == 0.0) return 0.0;
  const double ni = prm[P_N_INTRINSIC];
  return (ni * ni - electron_density * hole_density) /
         (prm[P_TAU_N] * (electron_density - ni) + prm[P_TAU_P] * (hole_density - ni));
}

namespace {
// FEValues[Density].get_function_values
void density_values(const FEValuesBase& fe, const std::vector<double>& solution, int n_cells, int cell, double* out) {
  for (int q = 0; q < fe.n_q; ++q) {
    double s = 0;
    for (int a = 0; a < 4; ++a) s += solution[8 * n_cells + 4 * cell + a] * fe.N_[a][q];
    out[q] = s;
  }
}
// Poisson_fe_values[ElectricField].get_function_values
void field_values(const FEValuesBase& fe, const PoissonData& P, int pcell, Tensor1* out) {
  for (int q = 0; q < fe.n_q; ++q) {
    Tensor1 s{{0, 0}};
    for (int f = 0; f < 4; ++f) {
      const double X = P.solution[P.face_dof[4 * pcell + f]];
      s.c[0] += X * fe.rt_[f][q].c[0];
      s.c[1] += X * fe.rt_[f][q].c[1];
    }
    out[q] = s;
  }
}
} // namespace

void SolarCellProblem::assemble_local_semiconductor_rhs(int cell, std::vector<double>& rhs1,
                                                        std::vector<double>& rhs2) const {
  // reference SolarCell.cpp:1073-1414
  const Mesh& mesh = semiconductor_mesh;
  const int n = mesh.n_cells;
  const double h = mesh.diameter(cell);
  const double penalty = electron_hole_pair.penalty;
  const int Poisson_cell = s_2_p_map.at(cell);
  FEValues carrier_fe_values, Poisson_fe_values;
  carrier_fe_values.reinit(mesh.v(cell));
  Poisson_fe_values.reinit(Poisson_mesh.v(Poisson_cell));
  std::fill(rhs1.begin(), rhs1.end(), 0.0);
  std::fill(rhs2.begin(), rhs2.end(), 0.0);
  double old1[9], old2[9], gen[9];
  Tensor1 E[9];
  density_values(carrier_fe_values, electron_hole_pair.carrier_1.solution, n, cell, old1);
  density_values(carrier_fe_values, electron_hole_pair.carrier_2.solution, n, cell, old2);
  for (int q = 0; q < 9; ++q) gen[q] = generation(carrier_fe_values.quadrature_point(q));
  field_values(Poisson_fe_values, Poisson_object, Poisson_cell, E);
  const double inverse_perm = 1.0 / prm[P_EPS_S];
  const double z1 = electron_hole_pair.carrier_1.charge_number, z2 = electron_hole_pair.carrier_2.charge_number;
  for (int q = 0; q < 9; ++q)
    for (int i = 0; i < 12; ++i) {
      const double psi_i_density = carrier_fe_values.density_value(i, q);
      const Tensor1 psi_i_current = carrier_fe_values.current_value(i, q);
      rhs1[i] += (psi_i_density * gen[q] + psi_i_density * SRH_Recombination(old1[q], old2[q], prm) +
                  z1 * (psi_i_current * E[q]) * inverse_perm * old1[q]) * carrier_fe_values.JxW(q);
      rhs2[i] += (psi_i_density * gen[q] + psi_i_density * SRH_Recombination(old1[q], old2[q], prm) +
                  z2 * (psi_i_current * E[q]) * inverse_perm * old2[q]) * carrier_fe_values.JxW(q);
    }
  FEFaceValues face_values, neighbor_face_values;
  for (int face_no = 0; face_no < 4; ++face_no) {
    if (mesh.face_kind[4 * cell + face_no] != FACE_BOUNDARY) continue;
    face_values.reinit(mesh.v(cell), face_no);
    const int bid = mesh.boundary_id[4 * cell + face_no];
    const double bc1 = prm[P_RHO_N_E], bc2 = prm[P_RHO_P_E]; // Electrons_/Holes_Equilibrium, density component
    if (bid == Dirichlet) {
      for (int q = 0; q < 3; ++q)
        for (int i = 0; i < 12; ++i) {
          const double t = -1.0 * (face_values.current_value(i, q) * face_values.normal_vector(q)) +
                           (penalty / h) * face_values.density_value(i, q);
          rhs1[i] += t * bc1 * face_values.JxW(q);
          rhs2[i] += t * bc2 * face_values.JxW(q);
        }
    } else if (bid == Interface) {
      double ne[3], pe[3], red[3], ox[3];
      density_values(face_values, electron_hole_pair.carrier_1.solution, n, cell, ne);
      density_values(face_values, electron_hole_pair.carrier_2.solution, n, cell, pe);
      const int interface_index = semi_interface_map.at(cell);
      const int ncell = elec_interface_cells[interface_index];
      neighbor_face_values.reinit(electrolyte_mesh.v(ncell), elec_interface_faces[interface_index]);
      density_values(neighbor_face_values, redox_pair.carrier_1.solution, electrolyte_mesh.n_cells, ncell, red);
      density_values(neighbor_face_values, redox_pair.carrier_2.solution, electrolyte_mesh.n_cells, ncell, ox);
      for (int q = 0; q < 3; ++q)
        for (int i = 0; i < 12; ++i) {
          const double psi = face_values.density_value(i, q);
          rhs1[i] += -1.0 * psi * prm[P_K_ET] * (ne[q] - bc1) * ox[q] * face_values.JxW(q);
          rhs2[i] += +1.0 * psi * prm[P_K_HT] * (pe[q] - bc2) * red[q] * face_values.JxW(q);
        }
    } else if (bid == Schottky) {
      double ne[3], pe[3];
      density_values(face_values, electron_hole_pair.carrier_1.solution, n, cell, ne);
      density_values(face_values, electron_hole_pair.carrier_2.solution, n, cell, pe);
      for (int q = 0; q < 3; ++q)
        for (int i = 0; i < 12; ++i) {
          const double psi = face_values.density_value(i, q);
          rhs1[i] += -1.0 * psi * prm[P_V_N] * (ne[q] - bc1) * face_values.JxW(q);
          rhs2[i] += +1.0 * psi * prm[P_V_P] * (pe[q] - bc2) * face_values.JxW(q);
        }
    } else if (bid == Neumann) {
      // nothing to do if insulating
    } else {
      throw std::runtime_error("semiconductor rhs: unknown boundary id");
    }
  }
}

void SolarCellProblem::assemble_local_electrolyte_rhs(int cell, std::vector<double>& rhs1,
                                                      std::vector<double>& rhs2) const {
  // reference SolarCell.cpp:1451-1726
  const Mesh& mesh = electrolyte_mesh;
  const int n = mesh.n_cells;
  const double h = mesh.diameter(cell);
  const double penalty = redox_pair.penalty;
  const int Poisson_cell = e_2_p_map.at(cell);
  FEValues carrier_fe_values, Poisson_fe_values;
  carrier_fe_values.reinit(mesh.v(cell));
  Poisson_fe_values.reinit(Poisson_mesh.v(Poisson_cell));
  std::fill(rhs1.begin(), rhs1.end(), 0.0);
  std::fill(rhs2.begin(), rhs2.end(), 0.0);
  double old1[9], old2[9];
  Tensor1 E[9];
  density_values(carrier_fe_values, redox_pair.carrier_1.solution, n, cell, old1);
  density_values(carrier_fe_values, redox_pair.carrier_2.solution, n, cell, old2);
  field_values(Poisson_fe_values, Poisson_object, Poisson_cell, E);
  const double inverse_perm = 1.0 / prm[P_EPS_E];
  const double z1 = redox_pair.carrier_1.charge_number, z2 = redox_pair.carrier_2.charge_number;
  for (int q = 0; q < 9; ++q)
    for (int i = 0; i < 12; ++i) {
      const Tensor1 psi_i_current = carrier_fe_values.current_value(i, q);
      rhs1[i] += (z1 * (psi_i_current * E[q]) * inverse_perm * old1[q]) * carrier_fe_values.JxW(q);
      rhs2[i] += (z2 * (psi_i_current * E[q]) * inverse_perm * old2[q]) * carrier_fe_values.JxW(q);
    }
  FEFaceValues face_values, neighbor_face_values;
  for (int face_no = 0; face_no < 4; ++face_no) {
    if (mesh.face_kind[4 * cell + face_no] != FACE_BOUNDARY) continue;
    face_values.reinit(mesh.v(cell), face_no);
    const int bid = mesh.boundary_id[4 * cell + face_no];
    if (bid == Dirichlet) {
      const double bc1 = prm[P_RHO_R_E], bc2 = prm[P_RHO_O_E];
      for (int q = 0; q < 3; ++q)
        for (int i = 0; i < 12; ++i) {
          const double t = -1.0 * (face_values.current_value(i, q) * face_values.normal_vector(q)) +
                           (penalty / h) * face_values.density_value(i, q);
          rhs1[i] += t * bc1 * face_values.JxW(q);
          rhs2[i] += t * bc2 * face_values.JxW(q);
        }
    } else if (bid == Interface) {
      double red[3], ox[3], ne[3], pe[3];
      density_values(face_values, redox_pair.carrier_1.solution, n, cell, red);
      density_values(face_values, redox_pair.carrier_2.solution, n, cell, ox);
      const int interface_index = elec_interface_map.at(cell);
      const int ncell = semi_interface_cells[interface_index];
      neighbor_face_values.reinit(semiconductor_mesh.v(ncell), semi_interface_faces[interface_index]);
      const double bc1 = prm[P_RHO_N_E], bc2 = prm[P_RHO_P_E];
      density_values(neighbor_face_values, electron_hole_pair.carrier_1.solution, semiconductor_mesh.n_cells, ncell, ne);
      density_values(neighbor_face_values, electron_hole_pair.carrier_2.solution, semiconductor_mesh.n_cells, ncell, pe);
      for (int q = 0; q < 3; ++q)
        for (int i = 0; i < 12; ++i) {
          const double psi = face_values.density_value(i, q);
          const double current = -1.0 * psi * prm[P_K_ET] * (ne[q] - bc1) * ox[q] * face_values.JxW(q) +
                                 1.0 * psi * prm[P_K_HT] * (pe[q] - bc2) * red[q] * face_values.JxW(q);
          rhs1[i] += current;
          rhs2[i] += -1.0 * current;
        }
    } else if (bid == Neumann) {
      // nothing to do if insulating
    } else {
      throw std::runtime_error("electrolyte rhs: unknown boundary id");
    }
  }
}

namespace {
// the WorkStream pattern: parallel workers with private scratch; DG rows of different cells are disjoint, so the
// "copier" can write straight into the global vector and the summation order per entry is fixed.
template <class Local>
void run_pair_rhs(CarrierPair& pair, const Mesh& mesh, Local&& local) {
  // system_rhs = M u^{k-1}  (reference SolarCell.cpp:1043-1047)
  pair.mass_matrix.vmult(pair.carrier_1.system_rhs, pair.carrier_1.solution);
  pair.mass_matrix.vmult(pair.carrier_2.system_rhs, pair.carrier_2.solution);
#pragma omp parallel
  {
    std::vector<double> r1(12), r2(12);
    std::vector<int> dofs(12);
#pragma omp for schedule(static)
    for (int cell = 0; cell < mesh.n_cells; ++cell) {
      local(cell, r1, r2);
      pair.get_dof_indices(cell, dofs);
      for (int i = 0; i < 12; ++i) {
        pair.carrier_1.system_rhs[dofs[i]] += r1[i];
        pair.carrier_2.system_rhs[dofs[i]] += r2[i];
      }
    }
  }
}
} // namespace

void SolarCellProblem::assemble_semiconductor_rhs() { // reference SolarCell.cpp:1037-1071
  run_pair_rhs(electron_hole_pair, semiconductor_mesh,
               [&](int c, std::vector<double>& a, std::vector<double>& b) { assemble_local_semiconductor_rhs(c, a, b); });
}
void SolarCellProblem::assemble_electrolyte_rhs() { // reference SolarCell.cpp:1417-1449
  run_pair_rhs(redox_pair, electrolyte_mesh,
               [&](int c, std::vector<double>& a, std::vector<double>& b) { assemble_local_electrolyte_rhs(c, a, b); });
}

void SolarCellProblem::solve_full_system() { // reference SolarCell.cpp:1758-1782: four concurrent tasks
#pragma omp parallel sections
  {
#pragma omp section
    electron_hole_pair.carrier_1.solve();
#pragma omp section
    electron_hole_pair.carrier_2.solve();
#pragma omp section
    { if (full_system) redox_pair.carrier_1.solve(); }
#pragma omp section
    { if (full_system) redox_pair.carrier_2.solve(); }
  }
}

// ------------------------------------------------------------------------------------------------ Poisson rhs
void SolarCellProblem::assemble_local_Poisson_rhs(int cell, bool semiconductor, std::vector<int>& dofs,
                                                  std::vector<double>& rhs) const {
  // reference SolarCell.cpp:487-684 (semiconductor) and 686-815 (electrolyte)
  const Mesh& mesh = semiconductor ? semiconductor_mesh : electrolyte_mesh;
  const CarrierPair& pair = semiconductor ? electron_hole_pair : redox_pair;
  const int Poisson_cell = semiconductor ? s_2_p_map.at(cell) : e_2_p_map.at(cell);
  std::fill(rhs.begin(), rhs.end(), 0.0);
  Poisson_object.get_dof_indices(Poisson_cell, dofs);
  FEValues Poisson_fe_values, carrier_fe_values;
  Poisson_fe_values.reinit(Poisson_mesh.v(Poisson_cell));
  carrier_fe_values.reinit(mesh.v(cell));
  double old1[9], old2[9];
  density_values(carrier_fe_values, pair.carrier_1.solution, mesh.n_cells, cell, old1);
  density_values(carrier_fe_values, pair.carrier_2.solution, mesh.n_cells, cell, old2);
  const double donor = semiconductor ? prm[P_RHO_N_E] : 0.0, acceptor = semiconductor ? prm[P_RHO_P_E] : 0.0;
  for (int q = 0; q < 9; ++q)
    for (int i = 0; i < 5; ++i)
      rhs[i] += -Poisson_fe_values.potential_value(i, q) *
                ((donor - acceptor) + (pair.carrier_1.charge_number * old1[q] + pair.carrier_2.charge_number * old2[q])) *
                Poisson_fe_values.JxW(q);
  FEFaceValues face_values;
  for (int face_no = 0; face_no < 4; ++face_no) {
    if (Poisson_mesh.face_kind[4 * Poisson_cell + face_no] != FACE_BOUNDARY) continue;
    const int bid = Poisson_mesh.boundary_id[4 * Poisson_cell + face_no];
    if (!(bid == Dirichlet || (semiconductor && bid == Schottky))) continue;
    face_values.reinit(Poisson_mesh.v(Poisson_cell), face_no);
    for (int q = 0; q < 3; ++q) {
      const Tensor1& p = face_values.quadrature_point(q);
      double value;
      if (semiconductor) {
        const double bi = (p.c[0] == 0.0) ? prm[P_PHI_BI] : 0.0; // Built_In_Bias, reference BiasValues.cpp:11-31
        double bc;
        if (bid == Dirichlet)
          bc = (p.c[0] == 0.0) ? prm[P_PHI_APP] : 0.0; // Applied_Bias, BiasValues.cpp:75-94
        else
          bc = (p.c[1] == prm[P_SCH_LOCATION]) ? prm[P_PHI_SCH] : 0.0; // Schottky_Bias, BiasValues.cpp:49-65
        value = bi - bc;
      } else {
        value = 0.0; // Bulk_Bias, BiasValues.cpp:99-115
      }
      for (int i = 0; i < 5; ++i)
        rhs[i] += -((face_values.field_value(i, q) * face_values.normal_vector(q)) * value * face_values.JxW(q));
    }
  }
}

void SolarCellProblem::assemble_Poisson_rhs() { // reference SolarCell.cpp:430-485
  std::fill(Poisson_object.system_rhs.begin(), Poisson_object.system_rhs.end(), 0.0);
  std::vector<int> dofs(5);
  std::vector<double> rhs(5);
  // the copier runs serially in cell order (constrained scatter), reference SolarCell.cpp:419-428
  for (int cell = 0; cell < semiconductor_mesh.n_cells; ++cell) {
    assemble_local_Poisson_rhs(cell, true, dofs, rhs);
    Poisson_object.constraints.distribute_local_to_global(rhs, dofs, Poisson_object.system_rhs);
  }
  if (full_system)
    for (int cell = 0; cell < electrolyte_mesh.n_cells; ++cell) {
      assemble_local_Poisson_rhs(cell, false, dofs, rhs);
      Poisson_object.constraints.distribute_local_to_global(rhs, dofs, Poisson_object.system_rhs);
    }
}

void SolarCellProblem::solve_Poisson() { Poisson_object.solve(); } // reference SolarCell.cpp:1750-1756

// ------------------------------------------------------------------------------------------------ manufactured tests
namespace test_functions { // reference source/test_functions.cpp
const double two_pi = 2 * M_PI;
// test_Poisson (:13-93)
double Poisson_rhs(double x, double y) { return 4 * M_PI * M_PI * (std::cos(two_pi * y) - std::sin(two_pi * x)); }
double Poisson_bc(double x, double y) { return std::cos(two_pi * y) - std::sin(two_pi * x) - x; }
void Poisson_solution(double x, double y, double v[3]) {
  v[0] = 1 + two_pi * std::cos(two_pi * x);
  v[1] = two_pi * std::sin(two_pi * y);
  v[2] = std::cos(two_pi * y) - std::sin(two_pi * x) - x;
}
// test_LDG_IMEX (:99-220)
double LDG_rhs(double x, double y, double t) {
  return -std::exp(-t) + two_pi * two_pi * std::cos(two_pi * x) + two_pi * two_pi * std::cos(two_pi * y) +
         two_pi * std::sin(two_pi * x);
}
double LDG_bc(double x, double y, double t) { return std::exp(-t) + std::cos(two_pi * x) + std::cos(two_pi * y); }
double LDG_interface(double, double y, double t) { return -std::exp(-t) - std::cos(two_pi * y) - 1; }
void LDG_solution(double x, double y, double t, double v[3]) {
  v[2] = std::exp(-t) + std::cos(two_pi * x) + std::cos(two_pi * y);
  v[0] = two_pi * std::sin(two_pi * x) - v[2];
  v[1] = two_pi * std::sin(two_pi * y);
}
// test_DD_Poisson (:227-361)
double DD_rhs(double x, double y, double t) {
  const double u = std::exp(-t) + std::cos(two_pi * x) + std::cos(two_pi * y);
  const double div_E_u = two_pi * two_pi * (std::cos(two_pi * y) - std::sin(two_pi * x)) * u -
                         two_pi * (two_pi * std::cos(two_pi * x) + 1) * std::sin(two_pi * x) -
                         two_pi * two_pi * std::sin(two_pi * y) * std::sin(two_pi * y);
  return -std::exp(-t) + two_pi * two_pi * std::cos(two_pi * x) + two_pi * two_pi * std::cos(two_pi * y) - div_E_u;
}
double DD_Poisson_rhs(double x, double y, double t) {
  return 4 * M_PI * M_PI * (std::cos(2 * M_PI * y) - std::sin(2 * M_PI * x)) + std::exp(-t) + std::cos(two_pi * x) +
         std::cos(two_pi * y);
}
void DD_solution(double x, double y, double t, double v[3]) {
  v[2] = std::exp(-t) + std::cos(two_pi * x) + std::cos(two_pi * y);
  v[0] = two_pi * std::sin(two_pi * x) - (two_pi * std::cos(two_pi * x) + 1) * v[2];
  v[1] = two_pi * std::sin(two_pi * y) - two_pi * std::sin(two_pi * y) * v[2];
}
// test_interface_problem::InitialConditions (:429-448): used as IC by ALL transient tests
double initial_condition(double x, double y) { return 1 + std::cos(two_pi * x) * std::cos(two_pi * y); }
} // namespace test_functions

void SolarCellProblem::project_test_initial_condition() {
  // VectorTools::project with QGauss(degree+1) (reference SolarCell.cpp:2898-2902): cell-local 4x4 mass solve
  const Mesh& mesh = semiconductor_mesh;
  std::vector<double>& u = electron_hole_pair.carrier_1.solution;
  std::fill(u.begin(), u.end(), 0.0);
  const double g2[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  for (int c = 0; c < mesh.n_cells; ++c) {
    const double* vt = mesh.v(c);
    double M[4][5] = {};
    for (int qy = 0; qy < 2; ++qy)
      for (int qx = 0; qx < 2; ++qx) {
        const double xi = g2[qx], eta = g2[qy];
        const double N[4] = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
        const double dxi[4] = {-(1 - eta), (1 - eta), -eta, eta}, deta[4] = {-(1 - xi), -xi, (1 - xi), xi};
        double J00 = 0, J01 = 0, J10 = 0, J11 = 0, x = 0, y = 0;
        for (int a = 0; a < 4; ++a) {
          J00 += vt[2 * a] * dxi[a];
          J01 += vt[2 * a] * deta[a];
          J10 += vt[2 * a + 1] * dxi[a];
          J11 += vt[2 * a + 1] * deta[a];
          x += vt[2 * a] * N[a];
          y += vt[2 * a + 1] * N[a];
        }
        const double JxW = (J00 * J11 - J01 * J10) * 0.25;
        const double f = test_functions::initial_condition(x, y);
        for (int a = 0; a < 4; ++a) {
          for (int b = 0; b < 4; ++b) M[a][b] += N[a] * N[b] * JxW;
          M[a][4] += N[a] * f * JxW;
        }
      }
    // Gaussian elimination with partial pivoting on the 4x4 system
    for (int k = 0; k < 4; ++k) {
      int p = k;
      for (int r = k + 1; r < 4; ++r)
        if (std::fabs(M[r][k]) > std::fabs(M[p][k])) p = r;
      if (p != k)
        for (int j = 0; j < 5; ++j) std::swap(M[k][j], M[p][j]);
      for (int r = k + 1; r < 4; ++r) {
        const double m = M[r][k] / M[k][k];
        for (int j = k; j < 5; ++j) M[r][j] -= m * M[k][j];
      }
    }
    double sol[4];
    for (int k = 3; k >= 0; --k) {
      double s = M[k][4];
      for (int j = k + 1; j < 4; ++j) s -= M[k][j] * sol[j];
      sol[k] = s / M[k][k];
    }
    for (int a = 0; a < 4; ++a) u[8 * mesh.n_cells + 4 * c + a] = sol[a];
  }
}

void SolarCellProblem::assemble_test_steady_rhs() {
  // Poisson: reference MixedFEM.cpp:166-254 ; LDG (carrier 1): reference LDG.cpp:681-800
  std::fill(Poisson_object.system_rhs.begin(), Poisson_object.system_rhs.end(), 0.0);
  std::vector<int> dofs(5);
  std::vector<double> rhs(5);
  FEValues fe;
  FEFaceValues ffe;
  for (int cell = 0; cell < Poisson_mesh.n_cells; ++cell) {
    fe.reinit(Poisson_mesh.v(cell));
    Poisson_object.get_dof_indices(cell, dofs);
    std::fill(rhs.begin(), rhs.end(), 0.0);
    for (int q = 0; q < 9; ++q)
      for (int i = 0; i < 5; ++i)
        rhs[i] += -fe.potential_value(i, q) *
                  test_functions::Poisson_rhs(fe.quadrature_point(q).c[0], fe.quadrature_point(q).c[1]) * fe.JxW(q);
    for (int face_no = 0; face_no < 4; ++face_no)
      if (Poisson_mesh.face_kind[4 * cell + face_no] == FACE_BOUNDARY &&
          Poisson_mesh.boundary_id[4 * cell + face_no] == Dirichlet) {
        ffe.reinit(Poisson_mesh.v(cell), face_no);
        for (int q = 0; q < 3; ++q)
          for (int i = 0; i < 5; ++i)
            rhs[i] += -((ffe.field_value(i, q) * ffe.normal_vector(q)) *
                        test_functions::Poisson_bc(ffe.quadrature_point(q).c[0], ffe.quadrature_point(q).c[1]) * ffe.JxW(q));
      }
    Poisson_object.constraints.distribute_local_to_global(rhs, dofs, Poisson_object.system_rhs);
  }
  const Mesh& mesh = semiconductor_mesh;
  std::vector<double>& g = electron_hole_pair.carrier_1.system_rhs;
  std::fill(g.begin(), g.end(), 0.0);
  std::vector<int> cdofs(12);
  for (int cell = 0; cell < mesh.n_cells; ++cell) {
    fe.reinit(mesh.v(cell));
    electron_hole_pair.get_dof_indices(cell, cdofs);
    const double h = mesh.diameter(cell);
    double r[12] = {};
    for (int q = 0; q < 9; ++q)
      for (int i = 0; i < 12; ++i)
        r[i] += fe.density_value(i, q) *
                test_functions::Poisson_rhs(fe.quadrature_point(q).c[0], fe.quadrature_point(q).c[1]) * fe.JxW(q);
    for (int face_no = 0; face_no < 4; ++face_no)
      if (mesh.face_kind[4 * cell + face_no] == FACE_BOUNDARY && mesh.boundary_id[4 * cell + face_no] == Dirichlet) {
        ffe.reinit(mesh.v(cell), face_no);
        for (int q = 0; q < 3; ++q)
          for (int i = 0; i < 12; ++i)
            r[i] += (-1.0 * (ffe.current_value(i, q) * ffe.normal_vector(q)) +
                     (electron_hole_pair.penalty / h) * ffe.density_value(i, q)) *
                    test_functions::Poisson_bc(ffe.quadrature_point(q).c[0], ffe.quadrature_point(q).c[1]) * ffe.JxW(q);
      }
    for (int i = 0; i < 12; ++i) g[cdofs[i]] += r[i];
  }
}

void SolarCellProblem::assemble_test_transient_rhs(double time) {
  // reference SolarCell.cpp:2908-2932 + LDG.cpp:802-982; field fixed to (1,0)
  CarrierPair& pair = electron_hole_pair;
  const Mesh& mesh = semiconductor_mesh;
  pair.mass_matrix.vmult(pair.carrier_1.system_rhs, pair.carrier_1.solution);
  std::vector<int> dofs(12);
  FEValues fe;
  FEFaceValues ffe;
  const Tensor1 field{{1.0, 0.0}};
  for (int cell = 0; cell < mesh.n_cells; ++cell) {
    fe.reinit(mesh.v(cell));
    pair.get_dof_indices(cell, dofs);
    const double h = mesh.diameter(cell);
    double r[12] = {}, old1[9];
    density_values(fe, pair.carrier_1.solution, mesh.n_cells, cell, old1);
    for (int q = 0; q < 9; ++q)
      for (int i = 0; i < 12; ++i)
        r[i] += (fe.density_value(i, q) *
                     test_functions::LDG_rhs(fe.quadrature_point(q).c[0], fe.quadrature_point(q).c[1], time) -
                 (fe.current_value(i, q) * field) * old1[q]) * fe.JxW(q);
    for (int face_no = 0; face_no < 4; ++face_no) {
      if (mesh.face_kind[4 * cell + face_no] != FACE_BOUNDARY) continue;
      ffe.reinit(mesh.v(cell), face_no);
      const int bid = mesh.boundary_id[4 * cell + face_no];
      if (bid == Dirichlet) {
        for (int q = 0; q < 3; ++q)
          for (int i = 0; i < 12; ++i)
            r[i] += (-1.0 * (ffe.current_value(i, q) * ffe.normal_vector(q)) + (pair.penalty / h) * ffe.density_value(i, q)) *
                    test_functions::LDG_bc(ffe.quadrature_point(q).c[0], ffe.quadrature_point(q).c[1], time) * ffe.JxW(q);
      } else if (bid == Interface) {
        for (int q = 0; q < 3; ++q)
          for (int i = 0; i < 12; ++i)
            r[i] += -1.0 * ffe.density_value(i, q) *
                    test_functions::LDG_interface(ffe.quadrature_point(q).c[0], ffe.quadrature_point(q).c[1], time) *
                    ffe.JxW(q);
      }
    }
    for (int i = 0; i < 12; ++i) pair.carrier_1.system_rhs[dofs[i]] += r[i];
  }
}

void SolarCellProblem::assemble_coupled_Poisson_test_rhs(double time) {
  // reference SolarCell.cpp:2108-2223 (loop over semiconductor cells, scatter into the mapped Poisson cell)
  std::fill(Poisson_object.system_rhs.begin(), Poisson_object.system_rhs.end(), 0.0);
  std::vector<int> dofs(5);
  std::vector<double> rhs(5);
  FEValues pfe, cfe;
  FEFaceValues ffe;
  for (int cell = 0; cell < semiconductor_mesh.n_cells; ++cell) {
    const int pc = s_2_p_map.at(cell);
    Poisson_object.get_dof_indices(pc, dofs);
    pfe.reinit(Poisson_mesh.v(pc));
    cfe.reinit(semiconductor_mesh.v(cell));
    double old1[9];
    density_values(cfe, electron_hole_pair.carrier_1.solution, semiconductor_mesh.n_cells, cell, old1);
    std::fill(rhs.begin(), rhs.end(), 0.0);
    for (int q = 0; q < 9; ++q)
      for (int i = 0; i < 5; ++i)
        rhs[i] += -pfe.potential_value(i, q) *
                  (test_functions::DD_Poisson_rhs(pfe.quadrature_point(q).c[0], pfe.quadrature_point(q).c[1], time) - old1[q]) *
                  pfe.JxW(q);
    for (int face_no = 0; face_no < 4; ++face_no)
      if (Poisson_mesh.face_kind[4 * pc + face_no] == FACE_BOUNDARY && Poisson_mesh.boundary_id[4 * pc + face_no] == Dirichlet) {
        ffe.reinit(Poisson_mesh.v(pc), face_no);
        for (int q = 0; q < 3; ++q)
          for (int i = 0; i < 5; ++i)
            rhs[i] += -((ffe.field_value(i, q) * ffe.normal_vector(q)) *
                        test_functions::Poisson_bc(ffe.quadrature_point(q).c[0], ffe.quadrature_point(q).c[1]) * ffe.JxW(q));
      }
    Poisson_object.constraints.distribute_local_to_global(rhs, dofs, Poisson_object.system_rhs);
  }
}

void SolarCellProblem::assemble_coupled_DD_test_rhs(double time) {
  // reference SolarCell.cpp:3052-3075 + 2228-2330
  CarrierPair& pair = electron_hole_pair;
  const Mesh& mesh = semiconductor_mesh;
  pair.mass_matrix.vmult(pair.carrier_1.system_rhs, pair.carrier_1.solution);
  std::vector<int> dofs(12);
  FEValues cfe, pfe;
  FEFaceValues ffe;
  for (int cell = 0; cell < mesh.n_cells; ++cell) {
    const int pc = s_2_p_map.at(cell);
    pfe.reinit(Poisson_mesh.v(pc));
    cfe.reinit(mesh.v(cell));
    pair.get_dof_indices(cell, dofs);
    const double h = mesh.diameter(cell);
    double r[12] = {}, old1[9];
    Tensor1 E[9];
    density_values(cfe, pair.carrier_1.solution, mesh.n_cells, cell, old1);
    field_values(pfe, Poisson_object, pc, E);
    for (int q = 0; q < 9; ++q)
      for (int i = 0; i < 12; ++i)
        r[i] += (cfe.density_value(i, q) *
                     test_functions::DD_rhs(cfe.quadrature_point(q).c[0], cfe.quadrature_point(q).c[1], time) -
                 (cfe.current_value(i, q) * E[q]) * old1[q]) * cfe.JxW(q);
    for (int face_no = 0; face_no < 4; ++face_no)
      if (mesh.face_kind[4 * cell + face_no] == FACE_BOUNDARY && mesh.boundary_id[4 * cell + face_no] == Dirichlet) {
        ffe.reinit(mesh.v(cell), face_no);
        for (int q = 0; q < 3; ++q)
          for (int i = 0; i < 12; ++i)
            r[i] += (-1.0 * (ffe.current_value(i, q) * ffe.normal_vector(q)) + (pair.penalty / h) * ffe.density_value(i, q)) *
                    test_functions::LDG_bc(ffe.quadrature_point(q).c[0], ffe.quadrature_point(q).c[1], time) * ffe.JxW(q);
      }
    for (int i = 0; i < 12; ++i) pair.carrier_1.system_rhs[dofs[i]] += r[i];
  }
}

void SolarCellProblem::ldg_errors(int which, double time, double& density_error, double& current_error) const {
  // reference LDG.cpp:984-1133: L2 norms with QIterated(QTrapez, degree+2); which: 0 test_Poisson, 1 LDG_IMEX, 2 DD
  const Mesh& mesh = semiconductor_mesh;
  const std::vector<double>& u = electron_hole_pair.carrier_1.solution;
  const int n = mesh.n_cells;
  FEErrorValues fe;
  double e_u = 0, e_q = 0;
  for (int c = 0; c < n; ++c) {
    fe.reinit(mesh.v(c));
    for (int q = 0; q < 16; ++q) {
      double uh[3] = {0, 0, 0}, ex[3];
      for (int a = 0; a < 4; ++a)
        for (int comp = 0; comp < 3; ++comp) uh[comp] += u[comp * 4 * n + 4 * c + a] * fe.N16[a][q];
      const double x = fe.point16[q].c[0], y = fe.point16[q].c[1];
      if (which == 0)
        test_functions::Poisson_solution(x, y, ex);
      else if (which == 1)
        test_functions::LDG_solution(x, y, time, ex);
      else
        test_functions::DD_solution(x, y, time, ex);
      e_u += (uh[2] - ex[2]) * (uh[2] - ex[2]) * fe.JxW16[q];
      e_q += ((uh[0] - ex[0]) * (uh[0] - ex[0]) + (uh[1] - ex[1]) * (uh[1] - ex[1])) * fe.JxW16[q];
    }
  }
  density_error = std::sqrt(e_u);
  current_error = std::sqrt(e_q);
}

void SolarCellProblem::mixed_errors(double& potential_error, double& field_error) const {
  // reference MixedFEM.cpp:256-295 against test_Poisson::TrueSolution
  const Mesh& mesh = Poisson_mesh;
  FEErrorValues fe;
  double e_p = 0, e_d = 0;
  for (int c = 0; c < mesh.n_cells; ++c) {
    fe.reinit(mesh.v(c));
    const double phi = Poisson_object.solution[Poisson_object.n_rt + c];
    for (int q = 0; q < 16; ++q) {
      double D[2] = {0, 0}, ex[3];
      for (int f = 0; f < 4; ++f) {
        const double X = Poisson_object.solution[Poisson_object.face_dof[4 * c + f]];
        D[0] += X * fe.rt16[f][q].c[0];
        D[1] += X * fe.rt16[f][q].c[1];
      }
      test_functions::Poisson_solution(fe.point16[q].c[0], fe.point16[q].c[1], ex);
      e_p += (phi - ex[2]) * (phi - ex[2]) * fe.JxW16[q];
      e_d += ((D[0] - ex[0]) * (D[0] - ex[0]) + (D[1] - ex[1]) * (D[1] - ex[1])) * fe.JxW16[q];
    }
  }
  potential_error = std::sqrt(e_p);
  field_error = std::sqrt(e_d);
}

} // namespace oracle
