// SparseDirect.hpp -- nested-dissection multifrontal factorisation plan for the fixed system matrices.
//
// The reference factorises each constant matrix once with UMFPACK (SparseDirectUMFPACK::initialize, reference
// source/Carrier.cpp:26-32, source/Poisson.cpp:92-96) and then, every time step, runs UMFPACK's sequential sparse
// triangular substitutions (solver.vmult, source/Carrier.cpp:34-40, source/Poisson.cpp:98-105).  Sparse triangular
// substitution is the one thing a GPU is bad at, so the factorisation is organised differently here:
//
//   * unknowns are grouped into nodes (a cell's 12 LDG dofs; a Poisson cell's potential + the edge fluxes it owns);
//   * geometric nested dissection on the node graph gives a binary elimination tree of FRONTS (separators and
//     leaves); every front has np pivot unknowns and nb boundary unknowns (all of them in ancestor fronts);
//   * instead of L and U the explicit block operators are stored for every front
//         Inv = F_PP^-1,   G = F_BP F_PP^-1   (nb x np),   H = F_PP^-1 F_PB   (np x nb)
//     where F is the frontal matrix (original entries + Schur updates of the children);
//   * a solve is then two sweeps over the tree LEVELS made only of dense mat-vecs (no substitutions):
//         forward  (leaves -> root):  w_P = b_P - children's updates,   t = children's updates on B + G w_P
//         backward (root -> leaves):  x_P = Inv w_P - H x_B
//     All fronts of one level are independent: one batched, HBM-streaming kernel per level and sweep.
//
// This header holds the symbolic part (tree, index maps, level schedule) and the host numeric factorisation used
// for small problems and as the checker of the device factorisation.
#pragma once
#include <cstdint>
#include <vector>

#include "Csr.hpp"

namespace pecs {

struct Front {
  int np = 0, nb = 0;       // pivot / boundary unknowns
  int p0 = 0;               // pivots occupy positions [p0, p0+np) of the permuted vector
  int parent = -1;
  int child[2] = {-1, -1};
  int depth = 0;            // root = 0
  int64_t bd_off = 0;       // offset of this front's boundary index list in SolvePlan::bd_index
  int64_t fwd_off = 0;      // offset of G            (nb x np, row-major)          in the forward table
  int64_t bwd_off = 0;      // offset of [Inv | -H]   (np x (np+nb), row-major)     in the backward table
  int64_t upd_off = 0;      // offset of this front's update vector t (nb entries) in the update buffer
  int64_t cmap_off[2] = {0, 0}; // per child: for every local index l in [0, np+nb) the child's boundary slot or -1
};

struct SolvePlan {
  int n = 0;
  std::vector<int> perm;     // perm[dof] = position in the elimination order
  std::vector<int> iperm;    // iperm[position] = dof
  std::vector<Front> fronts; // postorder: children before parents, root last
  std::vector<int> bd_index; // concatenated boundary lists (positions in the permuted vector, ascending)
  std::vector<int> child_map; // concatenated inverse child maps (see Front::cmap_off)
  std::vector<std::vector<int>> levels; // fronts per depth
  int64_t fwd_entries = 0, bwd_entries = 0, upd_entries = 0;
  int max_np = 0, max_nb = 0;
  const int* bd(const Front& f) const { return bd_index.data() + f.bd_off; }
};

// node_of_dof[n] -> node id in [0, n_nodes); node_x/node_y: coordinates used for the geometric bisection.
// leaf_nodes: recursion stops at this many nodes.
SolvePlan build_solve_plan(const CsrMatrix& A, const std::vector<int>& node_of_dof, const std::vector<double>& node_x,
                           const std::vector<double>& node_y, int leaf_nodes);

// P A P^T (or its transpose) in CSR, P = plan.perm
CsrMatrix permute_csr(const CsrMatrix& A, const std::vector<int>& perm, bool transpose);

// Host numeric factorisation: fills fwd (all G) and bwd (all [Inv | -H]) tables.  Throws StatusError(PECS_ERR_SINGULAR)
// when a pivot block cannot be inverted.
void factorize_host(const SolvePlan& plan, const CsrMatrix& A, std::vector<double>& fwd, std::vector<double>& bwd);

// Host reference of the two solve sweeps (used by the CPU tests to validate plan + factor tables).
void solve_host(const SolvePlan& plan, const std::vector<double>& fwd, const std::vector<double>& bwd, const double* b,
                double* x);

} // namespace pecs
