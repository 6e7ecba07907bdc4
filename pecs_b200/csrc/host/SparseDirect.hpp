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
//   * a front hands its update t to its parent in the PARENT's local numbering ([pivots | boundary]): every front
//     owns one dense buffer per child, the child scatters into it through a precomputed map, the parent reads it
//     with unit stride.  Slots a child never writes stay zero from setup.
//
// Table layouts (what the device kernels stream): every row is padded to an even length so that all loads can be
// 16 bytes wide.  Backward table [Inv | -H]: row-major, np rows, leading dimension ld_bwd = even(np+nb).
// Forward table G: row-major (nb rows, ld_fwd = even(np)) for large fronts -- one warp per row -- and column-major
// (np columns of ld_fwd = even(nb) entries) for fronts with np <= kColMajorMaxNp -- one thread per row.
//
// This header holds the symbolic part (tree, index maps, level schedule) and the host numeric factorisation used
// for small problems and as the checker of the device factorisation.
#pragma once
#include <cstdint>
#include <vector>

#include "Csr.hpp"

namespace pecs {

constexpr int kColMajorMaxNp = 128;

struct Front {
  int np = 0, nb = 0;       // pivot / boundary unknowns
  int p0 = 0;               // pivots occupy positions [p0, p0+np) of the permuted vector
  int parent = -1;
  int child[2] = {-1, -1};
  int which_child = 0;      // this front is child[which_child] of its parent
  int depth = 0;            // root = 0
  int fwd_colmajor = 0;     // layout of G, see above
  int ld_fwd = 0, ld_bwd = 0;
  int64_t bd_off = 0;       // offset of this front's boundary index list in SolvePlan::bd_index (also of its out map)
  int64_t fwd_off = 0;      // offset of G in the forward table
  int64_t bwd_off = 0;      // offset of [Inv | -H] in the backward table
  int64_t cbuf_off[2] = {-1, -1}; // this front's dense update buffers, one per child, np+nb entries each (-1: no child)
  // G(i, j), i < nb, j < np
  int64_t fwd_index(int i, int j) const { return fwd_off + (fwd_colmajor ? (int64_t)j * ld_fwd + i : (int64_t)i * ld_fwd + j); }
  int64_t fwd_size() const { return (int64_t)(fwd_colmajor ? np : nb) * ld_fwd; }
  int64_t bwd_size() const { return (int64_t)np * ld_bwd; }
};

struct SolvePlan {
  int n = 0;
  std::vector<int> perm;     // perm[dof] = position in the elimination order
  std::vector<int> iperm;    // iperm[position] = dof
  std::vector<Front> fronts; // postorder: children before parents, root last
  std::vector<int> bd_index; // concatenated boundary lists (positions in the permuted vector, ascending)
  std::vector<int> out_map;  // same layout as bd_index: the parent-local index of every boundary unknown
  std::vector<std::vector<int>> levels; // fronts per depth
  int64_t fwd_entries = 0, bwd_entries = 0, upd_entries = 0; // upd_entries: total size of all child buffers
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
