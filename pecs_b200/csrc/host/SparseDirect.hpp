// SparseDirect.hpp -- nested-dissection multifrontal factorisation plan for the fixed system matrices.
//
// The reference factorises each constant matrix once with UMFPACK (SparseDirectUMFPACK::initialize, reference
// source/Carrier.cpp:26-32, source/Poisson.cpp:92-96) and then, every time step, runs UMFPACK's sequential sparse
// triangular substitutions (solver.vmult, source/Carrier.cpp:34-40, source/Poisson.cpp:98-105).  Sparse triangular
// substitution is the one thing a GPU is bad at, so the factorisation is organised differently here:
//
//   * an ELIMINATION TREE of fronts is built by geometric nested dissection (three builders below: vertex separators
//     on cell nodes, vertex separators on single unknowns with the bisection done on cells, and -- SolverSetup.cpp --
//     edge-flux separators for the saddle-point Poisson matrix); every front eliminates np pivot unknowns and has nb
//     boundary unknowns (all of them pivots of ancestor fronts);
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
// Table layout (what the device kernels stream, cuda/solve_kernels.cu): PANELS.  A table with R rows and N columns is
// cut into panels of P consecutive rows (P a power of two <= 32, chosen per front so that a panel is about 64 KB and a
// tree level always offers enough panels to occupy the whole GPU); a panel is stored column-major and contiguously:
//         entry (i, j)  ->  off + (i / P) * P * cols_pad + j * P + (i % P)
// Rows are zero-padded to a multiple of P, columns to a multiple of 32 / P.  One warp streams one panel with bulk
// asynchronous copies: 32 consecutive doubles of a panel are 32 / P columns of its P rows, so lane l always works
// for row l % P and the per-row sums need log2(32 / P) shuffles at the very end of the panel, nothing in between.
//
// This header holds the symbolic part (tree, index maps, level schedule) and the host numeric factorisation used
// for small problems and as the checker of the device factorisation.
#pragma once
#include <cstdint>
#include <vector>

#include "Csr.hpp"

namespace pecs {

constexpr int kSmallFrontMaxNp = 128;  // fronts up to this many pivots are factorised by one thread block each
constexpr int kTargetPanelsPerLevel = 148 * 16;
constexpr int kSmallTableDoubles = 2048; // tables up to 16 KB and 64 rows are one warp's work (per-warp mode of the kernels)
constexpr int kPanelTargetDoubles = 8192; // otherwise a panel holds about 64 KB: equal loads for the warps of a level
constexpr int kWarpsPerFront = 4;      // warps of the thread block that streams a front's panels (cuda/solve_kernels.cuh)

struct PanelTable {
  int rows = 0, cols = 0; // logical size
  int log2P = 0;          // panel height P = 1 << log2P
  int rows_pad = 0;       // multiple of P
  int cols_pad = 0;       // multiple of 32 / P
  int small = 0;          // the whole table is streamed by ONE warp (<= kSmallTableDoubles entries)
  int64_t off = 0;        // offset of the first panel in the table array (in doubles)
  int P() const { return 1 << log2P; }
  int n_panels() const { return rows_pad >> log2P; }
  int64_t panel_stride() const { return (int64_t)cols_pad << log2P; }
  int64_t size() const { return (int64_t)rows_pad * cols_pad; }
  int64_t index(int i, int j) const { return off + (int64_t)(i >> log2P) * panel_stride() + ((int64_t)j << log2P) + (i & (P() - 1)); }
};

struct Front {
  int np = 0, nb = 0;       // pivot / boundary unknowns
  int p0 = 0;               // pivots occupy positions [p0, p0+np) of the permuted vector
  int parent = -1;
  int child[2] = {-1, -1};
  int which_child = 0;      // this front is child[which_child] of its parent
  int depth = 0;            // root = 0
  PanelTable fwd;           // G          (nb x np)
  PanelTable bwd;           // [Inv | -H] (np x (np + nb))
  int64_t bd_off = 0;       // offset of this front's boundary index list in SolvePlan::bd_index (also of its out map)
  int64_t cbuf_off[2] = {-1, -1}; // this front's dense update buffers, one per child, np+nb entries each (-1: no child)
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
  // entries of all front operators without the padding of the panel layout: np^2 + 2 np nb per front
  int64_t logical_entries() const {
    int64_t e = 0;
    for (const Front& f : fronts) e += (int64_t)f.np * f.np + 2 * (int64_t)f.np * f.nb;
    return e;
  }
  const int* bd(const Front& f) const { return bd_index.data() + f.bd_off; }
};

// Elimination tree on graph NODES (a node = a group of unknowns that is never split: a single unknown, or all the
// unknowns of a cell).  Postorder, root last; a tree node lists the graph nodes its front eliminates.
struct EliminationTree {
  struct Node {
    std::vector<int> nodes;
    int child[2] = {-1, -1};
  };
  std::vector<Node> tree;
};

// Geometric nested dissection with VERTEX separators.  Graph nodes are bisected through their GROUP (a mesh cell:
// group_of_node[v], coordinates group_x/y; pass an empty group_of_node when every node is its own group): the cells
// of a region are split at the median coordinate, in x and in y, and the separator is the smaller of "nodes of one
// side that touch the other side", both sides tried.  With single unknowns as nodes the separator of the LDG density
// system is 6 unknowns per cell row (a cell and the facing nodes of its neighbour) instead of the 8 of two whole cells.
// Recursion stops at leaf_groups cells.  threads: the two halves of a region are dissected concurrently near the root
// (regions are disjoint; the tree does not depend on it).
EliminationTree nested_dissection(const std::vector<std::vector<int>>& adj, const std::vector<int>& group_of_node,
                                  const std::vector<double>& group_x, const std::vector<double>& group_y, int leaf_groups,
                                  int threads = 1);

// adjacency of the graph nodes induced by the (symmetrised) pattern of A
std::vector<std::vector<int>> node_adjacency(const CsrMatrix& A, const std::vector<int>& node_of_dof, int n_nodes,
                                             int threads = 1);

// Symbolic analysis + table layout for a given tree.
SolvePlan build_solve_plan(const CsrMatrix& A, const std::vector<int>& node_of_dof, int n_nodes,
                           const std::vector<std::vector<int>>& adj, const EliminationTree& tree, int threads = 1);

// convenience: node_adjacency + nested_dissection + build_solve_plan
SolvePlan build_solve_plan(const CsrMatrix& A, const std::vector<int>& node_of_dof, const std::vector<int>& group_of_node,
                           const std::vector<double>& group_x, const std::vector<double>& group_y, int leaf_groups,
                           int threads = 1);

// P A P^T (or its transpose) in CSR, P = plan.perm
CsrMatrix permute_csr(const CsrMatrix& A, const std::vector<int>& perm, bool transpose, int threads = 1);

// Host numeric factorisation: fills fwd (all G) and bwd (all [Inv | -H]) tables.  Throws StatusError(PECS_ERR_SINGULAR)
// when a pivot block cannot be inverted.
void factorize_host(const SolvePlan& plan, const CsrMatrix& A, std::vector<double>& fwd, std::vector<double>& bwd);

// (the host reference of the two solve sweeps lives in the test library, csrc/selftest/selftest.cpp)

} // namespace pecs
