// SparseDirect.cpp -- see SparseDirect.hpp.
#include "SparseDirect.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <future>
#include <numeric>
#include <stdexcept>

#include "../error.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace pecs {

namespace {

int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct NdBuilder {
  typedef std::vector<EliminationTree::Node> Tree; // postorder, root last, child indices local to the vector
  const std::vector<std::vector<int>>& adj;
  const std::vector<int>& group_of; // empty: every node is its own group
  const std::vector<double>&x, &y;  // per group
  int leaf_groups;
  int parallel_depth;      // the two halves of a region are dissected concurrently down to this depth
  std::vector<int> side;   // per node: 0 = outside the current region, 1 = A, 2 = B
  std::vector<int> gside;  // per group, scratch
  std::vector<int> gstamp; // per group: last region that counted it
  std::atomic<int> stamp{0};
  // Concurrent regions are disjoint in nodes AND in groups (a region is bisected through whole groups), so the three
  // scratch arrays are shared without locks; only the region counter is atomic.

  int group(int v) const { return group_of.empty() ? v : group_of[v]; }

  static Tree leaf(std::vector<int>& nodes) {
    Tree t(1);
    std::sort(nodes.begin(), nodes.end());
    t[0].nodes.swap(nodes);
    return t;
  }

  Tree build(std::vector<int>& nodes, int depth) {
    std::vector<int> groups;
    const int my_stamp = ++stamp;
    for (int v : nodes) {
      const int g = group(v);
      if (gstamp[g] != my_stamp) {
        gstamp[g] = my_stamp;
        groups.push_back(g);
      }
    }
    if ((int)groups.size() <= leaf_groups) return leaf(nodes);
    std::vector<int> bestA, bestB, bestS;
    bool have = false;
    long long best_imbalance = 0;
    for (int dir = 0; dir < 2; ++dir) {
      const std::vector<double>& c = dir == 0 ? x : y;
      std::vector<int> sorted = groups;
      const size_t half = sorted.size() / 2;
      std::nth_element(sorted.begin(), sorted.begin() + half, sorted.end(),
                       [&](int a, int b) { return c[a] < c[b] || (c[a] == c[b] && a < b); });
      for (size_t k = 0; k < sorted.size(); ++k) gside[sorted[k]] = k < half ? 1 : 2;
      for (int v : nodes) side[v] = gside[group(v)];
      for (int from = 1; from <= 2; ++from) {
        // separator = nodes of side `from` that touch the other side
        std::vector<int> A, B, S;
        for (int v : nodes) {
          if (side[v] != from) {
            (side[v] == 1 ? A : B).push_back(v);
            continue;
          }
          bool touches = false;
          for (int w : adj[v])
            if (side[w] == 3 - from) {
              touches = true;
              break;
            }
          (touches ? S : (from == 1 ? A : B)).push_back(v);
        }
        if (A.empty() || B.empty() || S.empty()) continue;
        const long long imbalance = std::llabs((long long)A.size() - (long long)B.size());
        if (!have || S.size() < bestS.size() || (S.size() == bestS.size() && imbalance < best_imbalance)) {
          bestA.swap(A);
          bestB.swap(B);
          bestS.swap(S);
          best_imbalance = imbalance;
          have = true;
        }
      }
      for (int v : nodes) side[v] = 0;
    }
    if (!have) return leaf(nodes);
    nodes.clear();
    nodes.shrink_to_fit();
    Tree a, b;
    if (depth < parallel_depth) {
      std::future<Tree> other = std::async(std::launch::async, [&] { return build(bestB, depth + 1); });
      a = build(bestA, depth + 1);
      b = other.get();
    } else {
      a = build(bestA, depth + 1);
      b = build(bestB, depth + 1);
    }
    // postorder: the first half's subtree, the second half's, then the separator
    const int na = (int)a.size(), nb = (int)b.size();
    a.reserve((size_t)na + nb + 1);
    for (EliminationTree::Node& t : b) {
      for (int k = 0; k < 2; ++k)
        if (t.child[k] >= 0) t.child[k] += na;
      a.push_back(std::move(t));
    }
    EliminationTree::Node t;
    t.child[0] = na - 1;
    t.child[1] = na + nb - 1;
    std::sort(bestS.begin(), bestS.end());
    t.nodes.swap(bestS);
    a.push_back(std::move(t));
    return a;
  }
};

int64_t target_panels_per_level() {
  static const int64_t v = [] {
    const char* e = std::getenv("PECS_B200_TARGET_PANELS");
    return e && std::atoll(e) > 0 ? (int64_t)std::atoll(e) : (int64_t)kTargetPanelsPerLevel;
  }();
  return v;
}

int64_t env_or(const char* name, int64_t fallback) {
  const char* e = std::getenv(name);
  return e ? (int64_t)std::atoll(e) : fallback;
}
int64_t small_table_doubles() {
  static const int64_t v = env_or("PECS_B200_SMALL_TABLE", kSmallTableDoubles);
  return v;
}
int64_t panel_target_doubles() {
  static const int64_t v = std::max<int64_t>(32, env_or("PECS_B200_PANEL_DOUBLES", kPanelTargetDoubles));
  return v;
}
int pow2_ceil_log2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// Shape of one front operator.  log2_cap: level-wide cap on the panel height (keeps enough panels per level).
void shape_table(PanelTable& t, int rows, int cols, int log2_cap) {
  t = PanelTable{};
  if (rows == 0 || cols == 0) return;
  t.rows = rows;
  t.cols = cols;
  if (rows <= 64) {
    // candidate for "one warp per front": one or two panels as tall as the table, no shuffles when P = 32
    const int l = rows <= 32 ? std::max(2, pow2_ceil_log2(rows)) : 5;
    const int rows_pad = round_up(rows, 1 << l), cols_pad = round_up(cols, 32 >> l);
    if ((int64_t)rows_pad * cols_pad <= small_table_doubles()) {
      t.small = 1;
      t.log2P = l;
      t.rows_pad = rows_pad;
      t.cols_pad = cols_pad;
      return;
    }
  }
  // thread-block mode: panels of about panel_target_doubles entries, so that the warps of a level carry equal loads
  int l = 0;
  while (l < 5 && l < log2_cap && ((int64_t)cols << (l + 1)) <= panel_target_doubles()) ++l;
  // not more than 1/16 of padding rows (panels of up to 4 rows are always fine)
  while (l > 2 && (round_up(rows, 1 << l) - rows) * 16 > rows) --l;
  // a front is streamed by the kWarpsPerFront warps of one thread block: give every warp a panel when the front allows it
  while (l > 0 && (round_up(rows, 1 << l) >> l) < kWarpsPerFront) --l;
  t.log2P = l;
  t.rows_pad = round_up(rows, 1 << l);
  t.cols_pad = round_up(cols, 32 >> l);
}

} // namespace

std::vector<std::vector<int>> node_adjacency(const CsrMatrix& A, const std::vector<int>& node_of_dof, int n_nodes,
                                             int threads) {
  std::vector<std::vector<int>> adj(n_nodes);
  {
    std::vector<int> degree(n_nodes, 0); // with duplicates: an upper bound, reserved once
    for (int i = 0; i < A.n; ++i)
      for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
        const int a = node_of_dof[i], b = node_of_dof[A.col[k]];
        if (a != b) {
          ++degree[a];
          ++degree[b];
        }
      }
#pragma omp parallel for schedule(static) num_threads(std::max(1, threads))
    for (int v = 0; v < n_nodes; ++v) adj[v].reserve((size_t)degree[v]);
  }
  for (int i = 0; i < A.n; ++i)
    for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
      const int a = node_of_dof[i], b = node_of_dof[A.col[k]];
      if (a != b) {
        adj[a].push_back(b);
        adj[b].push_back(a);
      }
    }
#pragma omp parallel for schedule(static) num_threads(std::max(1, threads))
  for (int v = 0; v < n_nodes; ++v) {
    std::vector<int>& a = adj[v];
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  return adj;
}

EliminationTree nested_dissection(const std::vector<std::vector<int>>& adj, const std::vector<int>& group_of_node,
                                  const std::vector<double>& group_x, const std::vector<double>& group_y, int leaf_groups,
                                  int threads) {
  const int n_nodes = (int)adj.size();
  const int n_groups = (int)group_x.size();
  if (!group_of_node.empty() && (int)group_of_node.size() != n_nodes)
    throw StatusError(PECS_ERR_INVALID, "nested_dissection: group_of_node size");
  if (group_of_node.empty() && n_groups != n_nodes) throw StatusError(PECS_ERR_INVALID, "nested_dissection: coordinates size");
  int depth = 0;
  while ((1 << (depth + 1)) <= threads) ++depth; // 2^depth concurrent subtrees
  NdBuilder nd{adj, group_of_node, group_x, group_y, std::max(1, leaf_groups), depth, std::vector<int>(n_nodes, 0),
               std::vector<int>(n_groups, 0), std::vector<int>(n_groups, 0)};
  std::vector<int> all(n_nodes);
  std::iota(all.begin(), all.end(), 0);
  EliminationTree out;
  out.tree = nd.build(all, 0);
  return out;
}

SolvePlan build_solve_plan(const CsrMatrix& A, const std::vector<int>& node_of_dof, const std::vector<int>& group_of_node,
                           const std::vector<double>& group_x, const std::vector<double>& group_y, int leaf_groups,
                           int threads) {
  if ((int)node_of_dof.size() != A.n) throw StatusError(PECS_ERR_INVALID, "build_solve_plan: node_of_dof size");
  const int n_nodes = group_of_node.empty() ? (int)group_x.size() : (int)group_of_node.size();
  const std::vector<std::vector<int>> adj = node_adjacency(A, node_of_dof, n_nodes, threads);
  const EliminationTree tree = nested_dissection(adj, group_of_node, group_x, group_y, leaf_groups, threads);
  return build_solve_plan(A, node_of_dof, n_nodes, adj, tree, threads);
}

SolvePlan build_solve_plan(const CsrMatrix& A, const std::vector<int>& node_of_dof, int n_nodes,
                           const std::vector<std::vector<int>>& adj, const EliminationTree& etree, int threads) {
  threads = std::max(1, threads);
  const int n = A.n;
  if ((int)node_of_dof.size() != n) throw StatusError(PECS_ERR_INVALID, "build_solve_plan: node_of_dof size");
  std::vector<std::vector<int>> node_dofs(n_nodes);
  for (int i = 0; i < n; ++i) node_dofs[node_of_dof[i]].push_back(i);
  const std::vector<EliminationTree::Node>& tree = etree.tree;
  const int nf = (int)tree.size();

  SolvePlan plan;
  plan.n = n;
  plan.fronts.resize(nf);
  // node positions in elimination order + dof permutation
  std::vector<int> node_pos(n_nodes, -1), node_first_dofpos(n_nodes, -1);
  plan.perm.assign(n, -1);
  plan.iperm.assign(n, -1);
  int next_node = 0, next_dof = 0;
  std::vector<int> last_node_pos(nf);
  for (int f = 0; f < nf; ++f) {
    Front& F = plan.fronts[f];
    F.p0 = next_dof;
    for (int v : tree[f].nodes) {
      if (node_pos[v] >= 0) throw StatusError(PECS_ERR_INTERNAL, "build_solve_plan: a node is eliminated twice");
      node_pos[v] = next_node++;
      node_first_dofpos[v] = next_dof;
      for (int d : node_dofs[v]) {
        plan.perm[d] = next_dof;
        plan.iperm[next_dof] = d;
        ++next_dof;
      }
    }
    F.np = next_dof - F.p0;
    last_node_pos[f] = next_node - 1;
    for (int k = 0; k < 2; ++k) {
      F.child[k] = tree[f].child[k];
      if (F.child[k] >= 0) plan.fronts[F.child[k]].parent = f;
    }
  }
  if (next_dof != n) throw StatusError(PECS_ERR_INTERNAL, "build_solve_plan: not every unknown was ordered");
  // depth (root is the last front in postorder)
  for (int f = nf - 1; f >= 0; --f) plan.fronts[f].depth = plan.fronts[f].parent < 0 ? 0 : plan.fronts[plan.fronts[f].parent].depth + 1;

  // symbolic structure on nodes: boundary nodes as their node positions, ascending
  std::vector<int> pos_to_node(n_nodes, -1);
  for (int v = 0; v < n_nodes; ++v)
    if (node_pos[v] >= 0) pos_to_node[node_pos[v]] = v;
  std::vector<std::vector<int>> bd_nodes(nf);
  int max_depth = 0;
  for (const Front& F : plan.fronts) max_depth = std::max(max_depth, F.depth);
  plan.levels.assign(max_depth + 1, {});
  for (int f = 0; f < nf; ++f) plan.levels[plan.fronts[f].depth].push_back(f);
  // children before parents: level by level from the bottom, the fronts of a level side by side
  for (int depth = max_depth; depth >= 0; --depth) {
    const std::vector<int>& lvl = plan.levels[depth];
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads) if (lvl.size() > 64)
    for (int q = 0; q < (int)lvl.size(); ++q) {
      const int f = lvl[q];
      std::vector<int> s;
      for (int v : tree[f].nodes)
        for (int w : adj[v])
          if (node_pos[w] > last_node_pos[f]) s.push_back(node_pos[w]);
      for (int k = 0; k < 2; ++k)
        if (plan.fronts[f].child[k] >= 0) {
          for (int p : bd_nodes[plan.fronts[f].child[k]])
            if (p > last_node_pos[f]) s.push_back(p);
        }
      std::sort(s.begin(), s.end());
      s.erase(std::unique(s.begin(), s.end()), s.end());
      bd_nodes[f].swap(s);
    }
  }
  // expand to unknowns
  for (int f = 0; f < nf; ++f) {
    Front& F = plan.fronts[f];
    F.bd_off = (int64_t)plan.bd_index.size();
    for (int p : bd_nodes[f]) {
      const int v = pos_to_node[p];
      for (int k = 0; k < (int)node_dofs[v].size(); ++k) plan.bd_index.push_back(node_first_dofpos[v] + k);
    }
    F.nb = (int)((int64_t)plan.bd_index.size() - F.bd_off);
    std::vector<int>().swap(bd_nodes[f]);
  }
  // panel heights: as tall as possible, capped per level so that the level offers enough panels for the whole GPU
  for (const std::vector<int>& lvl : plan.levels)
    for (int which = 0; which < 2; ++which) {
      int cap = 5;
      for (; cap > 0; --cap) {
        int64_t panels = 0;
        for (int f : lvl) {
          const Front& F = plan.fronts[f];
          PanelTable t;
          shape_table(t, which == 0 ? F.nb : F.np, which == 0 ? F.np : F.np + F.nb, cap);
          panels += t.n_panels();
        }
        if (panels >= target_panels_per_level()) break;
      }
      for (int f : lvl) {
        Front& F = plan.fronts[f];
        if (which == 0)
          shape_table(F.fwd, F.nb, F.np, cap);
        else
          shape_table(F.bwd, F.np, F.np + F.nb, cap);
      }
    }
  // lay out tables and child buffers
  for (int f = 0; f < nf; ++f) {
    Front& F = plan.fronts[f];
    F.fwd.off = plan.fwd_entries;
    plan.fwd_entries += F.fwd.size();
    F.bwd.off = plan.bwd_entries;
    plan.bwd_entries += F.bwd.size();
    for (int k = 0; k < 2; ++k)
      if (F.child[k] >= 0) {
        plan.fronts[F.child[k]].which_child = k;
        F.cbuf_off[k] = plan.upd_entries;
        plan.upd_entries += F.np + F.nb;
        plan.upd_entries += plan.upd_entries & 1;
      }
    plan.max_np = std::max(plan.max_np, F.np);
    plan.max_nb = std::max(plan.max_nb, F.nb);
  }
  // where every boundary unknown of a front lives in its parent's local numbering [pivots | boundary]
  plan.out_map.assign(plan.bd_index.size(), -1);
  bool outside = false;
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads) reduction(|| : outside)
  for (int f = 0; f < nf; ++f) {
    const Front& C = plan.fronts[f];
    if (C.parent < 0) continue;
    const Front& F = plan.fronts[C.parent];
    const int* cbd = plan.bd(C);
    const int* fbd = plan.bd(F);
    for (int s = 0; s < C.nb; ++s) {
      const int pos = cbd[s];
      int l;
      if (pos < F.p0 + F.np) {
        if (pos < F.p0) outside = true; // below the parent's pivots
        l = pos - F.p0;
      } else {
        const int* it = std::lower_bound(fbd, fbd + F.nb, pos);
        if (it == fbd + F.nb || *it != pos) outside = true;
        l = F.np + (int)(it - fbd);
      }
      plan.out_map[C.bd_off + s] = l;
    }
  }
  if (outside) throw StatusError(PECS_ERR_INTERNAL, "build_solve_plan: child boundary not contained in parent front");
  return plan;
}

// ---------------------------------------------------------------------------------------------- host numeric
namespace {

// C (m x n) -= / = A (m x k) * B (k x n), all row-major with leading dimensions
void gemm_sub(int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc, bool par) {
#pragma omp parallel for schedule(static) if (par)
  for (int i = 0; i < m; ++i) {
    double* c = C + (size_t)i * ldc;
    for (int p = 0; p < k; ++p) {
      const double a = A[(size_t)i * lda + p];
      if (a == 0.0) continue;
      const double* b = B + (size_t)p * ldb;
      for (int j = 0; j < n; ++j) c[j] -= a * b[j];
    }
  }
}

// in-place inverse of an n x n row-major matrix by LU with partial pivoting; returns false if singular
bool invert(int n, double* M, bool par) {
  std::vector<int> piv(n);
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = std::fabs(M[(size_t)k * n + k]);
    for (int r = k + 1; r < n; ++r)
      if (std::fabs(M[(size_t)r * n + k]) > best) {
        best = std::fabs(M[(size_t)r * n + k]);
        p = r;
      }
    if (best == 0.0 || !std::isfinite(best)) return false;
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < n; ++j) std::swap(M[(size_t)k * n + j], M[(size_t)p * n + j]);
    const double inv = 1.0 / M[(size_t)k * n + k];
#pragma omp parallel for schedule(static) if (par && n - k > 256)
    for (int r = k + 1; r < n; ++r) {
      const double l = M[(size_t)r * n + k] * inv;
      M[(size_t)r * n + k] = l;
      if (l == 0.0) continue;
      const double* rk = M + (size_t)k * n;
      double* rr = M + (size_t)r * n;
      for (int j = k + 1; j < n; ++j) rr[j] -= l * rk[j];
    }
  }
  // inverse from the factors: solve L U X = P I column block-wise (row-major friendly formulation)
  std::vector<double> X((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) X[(size_t)i * n + i] = 1.0;
  for (int k = 0; k < n; ++k)
    if (piv[k] != k)
      for (int j = 0; j < n; ++j) std::swap(X[(size_t)k * n + j], X[(size_t)piv[k] * n + j]);
  // forward: rows of X updated in order (L unit lower)
  for (int i = 1; i < n; ++i) {
    double* xi = X.data() + (size_t)i * n;
    for (int k = 0; k < i; ++k) {
      const double l = M[(size_t)i * n + k];
      if (l == 0.0) continue;
      const double* xk = X.data() + (size_t)k * n;
      for (int j = 0; j < n; ++j) xi[j] -= l * xk[j];
    }
  }
  // backward
  for (int i = n - 1; i >= 0; --i) {
    double* xi = X.data() + (size_t)i * n;
    for (int k = i + 1; k < n; ++k) {
      const double u = M[(size_t)i * n + k];
      if (u == 0.0) continue;
      const double* xk = X.data() + (size_t)k * n;
      for (int j = 0; j < n; ++j) xi[j] -= u * xk[j];
    }
    const double d = 1.0 / M[(size_t)i * n + i];
    for (int j = 0; j < n; ++j) xi[j] *= d;
  }
  std::copy(X.begin(), X.end(), M);
  return true;
}

} // namespace

CsrMatrix permute_csr(const CsrMatrix& A, const std::vector<int>& perm, bool transpose, int threads) {
  // entry (i, j) goes to row perm[i] (perm[j] when transposed); rows are filled by counting, then sorted by column
  // one by one (a matrix has no duplicate entries, so the order inside a row is all there is to fix)
  const int n = A.n;
  threads = std::max(1, threads);
  CsrMatrix B;
  B.n = n;
  B.row_ptr.assign((size_t)n + 1, 0);
  if (transpose) {
    for (size_t k = 0; k < A.col.size(); ++k) ++B.row_ptr[(size_t)perm[A.col[k]] + 1];
  } else {
    for (int i = 0; i < n; ++i) B.row_ptr[(size_t)perm[i] + 1] = A.row_ptr[i + 1] - A.row_ptr[i];
  }
  for (int i = 0; i < n; ++i) B.row_ptr[i + 1] += B.row_ptr[i];
  B.col.resize(A.col.size());
  B.val.resize(A.val.size());
  if (transpose) {
    std::vector<int> next(B.row_ptr.begin(), B.row_ptr.end() - 1);
    for (int i = 0; i < n; ++i)
      for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
        const int q = next[perm[A.col[k]]]++;
        B.col[q] = perm[i];
        B.val[q] = A.val[k];
      }
  } else {
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < n; ++i) {
      int q = B.row_ptr[perm[i]];
      for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k, ++q) {
        B.col[q] = perm[A.col[k]];
        B.val[q] = A.val[k];
      }
    }
  }
#pragma omp parallel num_threads(threads)
  {
    std::vector<std::pair<int, double>> row;
#pragma omp for schedule(static)
    for (int i = 0; i < n; ++i) {
      const int b = B.row_ptr[i], e = B.row_ptr[i + 1];
      row.resize((size_t)(e - b));
      for (int k = b; k < e; ++k) row[(size_t)(k - b)] = {B.col[k], B.val[k]};
      std::sort(row.begin(), row.end(), [](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
      for (int k = b; k < e; ++k) {
        B.col[k] = row[(size_t)(k - b)].first;
        B.val[k] = row[(size_t)(k - b)].second;
      }
    }
  }
  return B;
}

void factorize_host(const SolvePlan& plan, const CsrMatrix& A, std::vector<double>& fwd, std::vector<double>& bwd) {
  const CsrMatrix Ap = permute_csr(A, plan.perm, false), Apt = permute_csr(A, plan.perm, true);
  fwd.assign((size_t)plan.fwd_entries, 0.0);
  bwd.assign((size_t)plan.bwd_entries, 0.0);
  const int nf = (int)plan.fronts.size();
  std::vector<std::vector<double>> update(nf); // Schur complements waiting for their parent
  bool singular = false;

  auto process = [&](int f, bool inner_parallel) {
    const Front& F = plan.fronts[f];
    const int np = F.np, nb = F.nb, m = np + nb;
    const int* bd = plan.bd(F);
    std::vector<double> M((size_t)m * m, 0.0);
    auto local = [&](int pos) -> int {
      if (pos < F.p0 + np) return pos - F.p0;
      const int* it = std::lower_bound(bd, bd + nb, pos);
      return (it != bd + nb && *it == pos) ? np + (int)(it - bd) : -1;
    };
    for (int i = 0; i < np; ++i) {
      const int r = F.p0 + i;
      for (int k = Ap.row_ptr[r]; k < Ap.row_ptr[r + 1]; ++k) {
        const int q = Ap.col[k];
        if (q < F.p0) continue;
        const int l = local(q);
        if (l < 0) throw StatusError(PECS_ERR_INTERNAL, "factorize_host: entry outside the symbolic front");
        M[(size_t)i * m + l] = Ap.val[k];
      }
      for (int k = Apt.row_ptr[r]; k < Apt.row_ptr[r + 1]; ++k) {
        const int q = Apt.col[k];
        if (q < F.p0 + np) continue;
        const int l = local(q);
        if (l < 0) throw StatusError(PECS_ERR_INTERNAL, "factorize_host: entry outside the symbolic front");
        M[(size_t)l * m + i] = Apt.val[k];
      }
    }
    for (int c = 0; c < 2; ++c) {
      if (F.child[c] < 0) continue;
      const Front& C = plan.fronts[F.child[c]];
      std::vector<double>& U = update[F.child[c]];
      std::vector<int> map(C.nb);
      const int* cbd = plan.bd(C);
      for (int s = 0; s < C.nb; ++s) map[s] = local(cbd[s]);
      for (int s = 0; s < C.nb; ++s) {
        double* row = M.data() + (size_t)map[s] * m;
        const double* u = U.data() + (size_t)s * C.nb;
        for (int t = 0; t < C.nb; ++t) row[map[t]] += u[t];
      }
      std::vector<double>().swap(U);
    }
    // split the frontal matrix
    std::vector<double> Fpp((size_t)np * np), Fpb((size_t)np * nb), Fbp((size_t)nb * np);
    for (int i = 0; i < np; ++i) {
      std::copy(M.begin() + (size_t)i * m, M.begin() + (size_t)i * m + np, Fpp.begin() + (size_t)i * np);
      std::copy(M.begin() + (size_t)i * m + np, M.begin() + (size_t)(i + 1) * m, Fpb.begin() + (size_t)i * nb);
    }
    for (int i = 0; i < nb; ++i)
      std::copy(M.begin() + (size_t)(np + i) * m, M.begin() + (size_t)(np + i) * m + np, Fbp.begin() + (size_t)i * np);
    if (!invert(np, Fpp.data(), inner_parallel)) {
      singular = true;
      return;
    }
    // G = Fbp * Inv (stored positively), H = Inv * Fpb (stored negated next to Inv), U = Fbb - G * Fpb
    std::vector<double> negG((size_t)nb * np, 0.0), G((size_t)nb * np);
    gemm_sub(nb, np, np, Fbp.data(), np, Fpp.data(), np, negG.data(), np, inner_parallel);
    for (size_t k = 0; k < negG.size(); ++k) G[k] = -negG[k];
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j < np; ++j) fwd[(size_t)F.fwd.index(i, j)] = G[(size_t)i * np + j];
    std::vector<double> negH((size_t)np * nb, 0.0);
    gemm_sub(np, nb, np, Fpp.data(), np, Fpb.data(), nb, negH.data(), nb, inner_parallel); // -H = -Inv F_PB
    for (int i = 0; i < np; ++i) {
      for (int j = 0; j < np; ++j) bwd[(size_t)F.bwd.index(i, j)] = Fpp[(size_t)i * np + j];
      for (int j = 0; j < nb; ++j) bwd[(size_t)F.bwd.index(i, np + j)] = negH[(size_t)i * nb + j];
    }
    if (F.parent >= 0) {
      std::vector<double>& U = update[f];
      U.resize((size_t)nb * nb);
      for (int i = 0; i < nb; ++i)
        std::copy(M.begin() + (size_t)(np + i) * m + np, M.begin() + (size_t)(np + i + 1) * m, U.begin() + (size_t)i * nb);
      gemm_sub(nb, nb, np, G.data(), np, Fpb.data(), nb, U.data(), nb, inner_parallel);
    }
  };

  int threads = 1;
#ifdef _OPENMP
  threads = omp_get_max_threads();
#endif
  for (int d = (int)plan.levels.size() - 1; d >= 0; --d) {
    const std::vector<int>& lvl = plan.levels[d];
    if ((int)lvl.size() >= 2 * threads) {
#pragma omp parallel for schedule(dynamic, 1)
      for (int k = 0; k < (int)lvl.size(); ++k) process(lvl[k], false);
    } else {
      for (int f : lvl) process(f, true);
    }
    if (singular) throw StatusError(PECS_ERR_SINGULAR, "factorize_host: singular pivot block");
  }
}

} // namespace pecs
