// SchurReduction.cpp -- see SchurReduction.hpp.
#include "SchurReduction.hpp"

#include <cmath>
#include <cstdlib>

#include <algorithm>
#include <cmath>
#include <vector>

#include "../error.hpp"

namespace pecs {

namespace {

// dense n x n inverse (row-major, in place) with partial pivoting; false if singular
bool invert_small(int n, double* M) {
  std::vector<double> X((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) X[(size_t)i * n + i] = 1.0;
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int r = k + 1; r < n; ++r)
      if (std::fabs(M[(size_t)r * n + k]) > std::fabs(M[(size_t)p * n + k])) p = r;
    if (M[(size_t)p * n + k] == 0.0) return false;
    if (p != k)
      for (int j = 0; j < n; ++j) {
        std::swap(M[(size_t)k * n + j], M[(size_t)p * n + j]);
        std::swap(X[(size_t)k * n + j], X[(size_t)p * n + j]);
      }
    const double inv = 1.0 / M[(size_t)k * n + k];
    for (int j = 0; j < n; ++j) {
      M[(size_t)k * n + j] *= inv;
      X[(size_t)k * n + j] *= inv;
    }
    for (int r = 0; r < n; ++r) {
      if (r == k) continue;
      const double f = M[(size_t)r * n + k];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) {
        M[(size_t)r * n + j] -= f * M[(size_t)k * n + j];
        X[(size_t)r * n + j] -= f * X[(size_t)k * n + j];
      }
    }
  }
  std::copy(X.begin(), X.end(), M);
  return true;
}

// sparse row accumulator with a dense value array and a touched list
struct RowAccumulator {
  std::vector<double> val;
  std::vector<char> used;
  std::vector<int> touched;
  explicit RowAccumulator(int n) : val(n, 0.0), used(n, 0) {}
  void add(int j, double v) {
    if (!used[j]) {
      used[j] = 1;
      touched.push_back(j);
    }
    val[j] += v;
  }
  // appends the row (sorted by column; exact zeros and entries not above rel_drop * max|row| dropped) to a CSR under
  // construction and resets
  void flush(CsrMatrix& A, double rel_drop = 0.0) {
    std::sort(touched.begin(), touched.end());
    double cut = 0.0;
    if (rel_drop > 0.0) {
      for (int j : touched) cut = std::max(cut, std::fabs(val[j]));
      cut *= rel_drop;
    }
    for (int j : touched) {
      if (val[j] != 0.0 && std::fabs(val[j]) > cut) {
        A.col.push_back(j);
        A.val.push_back(val[j]);
      }
      val[j] = 0.0;
      used[j] = 0;
    }
    touched.clear();
    A.row_ptr.push_back((int)A.col.size());
  }
};

} // namespace

double schur_drop_tolerance() {
  if (const char* e = std::getenv("PECS_B200_SCHUR_DROP")) return std::atof(e);
  return 1e-13;
}

namespace {

// Rows [0, n_rows) of one or two CSR matrices, built by `threads` threads over contiguous row ranges and concatenated:
// row i is produced by fill(i, acc[0], acc[1]) into the thread's own accumulators, so every row is summed in the same
// order whatever the number of threads.
template <class Fill>
void build_rows(int n_rows, int threads, int n_out, const int n_cols[2], const double rel_drop[2], CsrMatrix* const out[2],
                Fill fill) {
  threads = std::max(1, std::min(threads, n_rows / 1024 + 1));
  std::vector<CsrMatrix> piece((size_t)2 * threads);
#pragma omp parallel for schedule(static, 1) num_threads(threads)
  for (int t = 0; t < threads; ++t) {
    const int lo = (int)((long long)n_rows * t / threads), hi = (int)((long long)n_rows * (t + 1) / threads);
    RowAccumulator acc0(n_cols[0]), acc1(n_out > 1 ? n_cols[1] : 0);
    CsrMatrix* p0 = &piece[(size_t)2 * t];
    CsrMatrix* p1 = &piece[(size_t)2 * t + 1];
    p0->row_ptr.assign(1, 0);
    p1->row_ptr.assign(1, 0);
    for (int i = lo; i < hi; ++i) {
      fill(i, acc0, acc1);
      acc0.flush(*p0, rel_drop[0]);
      if (n_out > 1) acc1.flush(*p1, rel_drop[1]);
    }
  }
  for (int m = 0; m < n_out; ++m) {
    CsrMatrix& M = *out[m];
    M.n = n_rows;
    M.row_ptr.assign(1, 0);
    M.row_ptr.reserve((size_t)n_rows + 1);
    std::vector<size_t> first(threads + 1, 0);
    for (int t = 0; t < threads; ++t) {
      const CsrMatrix& P = piece[(size_t)2 * t + m];
      for (size_t r = 1; r < P.row_ptr.size(); ++r) M.row_ptr.push_back((int)(first[t] + (size_t)P.row_ptr[r]));
      first[t + 1] = first[t] + P.col.size();
    }
    M.col.resize(first[threads]);
    M.val.resize(first[threads]);
#pragma omp parallel for schedule(static, 1) num_threads(threads)
    for (int t = 0; t < threads; ++t) {
      const CsrMatrix& P = piece[(size_t)2 * t + m];
      std::copy(P.col.begin(), P.col.end(), M.col.begin() + first[t]);
      std::copy(P.val.begin(), P.val.end(), M.val.begin() + first[t]);
    }
  }
}

} // namespace

bool build_schur_reduction(const CsrView& A, int n_cells, SchurReduction& out, int threads) {
  const int nq = 8 * n_cells, nu = 4 * n_cells;
  if (A.n != nq + nu) throw StatusError(PECS_ERR_INVALID, "build_schur_reduction: matrix size is not 12 * n_cells");
  threads = std::max(1, threads);
  // q unknown i (component i / 4n, node i % 4): its cell and its slot 0..7 inside the cell block
  auto q_cell = [&](int i) { return (i % (4 * n_cells)) / 4; };
  auto q_slot = [&](int i) { return 4 * (i / (4 * n_cells)) + (i % 4); };
  auto q_index = [&](int cell, int slot) { return (slot / 4) * 4 * n_cells + 4 * cell + (slot % 4); };

  // 1. cell blocks of A_qq; refuse if A_qq couples different cells
  std::vector<double> blocks((size_t)n_cells * 64, 0.0);
  bool couples = false;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(|| : couples)
  for (int i = 0; i < nq; ++i)
    for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
      const int j = A.col[k];
      if (j >= nq) continue;
      if (q_cell(j) != q_cell(i)) {
        if (A.val[k] != 0.0) couples = true;
        continue;
      }
      blocks[(size_t)q_cell(i) * 64 + 8 * q_slot(i) + q_slot(j)] = A.val[k];
    }
  if (couples) return false;
  bool singular = false;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(|| : singular)
  for (int c = 0; c < n_cells; ++c)
    if (!invert_small(8, &blocks[(size_t)c * 64])) singular = true;
  if (singular) throw StatusError(PECS_ERR_SINGULAR, "build_schur_reduction: singular current mass block");

  SchurReduction R;
  R.n_cells = n_cells;
  const double no_drop[2] = {0.0, 0.0};
  // 2. Ainv as CSR
  {
    const int cols[2] = {nq, 0};
    CsrMatrix* const dst[2] = {&R.Ainv, nullptr};
    build_rows(nq, threads, 1, cols, no_drop, dst, [&](int i, RowAccumulator& acc, RowAccumulator&) {
      const int c = q_cell(i), s = q_slot(i);
      for (int t = 0; t < 8; ++t) acc.add(q_index(c, t), blocks[(size_t)c * 64 + 8 * s + t]);
    });
  }
  // 3. T2 = Ainv * G_qu   (8n x 4n): row i = sum_t Ainv(i, t) * G_qu(row t of the same cell, :)
  {
    const int cols[2] = {nu, 0};
    CsrMatrix* const dst[2] = {&R.T2, nullptr};
    build_rows(nq, threads, 1, cols, no_drop, dst, [&](int i, RowAccumulator& acc, RowAccumulator&) {
      const int c = q_cell(i), s = q_slot(i);
      for (int t = 0; t < 8; ++t) {
        const double a = blocks[(size_t)c * 64 + 8 * s + t];
        if (a == 0.0) continue;
        const int r = q_index(c, t);
        for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; ++k)
          if (A.col[k] >= nq) acc.add(A.col[k] - nq, a * A.val[k]);
      }
    });
  }
  // 4. T1 = G_uq * Ainv (4n x 8n) and S = S_uu - G_uq * T2 (4n x 4n)
  {
    // The LDG fluxes make S compact: the couplings of a cell with the cells two faces away cancel exactly.  In
    // floating point the cancellation leaves entries of the order of 1e-17 of the row, far below the rounding error
    // of the large entries of the same product; they are removed so that they do not widen the separators
    // (there is nothing between 1e-16 and 1e-6 of the row maximum; PECS_B200_SCHUR_DROP overrides the threshold).
    const int cols[2] = {nq, nu};
    const double drop[2] = {0.0, schur_drop_tolerance()};
    CsrMatrix* const dst[2] = {&R.T1, &R.S};
    build_rows(nu, threads, 2, cols, drop, dst, [&](int i, RowAccumulator& acc1, RowAccumulator& accS) {
      const int r = nq + i;
      for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; ++k) {
        const int j = A.col[k];
        const double v = A.val[k];
        if (j >= nq) {
          accS.add(j - nq, v);
          continue;
        }
        const int c = q_cell(j), s = q_slot(j);
        for (int t = 0; t < 8; ++t) acc1.add(q_index(c, t), v * blocks[(size_t)c * 64 + 8 * s + t]);
        for (int kk = R.T2.row_ptr[j]; kk < R.T2.row_ptr[j + 1]; ++kk) accS.add(R.T2.col[kk], -v * R.T2.val[kk]);
      }
    });
  }
  out = std::move(R);
  return true;
}

} // namespace pecs
