// Csr.hpp -- minimal compressed-sparse-row matrix for the fixed system matrices
// (stands in for dealii::SparseMatrix<double> + SparsityPattern on the host side).
#pragma once
#include <omp.h>

#include <algorithm>
#include <cstddef>
#include <memory>
#include <numeric>
#include <vector>

namespace pecs {

struct CsrMatrix {
  int n = 0;
  std::vector<int> row_ptr; // n+1
  std::vector<int> col;
  std::vector<double> val;
  size_t nnz() const { return col.size(); }

  void vmult(double* y, const double* x) const {
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) s += val[k] * x[col[k]];
      y[i] = s;
    }
  }
};

// Triplet accumulator; duplicates are summed in insertion order when compressed.
class TripletList {
public:
  explicit TripletList(int n) : n_(n) {}
  void add(int i, int j, double v) {
    r_.push_back(i);
    c_.push_back(j);
    v_.push_back(v);
  }
  void reserve(size_t m) {
    r_.reserve(m);
    c_.reserve(m);
    v_.reserve(m);
  }
  CsrMatrix compress(bool drop_zeros = false) const {
    CsrMatrix A;
    A.n = n_;
    std::vector<int> cnt(n_ + 1, 0);
    for (int i : r_) ++cnt[i + 1];
    std::partial_sum(cnt.begin(), cnt.end(), cnt.begin());
    std::vector<int> pos(cnt.begin(), cnt.end() - 1), tc(r_.size());
    std::vector<double> tv(r_.size());
    for (size_t k = 0; k < r_.size(); ++k) { // stable bucket by row
      const int p = pos[r_[k]]++;
      tc[p] = c_[k];
      tv[p] = v_[k];
    }
    A.row_ptr.assign(n_ + 1, 0);
    std::vector<int> order;
    for (int i = 0; i < n_; ++i) {
      const int b = cnt[i], e = cnt[i + 1];
      order.resize(e - b);
      std::iota(order.begin(), order.end(), b);
      std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return tc[a] < tc[c]; });
      size_t k = 0;
      while (k < order.size()) {
        const int cj = tc[order[k]];
        double s = 0;
        while (k < order.size() && tc[order[k]] == cj) s += tv[order[k++]];
        if (drop_zeros && s == 0.0) continue;
        A.col.push_back(cj);
        A.val.push_back(s);
      }
      A.row_ptr[i + 1] = (int)A.col.size();
    }
    return A;
  }

private:
  int n_;
  std::vector<int> r_, c_;
  std::vector<double> v_;
};

// Triplets of TWO matrices with a common pattern (the two carriers of a pair differ only in the mobility that scales
// the current-current mass blocks).  Filled in chunks -- one chunk per thread over a contiguous range of cells -- and
// compressed once for both: rows are bucketed in chunk order, so duplicates are summed in the order a sequential
// assembly would have inserted them, and the per-row sorts run in parallel.
class PairedTripletChunks {
public:
  struct Chunk {
    std::vector<int> r, c;
    std::vector<double> v1, v2;
    void add(int i, int j, double a, double b) {
      r.push_back(i);
      c.push_back(j);
      v1.push_back(a);
      v2.push_back(b);
    }
    void reserve(size_t m) {
      r.reserve(m);
      c.reserve(m);
      v1.reserve(m);
      v2.reserve(m);
    }
  };
  PairedTripletChunks(int n, int n_chunks) : n_(n), chunks_(n_chunks) {}
  Chunk& chunk(int k) { return chunks_[k]; }
  int n_chunks() const { return (int)chunks_.size(); }

  // entries whose two sums are both exactly zero are dropped when drop_zeros is set
  void compress(bool drop_zeros, CsrMatrix& A1, CsrMatrix& A2) const {
    const int n = n_;
    // stable bucket by row.  Every thread owns a contiguous range of ROWS and scans all chunks in order for the entries
    // of its rows (the scans are sequential reads; the scattered writes stay inside the thread's own segment).
    std::vector<size_t> start(n + 1, 0);
    int n_parts = 1;
#pragma omp parallel
    {
#pragma omp single
      n_parts = omp_get_num_threads();
    }
    auto row_lo = [&](int t) { return (int)((long long)n * t / n_parts); };
#pragma omp parallel for schedule(static, 1) num_threads(n_parts)
    for (int t = 0; t < n_parts; ++t) {
      const int lo = row_lo(t), hi = row_lo(t + 1);
      for (const Chunk& ch : chunks_)
        for (int i : ch.r)
          if (i >= lo && i < hi) ++start[i + 1];
    }
    std::partial_sum(start.begin(), start.end(), start.begin());
    const size_t total = start[n];
    std::unique_ptr<int[]> tc(new int[total]);
    std::unique_ptr<double[]> t1(new double[total]), t2(new double[total]);
#pragma omp parallel for schedule(static, 1) num_threads(n_parts)
    for (int t = 0; t < n_parts; ++t) {
      const int lo = row_lo(t), hi = row_lo(t + 1);
      std::vector<size_t> pos(start.begin() + lo, start.begin() + hi);
      for (const Chunk& ch : chunks_)
        for (size_t k = 0; k < ch.r.size(); ++k) {
          const int i = ch.r[k];
          if (i < lo || i >= hi) continue;
          const size_t q = pos[i - lo]++;
          tc[q] = ch.c[k];
          t1[q] = ch.v1[k];
          t2[q] = ch.v2[k];
        }
    }
    // per row: stable sort by column, sum duplicates in place at the head of the row's segment
    std::vector<int> kept(n, 0);
#pragma omp parallel
    {
      std::vector<int> order, cc;
      std::vector<double> c1, c2;
#pragma omp for schedule(dynamic, 4096)
      for (int i = 0; i < n; ++i) {
        const size_t b = start[i], e = start[i + 1];
        const int m = (int)(e - b);
        order.resize(m);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return tc[b + x] < tc[b + y]; });
        cc.clear();
        c1.clear();
        c2.clear();
        int k = 0;
        while (k < m) {
          const int cj = tc[b + order[k]];
          double s1 = 0, s2 = 0;
          while (k < m && tc[b + order[k]] == cj) {
            s1 += t1[b + order[k]];
            s2 += t2[b + order[k]];
            ++k;
          }
          if (drop_zeros && s1 == 0.0 && s2 == 0.0) continue;
          cc.push_back(cj);
          c1.push_back(s1);
          c2.push_back(s2);
        }
        std::copy(cc.begin(), cc.end(), tc.get() + b);
        std::copy(c1.begin(), c1.end(), t1.get() + b);
        std::copy(c2.begin(), c2.end(), t2.get() + b);
        kept[i] = (int)cc.size();
      }
    }
    A1.n = A2.n = n;
    A1.row_ptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) A1.row_ptr[i + 1] = A1.row_ptr[i] + kept[i];
    A2.row_ptr = A1.row_ptr;
    const size_t nnz = (size_t)A1.row_ptr[n];
    A1.col.resize(nnz);
    A1.val.resize(nnz);
    A2.val.resize(nnz);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      const size_t b = start[i], o = (size_t)A1.row_ptr[i];
      for (int k = 0; k < kept[i]; ++k) {
        A1.col[o + k] = tc[b + k];
        A1.val[o + k] = t1[b + k];
        A2.val[o + k] = t2[b + k];
      }
    }
    A2.col = A1.col;
  }

private:
  int n_;
  std::vector<Chunk> chunks_;
};

} // namespace pecs
