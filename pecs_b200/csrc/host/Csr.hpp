// Csr.hpp -- minimal compressed-sparse-row matrix for the fixed system matrices
// (stands in for dealii::SparseMatrix<double> + SparsityPattern on the host side).
#pragma once
#include <cstdint>
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

// a matrix somebody else owns (the arrays handed over the C ABI)
struct CsrView {
  int n = 0;
  const int* row_ptr = nullptr;
  const int* col = nullptr;
  const double* val = nullptr;
  CsrView() = default;
  CsrView(int n_, const int* r, const int* c, const double* v) : n(n_), row_ptr(r), col(c), val(v) {}
  CsrView(const CsrMatrix& A) : n(A.n), row_ptr(A.row_ptr.data()), col(A.col.data()), val(A.val.data()) {}
};

// ELL layout of a fixed matrix as the device kernels stream it (cuda/schur_kernels.cuh): slot-major, one thread per row,
// unit stride across the rows; block == 4: a slot is one aligned group of four columns (9 B per stored entry instead of
// 12).  Built on the host (host/EllTable.cpp) -- at setup by the preparation threads -- and uploaded as it is.
struct HostEll {
  int n = 0, width = 0, block = 1;
  std::vector<int> col;    // [width][n]
  std::vector<double> val; // [width * block][n]
};
// row_order (optional): ELL row i holds row (*row_order)[i] of A.  PECS_B200_ELL_SCALAR forces block = 1.
HostEll build_ell(const CsrMatrix& A, const std::vector<int>* row_order = nullptr, int threads = 1);

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

} // namespace pecs
