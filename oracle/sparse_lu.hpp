// oracle/sparse_lu.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// Stand-in for dealii::SparseDirectUMFPACK (reference include/Carrier.hpp:88, include/Poisson.hpp:74):
//   initialize(A)  = factor once        (reference source/Carrier.cpp:26-32, source/Poisson.cpp:92-96)
//   vmult(x, b)    = forward/back substitution with the stored factors (Carrier.cpp:34-40, Poisson.cpp:98-105)
// UMFPACK 5.x is a third-party dependency bundled with deal.II (version not pinned by the reference: deal.II >= 8.3)
// and is absent from this container.  Its published algorithm is: fill-reducing column pre-ordering, then sparse
// LU with threshold partial pivoting (threshold 0.1, preferring the diagonal on structurally symmetric matrices).
// This file restates that as a left-looking Gilbert-Peierls LU (the classic sparse partial-pivoting algorithm)
// with a caller-supplied column order.  Results of a backward-stable direct solve are determined by the matrix
// up to O(cond * eps), so this is what the 1e-9 state tolerance is anchored on; tests cross-check with SuperLU.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

namespace oracle {

struct CscMatrix {
  int n = 0;
  std::vector<int> colptr, rowind;
  std::vector<double> val;
};

class SparseLU {
public:
  // q: column elimination order (q[k] = original column eliminated k-th). threshold as in UMFPACK (0.1).
  void initialize(const CscMatrix& A, const std::vector<int>& q, double threshold = 0.1) {
    n_ = A.n;
    q_ = q;
    pinv_.assign(n_, -1);
    Lp_.assign(1, 0);
    Up_.assign(1, 0);
    Li_.clear(); Lx_.clear(); Ui_.clear(); Ux_.clear();
    std::vector<double> x(n_, 0.0);
    std::vector<int> xi(2 * n_), mark(n_, -1), stack(n_), pstack(n_);
    for (int k = 0; k < n_; ++k) {
      const int col = q[k];
      // --- symbolic: reach of A(:,col) in the graph of L (rows mapped through pinv) ---
      int top = n_;
      for (int p = A.colptr[col]; p < A.colptr[col + 1]; ++p) {
        const int i = A.rowind[p];
        if (mark[i] == k) continue;
        top = dfs(i, k, top, xi, mark, stack, pstack);
      }
      for (int p = top; p < n_; ++p) x[xi[p]] = 0.0;
      for (int p = A.colptr[col]; p < A.colptr[col + 1]; ++p) x[A.rowind[p]] = A.val[p];
      // --- numeric: x = L \ A(:,col) in topological order ---
      for (int px = top; px < n_; ++px) {
        const int i = xi[px];
        const int J = pinv_[i];
        if (J < 0) continue; // row not yet pivotal: stays in the L part
        const double xj = x[i]; // L has unit diagonal (stored first in the column)
        for (int p = Lp_[J] + 1; p < Lp_[J + 1]; ++p) x[Li_[p]] -= Lx_[p] * xj;
      }
      // --- pivot search among non-pivotal rows ---
      int ipiv = -1;
      double amax = -1.0;
      for (int p = top; p < n_; ++p) {
        const int i = xi[p];
        if (pinv_[i] < 0) {
          const double t = std::fabs(x[i]);
          if (t > amax) { amax = t; ipiv = i; }
        } else {
          Ui_.push_back(pinv_[i]);
          Ux_.push_back(x[i]);
        }
      }
      if (ipiv < 0 || amax <= 0.0) throw std::runtime_error("SparseLU: matrix is singular");
      if (pinv_[col] < 0 && std::fabs(x[col]) >= threshold * amax) ipiv = col; // prefer the diagonal
      const double pivot = x[ipiv];
      Ui_.push_back(k);
      Ux_.push_back(pivot);
      Up_.push_back((int)Ui_.size());
      pinv_[ipiv] = k;
      Li_.push_back(ipiv);
      Lx_.push_back(1.0);
      for (int p = top; p < n_; ++p) {
        const int i = xi[p];
        if (pinv_[i] < 0) {
          Li_.push_back(i);
          Lx_.push_back(x[i] / pivot);
        }
        x[i] = 0.0;
      }
      Lp_.push_back((int)Li_.size());
    }
    // row indices of L -> pivot order
    for (size_t p = 0; p < Li_.size(); ++p) Li_[p] = pinv_[Li_[p]];
  }

  // x = A^-1 b
  void vmult(double* xout, const double* b) const {
    std::vector<double> y(n_);
    for (int i = 0; i < n_; ++i) y[pinv_[i]] = b[i];
    for (int j = 0; j < n_; ++j) { // L y = Pb (unit diagonal first in column)
      const double yj = y[j];
      for (int p = Lp_[j] + 1; p < Lp_[j + 1]; ++p) y[Li_[p]] -= Lx_[p] * yj;
    }
    for (int j = n_ - 1; j >= 0; --j) { // U z = y (diagonal last in column)
      y[j] /= Ux_[Up_[j + 1] - 1];
      const double yj = y[j];
      for (int p = Up_[j]; p < Up_[j + 1] - 1; ++p) y[Ui_[p]] -= Ux_[p] * yj;
    }
    for (int k = 0; k < n_; ++k) xout[q_[k]] = y[k];
  }

  size_t nnz_factors() const { return Lx_.size() + Ux_.size(); }

private:
  // depth-first search from row i through the columns of L already computed
  int dfs(int i0, int k, int top, std::vector<int>& xi, std::vector<int>& mark, std::vector<int>& stack,
          std::vector<int>& pstack) const {
    int head = 0;
    stack[0] = i0;
    while (head >= 0) {
      const int i = stack[head];
      const int J = pinv_[i];
      if (mark[i] != k) {
        mark[i] = k;
        pstack[head] = (J < 0) ? 0 : Lp_[J] + 1;
      }
      bool done = true;
      const int pend = (J < 0) ? 0 : Lp_[J + 1];
      for (int p = pstack[head]; p < pend; ++p) {
        const int r = Li_[p]; // original row index while factoring
        if (mark[r] == k) continue;
        pstack[head] = p + 1;
        stack[++head] = r;
        done = false;
        break;
      }
      if (done) {
        --head;
        xi[--top] = i;
      }
    }
    return top;
  }

  int n_ = 0;
  std::vector<int> q_, pinv_, Lp_, Li_, Up_, Ui_;
  std::vector<double> Lx_, Ux_;
};

} // namespace oracle
