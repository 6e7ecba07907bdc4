// schur_kernels.cu -- see schur_kernels.cuh.
//
//   reduce : r~ = r_u - T1 r_q - S u_old             (before the nested-dissection solve of S du = r~)
//   recover: q  = A_qq^-1 r_q - (A_qq^-1 G_qu) u     (after it, u = u_old + du)
//   Poisson: r  = rhs - P x_old
// Both are HBM-bound streams of fixed ELL tables (12 B per stored entry), one thread per row, unit stride.
#include "schur_kernels.cuh"

#include <algorithm>

namespace pecs {

void DeviceEll::upload(const CsrMatrix& A, const std::vector<int>* row_order) {
  n = A.n;
  width = 0;
  for (int i = 0; i < n; ++i) width = std::max(width, A.row_ptr[i + 1] - A.row_ptr[i]);
  std::vector<int> c((size_t)n * width, 0);
  std::vector<double> v((size_t)n * width, 0.0);
  for (int i = 0; i < n; ++i) {
    const int r = row_order ? (*row_order)[i] : i;
    for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; ++k) {
      const size_t slot = (size_t)(k - A.row_ptr[r]) * n + i;
      c[slot] = A.col[k];
      v[slot] = A.val[k];
    }
  }
  col.upload(c);
  val.upload(v);
}

namespace {
struct EllView {
  int width;
  const int* col;
  const double* val;
  const double* x;
  double sign;
};
__device__ __forceinline__ double ell_row(const EllView& t, int n, int i) {
  double acc = 0.0;
#pragma unroll 4
  for (int k = 0; k < t.width; ++k) {
    const double a = __ldcs(t.val + (size_t)k * n + i);
    const int j = __ldcs(t.col + (size_t)k * n + i);
    acc += a * __ldg(t.x + j);
  }
  return t.sign * acc;
}
__global__ void __launch_bounds__(256) ell_combine_kernel(int n, const double* __restrict__ base,
                                                          const int* __restrict__ base_index, EllView t0, EllView t1, EllView t2,
                                                          double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = base ? base[base_index ? base_index[i] : i] : 0.0;
  acc += ell_row(t0, n, i);
  acc += ell_row(t1, n, i);
  acc += ell_row(t2, n, i);
  y[i] = acc;
}
EllView view(const EllTerm& t) {
  if (!t.A) return EllView{0, nullptr, nullptr, nullptr, 0.0};
  return EllView{t.A->width, t.A->col.get(), t.A->val.get(), t.x, t.sign};
}
} // namespace

void launch_ell_combine(int n_rows, const double* base, const int* base_index, EllTerm t0, EllTerm t1, EllTerm t2, double* y,
                        cudaStream_t s) {
  if (n_rows == 0) return;
  ell_combine_kernel<<<(n_rows + 255) / 256, 256, 0, s>>>(n_rows, base, base_index, view(t0), view(t1), view(t2), y);
}

} // namespace pecs
