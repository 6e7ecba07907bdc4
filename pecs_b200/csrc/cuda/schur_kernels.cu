// schur_kernels.cu -- see schur_kernels.cuh.
//
//   reduce : r~ = r_u - T1 r_q                       (before the nested-dissection solve of S u = r~)
//   recover: q  = A_qq^-1 r_q - (A_qq^-1 G_qu) u     (after it)
// Both are HBM-bound streams of fixed ELL tables (12 B per stored entry), one thread per row, unit stride.
#include "schur_kernels.cuh"

#include <algorithm>

namespace pecs {

void DeviceEll::upload(const CsrMatrix& A) {
  n = A.n;
  width = 0;
  for (int i = 0; i < n; ++i) width = std::max(width, A.row_ptr[i + 1] - A.row_ptr[i]);
  std::vector<int> c((size_t)n * width, 0);
  std::vector<double> v((size_t)n * width, 0.0);
  for (int i = 0; i < n; ++i)
    for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k) {
      const size_t slot = (size_t)(k - A.row_ptr[i]) * n + i;
      c[slot] = A.col[k];
      v[slot] = A.val[k];
    }
  col.upload(c);
  val.upload(v);
}

namespace {
__global__ void __launch_bounds__(256) ell_combine_kernel(int n, const double* __restrict__ base, int w1,
                                                          const int* __restrict__ c1, const double* __restrict__ v1,
                                                          const double* __restrict__ x1, int w2, const int* __restrict__ c2,
                                                          const double* __restrict__ v2, const double* __restrict__ x2,
                                                          double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = base ? base[i] : 0.0;
#pragma unroll 4
  for (int k = 0; k < w1; ++k) {
    const double a = __ldcs(v1 + (size_t)k * n + i);
    const int j = __ldcs(c1 + (size_t)k * n + i);
    acc += a * __ldg(x1 + j);
  }
#pragma unroll 4
  for (int k = 0; k < w2; ++k) {
    const double a = __ldcs(v2 + (size_t)k * n + i);
    const int j = __ldcs(c2 + (size_t)k * n + i);
    acc -= a * __ldg(x2 + j);
  }
  y[i] = acc;
}
} // namespace

void launch_ell_combine(int n_rows, const double* base, const DeviceEll* A1, const double* x1, const DeviceEll& A2,
                        const double* x2, double* y, cudaStream_t s) {
  if (n_rows == 0) return;
  ell_combine_kernel<<<(n_rows + 255) / 256, 256, 0, s>>>(n_rows, base, A1 ? A1->width : 0, A1 ? A1->col.get() : nullptr,
                                                        A1 ? A1->val.get() : nullptr, x1, A2.width, A2.col.get(),
                                                        A2.val.get(), x2, y);
}

} // namespace pecs
