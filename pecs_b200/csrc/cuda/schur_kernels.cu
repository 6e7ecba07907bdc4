// schur_kernels.cu -- see schur_kernels.cuh.
//
//   reduce : r~ = r_u - T1 r_q - S u_old             (before the nested-dissection solve of S du = r~)
//   recover: q  = A_qq^-1 r_q - (A_qq^-1 G_qu) u     (after it, u = u_old + du)
//   Poisson: r  = rhs - P x_old
// Both are HBM-bound streams of fixed ELL tables (12 B per stored entry), one thread per row, unit stride.
#include "schur_kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace pecs {

namespace {
struct EllView {
  int width, block;
  const int* col;
  const double* val;
  const double* x[2]; // up to two vectors share one pass over the table (launch_ell_combine2)
  double sign;
};
template <int NRHS>
__device__ __forceinline__ void ell_row(const EllView& t, int n, int i, double (&out)[NRHS]) {
  double acc[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) acc[r] = 0.0;
  if (t.block == 4) {
#pragma unroll 2
    for (int k = 0; k < t.width; ++k) {
      const int j = __ldcs(t.col + (size_t)k * n + i);
      const double* v = t.val + (size_t)4 * k * n + i;
      const double a0 = __ldcs(v), a1 = __ldcs(v + n), a2 = __ldcs(v + 2 * (size_t)n), a3 = __ldcs(v + 3 * (size_t)n);
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        const double2 x01 = ld_vec(reinterpret_cast<const double2*>(t.x[r] + j));
        const double2 x23 = ld_vec(reinterpret_cast<const double2*>(t.x[r] + j + 2));
        acc[r] += (a0 * x01.x + a1 * x01.y) + (a2 * x23.x + a3 * x23.y);
      }
    }
  } else {
#pragma unroll 4
    for (int k = 0; k < t.width; ++k) {
      const double a = __ldcs(t.val + (size_t)k * n + i);
      const int j = __ldcs(t.col + (size_t)k * n + i);
#pragma unroll
      for (int r = 0; r < NRHS; ++r) acc[r] += a * ld_vec(t.x[r] + j);
    }
  }
#pragma unroll
  for (int r = 0; r < NRHS; ++r) out[r] += t.sign * acc[r];
}
// base, the x of every term and y are step-varying vectors: coherent loads, no __restrict__ (device_util.cuh: ld_step)
template <int NRHS>
__global__ void __launch_bounds__(256) ell_combine_kernel(int n, EllBase base, const int* __restrict__ base_index, EllView t0,
                                                          EllView t1, EllView t2, EllOut y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  wait_for_predecessor(); // launched programmatically: the launch latency overlaps the previous kernel's tail
  if (i >= n) return;
  double acc[NRHS];
#pragma unroll
  for (int r = 0; r < NRHS; ++r) acc[r] = base.p[r] ? ld_vec(base.p[r] + (base_index ? base_index[i] : i)) : 0.0;
  // every term is added to the sum as a whole, in the order t0, t1, t2 (the summation order of round 1)
  if (t0.val) ell_row<NRHS>(t0, n, i, acc);
  if (t1.val) ell_row<NRHS>(t1, n, i, acc);
  if (t2.val) ell_row<NRHS>(t2, n, i, acc);
#pragma unroll
  for (int r = 0; r < NRHS; ++r) y.p[r][i] = acc[r];
}
EllView view(const EllTerm& t) {
  if (!t.A) return EllView{0, 1, nullptr, nullptr, {nullptr, nullptr}, 0.0};
  return EllView{t.A->width, t.A->block, t.A->col.get(), t.A->val.get(), {t.x, t.x2}, t.sign};
}
} // namespace

void launch_ell_combine(int n_rows, const double* base, const int* base_index, EllTerm t0, EllTerm t1, EllTerm t2, double* y,
                        cudaStream_t s) {
  if (n_rows == 0) return;
  launch_pdl(ell_combine_kernel<1>, (n_rows + 255) / 256, 256, 0, s, n_rows, EllBase{{base, nullptr}}, base_index, view(t0),
             view(t1), view(t2), EllOut{{y, nullptr}});
}

void launch_ell_combine2(int n_rows, const double* base, const double* base2, const int* base_index, EllTerm t0, EllTerm t1,
                         EllTerm t2, double* y, double* y2, cudaStream_t s) {
  if (n_rows == 0) return;
  launch_pdl(ell_combine_kernel<2>, (n_rows + 255) / 256, 256, 0, s, n_rows, EllBase{{base, base2}}, base_index, view(t0),
             view(t1), view(t2), EllOut{{y, y2}});
}

} // namespace pecs
