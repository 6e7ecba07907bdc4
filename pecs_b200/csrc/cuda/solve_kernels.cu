// solve_kernels.cu -- the two sweeps of the multifrontal solve as level-batched dense mat-vecs (sm_100a).
//
// Replaces SparseDirectUMFPACK::vmult, i.e. the sequential sparse triangular substitutions the reference runs for
// every Carrier::solve / PoissonData::solve (reference source/Carrier.cpp:34-40, source/Poisson.cpp:98-105).
// With the explicit front operators built at setup (host/SparseDirect.hpp) one solve is
//     forward level kernels  (deepest level -> root):  w_P = b_P - children,  t = children + G w_P
//     backward level kernels (root -> deepest level):  x_P = [Inv | -H] [w_P ; x_B]
// Every kernel streams its slice of the factor tables exactly once from HBM (that is all the traffic that matters:
// 8 bytes per stored factor entry per solve), rows are contiguous so a warp reads 256 B (or 512 B with 16-byte
// loads) per instruction, the small input vector of a front is staged in shared memory, and each output row is
// owned by one warp: no atomics, fixed summation order, bit-reproducible solves.
#include "solve_kernels.cuh"

namespace pecs {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dot products of TWO rows with the staged vector: 2 x 4 independent loads in flight per lane
template <bool VEC2>
__device__ __forceinline__ void dot2(const double* __restrict__ rowA, const double* __restrict__ rowB,
                                     const double* __restrict__ sv, int ncols, int lane, double& outA, double& outB) {
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  if (VEC2) {
    const int n2 = ncols >> 1;
    const double2* A2 = reinterpret_cast<const double2*>(rowA);
    const double2* B2 = reinterpret_cast<const double2*>(rowB);
    const double2* S2 = reinterpret_cast<const double2*>(sv);
    int j = lane;
    for (; j + 32 < n2; j += 64) {
      const double2 xa = __ldcs(A2 + j), ya = __ldcs(A2 + j + 32);
      const double2 xb = __ldcs(B2 + j), yb = __ldcs(B2 + j + 32);
      const double2 s0 = S2[j], s1 = S2[j + 32];
      a0 += xa.x * s0.x;
      a1 += xa.y * s0.y;
      a2 += ya.x * s1.x;
      a3 += ya.y * s1.y;
      b0 += xb.x * s0.x;
      b1 += xb.y * s0.y;
      b2 += yb.x * s1.x;
      b3 += yb.y * s1.y;
    }
    if (j < n2) {
      const double2 xa = __ldcs(A2 + j), xb = __ldcs(B2 + j), s0 = S2[j];
      a0 += xa.x * s0.x;
      a1 += xa.y * s0.y;
      b0 += xb.x * s0.x;
      b1 += xb.y * s0.y;
    }
  } else {
    int j = lane;
    for (; j + 96 < ncols; j += 128) {
      const double xa0 = __ldcs(rowA + j), xa1 = __ldcs(rowA + j + 32), xa2 = __ldcs(rowA + j + 64), xa3 = __ldcs(rowA + j + 96);
      const double xb0 = __ldcs(rowB + j), xb1 = __ldcs(rowB + j + 32), xb2 = __ldcs(rowB + j + 64), xb3 = __ldcs(rowB + j + 96);
      a0 += xa0 * sv[j];
      a1 += xa1 * sv[j + 32];
      a2 += xa2 * sv[j + 64];
      a3 += xa3 * sv[j + 96];
      b0 += xb0 * sv[j];
      b1 += xb1 * sv[j + 32];
      b2 += xb2 * sv[j + 64];
      b3 += xb3 * sv[j + 96];
    }
    for (; j < ncols; j += 32) {
      a0 += __ldcs(rowA + j) * sv[j];
      b0 += __ldcs(rowB + j) * sv[j];
    }
  }
  outA = warp_sum((a0 + a1) + (a2 + a3));
  outB = warp_sum((b0 + b1) + (b2 + b3));
}

template <bool VEC2>
__global__ void __launch_bounds__(kSolveThreads) forward_level_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                      const double* __restrict__ w_in,
                                                                      double* __restrict__ w_fin, double* upd) {
  extern __shared__ __align__(16) double sv[];
  const SolveTile tile = tiles[blockIdx.x];
  const DeviceFront F = t.fronts[tile.front];
  const int np = F.np;
  const int* cmap0 = t.child_map + F.cmap_off[0];
  const int* cmap1 = t.child_map + F.cmap_off[1];
  const double* upd0 = upd + F.child_upd_off[0];
  const double* upd1 = upd + F.child_upd_off[1];
  // finalise the pivot right-hand side: subtract what the children eliminated into it
  for (int l = threadIdx.x; l < np; l += blockDim.x) {
    double val = w_in[F.p0 + l];
    if (F.has_child[0]) {
      const int s = cmap0[l];
      if (s >= 0) val -= upd0[s];
    }
    if (F.has_child[1]) {
      const int s = cmap1[l];
      if (s >= 0) val -= upd1[s];
    }
    sv[l] = val;
    if (tile.first) w_fin[F.p0 + l] = val;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const double* G = t.fwd + F.fwd_off;
  double* out = upd + F.upd_off;
  for (int r = 2 * warp; r < tile.nrows; r += 2 * n_warps) {
    const int rowA = tile.row0 + r;
    const bool hasB = r + 1 < tile.nrows;
    const int rowB = hasB ? rowA + 1 : rowA;
    double dA, dB;
    dot2<VEC2>(G + (size_t)rowA * np, G + (size_t)rowB * np, sv, np, lane, dA, dB);
    if (lane < 2 && (lane == 0 || hasB)) {
      const int row = lane == 0 ? rowA : rowB;
      double carry = 0.0;
      if (F.has_child[0]) {
        const int s = cmap0[np + row];
        if (s >= 0) carry += upd0[s];
      }
      if (F.has_child[1]) {
        const int s = cmap1[np + row];
        if (s >= 0) carry += upd1[s];
      }
      out[row] = carry + (lane == 0 ? dA : dB);
    }
  }
}

template <bool VEC2>
__global__ void __launch_bounds__(kSolveThreads) backward_level_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                       const double* __restrict__ w_fin, double* x_perm) {
  extern __shared__ __align__(16) double sv[];
  const SolveTile tile = tiles[blockIdx.x];
  const DeviceFront F = t.fronts[tile.front];
  const int np = F.np, m = F.np + F.nb;
  const int* bd = t.bd_index + F.bd_off;
  for (int l = threadIdx.x; l < m; l += blockDim.x) sv[l] = l < np ? w_fin[F.p0 + l] : x_perm[bd[l - np]];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const double* B = t.bwd + F.bwd_off;
  for (int r = 2 * warp; r < tile.nrows; r += 2 * n_warps) {
    const int rowA = tile.row0 + r;
    const bool hasB = r + 1 < tile.nrows;
    const int rowB = hasB ? rowA + 1 : rowA;
    double dA, dB;
    dot2<VEC2>(B + (size_t)rowA * m, B + (size_t)rowB * m, sv, m, lane, dA, dB);
    if (lane == 0) x_perm[F.p0 + rowA] = dA;
    if (lane == 1 && hasB) x_perm[F.p0 + rowB] = dB;
  }
}

__global__ void gather_kernel(int n, const int* __restrict__ index, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[index[i]];
}

} // namespace

void configure_solve_kernels(int max_smem_bytes) {
  cudaFuncSetAttribute(forward_level_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
  cudaFuncSetAttribute(forward_level_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
  cudaFuncSetAttribute(backward_level_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
  cudaFuncSetAttribute(backward_level_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
}

void launch_forward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int smem_doubles, bool vec2,
                          const double* w_in, double* w_fin, double* upd, cudaStream_t s) {
  if (n_tiles == 0) return;
  const size_t smem = (size_t)smem_doubles * sizeof(double);
  if (vec2)
    forward_level_kernel<true><<<n_tiles, kSolveThreads, smem, s>>>(t, tiles, w_in, w_fin, upd);
  else
    forward_level_kernel<false><<<n_tiles, kSolveThreads, smem, s>>>(t, tiles, w_in, w_fin, upd);
}

void launch_backward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int smem_doubles, bool vec2,
                           const double* w_fin, double* x_perm, cudaStream_t s) {
  if (n_tiles == 0) return;
  const size_t smem = (size_t)smem_doubles * sizeof(double);
  if (vec2)
    backward_level_kernel<true><<<n_tiles, kSolveThreads, smem, s>>>(t, tiles, w_fin, x_perm);
  else
    backward_level_kernel<false><<<n_tiles, kSolveThreads, smem, s>>>(t, tiles, w_fin, x_perm);
}

void launch_gather(int n, const int* index, const double* in, double* out, cudaStream_t s) {
  if (n == 0) return;
  gather_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, index, in, out);
}

} // namespace pecs
