// solve_kernels.cu -- the two sweeps of the multifrontal solve as level-batched dense mat-vecs (sm_100a).
//
// Replaces SparseDirectUMFPACK::vmult, i.e. the sequential sparse triangular substitutions the reference runs for
// every Carrier::solve / PoissonData::solve (reference source/Carrier.cpp:34-40, source/Poisson.cpp:98-105).
// With the explicit front operators built at setup (host/SparseDirect.hpp) one solve is
//     forward level kernels  (deepest level -> root):  w_P = b_P - children,  t = children + G w_P  -> parent
//     backward level kernels (root -> deepest level):  x_P = [Inv | -H] [w_P ; x_B]
// The only traffic that matters is the factor tables: every entry is streamed from HBM exactly once per solve
// (8 B per stored entry).  Layout and mapping are chosen for that stream:
//   * rows padded to even length, 16-byte streaming loads (ld.global.cs), 8 independent loads in flight per lane;
//   * large fronts: row-major tables, one warp owns two rows at a time, the front's small input vector sits in
//     shared memory (or is gathered on the fly for the few root-level fronts whose vector would cost occupancy);
//   * small fronts (np <= 128, the vast majority): column-major forward table, one THREAD owns two rows, no
//     shuffles at all, a warp still reads 512 contiguous bytes per instruction;
//   * a front's update goes to a dense buffer in its parent's local numbering: the parent reads it with unit stride;
//   * every output row has exactly one owner: no atomics, fixed summation order, bit-reproducible solves.
#include "solve_kernels.cuh"

namespace pecs {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dot products of TWO table rows (n2 double2 each) with a vector delivered by vec(j) -> double2
template <class VecFn>
__device__ __forceinline__ void dot2(const double2* __restrict__ A, const double2* __restrict__ B, int n2, int lane,
                                     VecFn vec, double& outA, double& outB) {
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  int j = lane;
  for (; j + 96 < n2; j += 128) {
    const double2 xa0 = __ldcs(A + j), xa1 = __ldcs(A + j + 32), xa2 = __ldcs(A + j + 64), xa3 = __ldcs(A + j + 96);
    const double2 xb0 = __ldcs(B + j), xb1 = __ldcs(B + j + 32), xb2 = __ldcs(B + j + 64), xb3 = __ldcs(B + j + 96);
    const double2 s0 = vec(j), s1 = vec(j + 32), s2 = vec(j + 64), s3 = vec(j + 96);
    a0 += xa0.x * s0.x;
    a1 += xa0.y * s0.y;
    a2 += xa1.x * s1.x;
    a3 += xa1.y * s1.y;
    a0 += xa2.x * s2.x;
    a1 += xa2.y * s2.y;
    a2 += xa3.x * s3.x;
    a3 += xa3.y * s3.y;
    b0 += xb0.x * s0.x;
    b1 += xb0.y * s0.y;
    b2 += xb1.x * s1.x;
    b3 += xb1.y * s1.y;
    b0 += xb2.x * s2.x;
    b1 += xb2.y * s2.y;
    b2 += xb3.x * s3.x;
    b3 += xb3.y * s3.y;
  }
  for (; j < n2; j += 32) {
    const double2 xa = __ldcs(A + j), xb = __ldcs(B + j), s0 = vec(j);
    a0 += xa.x * s0.x;
    a1 += xa.y * s0.y;
    b0 += xb.x * s0.x;
    b1 += xb.y * s0.y;
  }
  outA = warp_sum((a0 + a1) + (a2 + a3));
  outB = warp_sum((b0 + b1) + (b2 + b3));
}

// finalised pivot right-hand side of a front into shared memory (and, once per front, into w_fin)
__device__ __forceinline__ void stage_pivot_rhs(const DeviceFront& F, bool publish, const double* __restrict__ w_in,
                                                double* __restrict__ w_fin, const double* cbuf, double* sv) {
  const double* c0 = F.cbuf_off[0] >= 0 ? cbuf + F.cbuf_off[0] : nullptr;
  const double* c1 = F.cbuf_off[1] >= 0 ? cbuf + F.cbuf_off[1] : nullptr;
  for (int l = threadIdx.x; l < F.np; l += blockDim.x) {
    double v = w_in[F.p0 + l];
    if (c0) v -= c0[l];
    if (c1) v -= c1[l];
    sv[l] = v;
    if (publish) w_fin[F.p0 + l] = v;
  }
  if (threadIdx.x == 0 && (F.np & 1)) sv[F.np] = 0.0;
}

// t[row] = (children's updates on row) + dot ; scattered into the parent's buffer in the parent's numbering
__device__ __forceinline__ void emit_update(const DeviceFront& F, const SolveTables& t, double* cbuf, int row, double dot) {
  double carry = 0.0;
  if (F.cbuf_off[0] >= 0) carry += cbuf[F.cbuf_off[0] + F.np + row];
  if (F.cbuf_off[1] >= 0) carry += cbuf[F.cbuf_off[1] + F.np + row];
  cbuf[F.out_off + t.out_map[F.bd_off + row]] = carry + dot;
}

__global__ void __launch_bounds__(kSolveThreads) forward_rows_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                     const double* __restrict__ w_in,
                                                                     double* __restrict__ w_fin, double* cbuf) {
  extern __shared__ __align__(16) double sv[];
  const SolveTile tile = tiles[blockIdx.x];
  const DeviceFront F = t.fronts[tile.front];
  stage_pivot_rhs(F, tile.first != 0, w_in, w_fin, cbuf, sv);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int n2 = F.ld_fwd >> 1;
  const double2* G = reinterpret_cast<const double2*>(t.fwd + F.fwd_off);
  const double2* S2 = reinterpret_cast<const double2*>(sv);
  for (int r = 2 * warp; r < tile.nrows; r += 2 * n_warps) {
    const int rowA = tile.row0 + r;
    const bool hasB = r + 1 < tile.nrows;
    const int rowB = hasB ? rowA + 1 : rowA;
    double dA, dB;
    dot2(G + (size_t)rowA * n2, G + (size_t)rowB * n2, n2, lane, [&](int j) { return S2[j]; }, dA, dB);
    if (lane == 0) emit_update(F, t, cbuf, rowA, dA);
    if (lane == 1 && hasB) emit_update(F, t, cbuf, rowB, dB);
  }
}

__global__ void __launch_bounds__(kSolveThreads) forward_cols_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                     const double* __restrict__ w_in,
                                                                     double* __restrict__ w_fin, double* cbuf) {
  __shared__ __align__(16) double sv[130];
  const SolveTile tile = tiles[blockIdx.x];
  const DeviceFront F = t.fronts[tile.front];
  stage_pivot_rhs(F, tile.first != 0, w_in, w_fin, cbuf, sv);
  __syncthreads();
  const int r = 2 * threadIdx.x;
  if (r >= tile.nrows) return;
  const int row = tile.row0 + r;
  const int ld2 = F.ld_fwd >> 1;
  const double2* G = reinterpret_cast<const double2*>(t.fwd + F.fwd_off + row);
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  int j = 0;
  for (; j + 8 <= F.np; j += 8) {
    double2 g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) g[u] = __ldcs(G + (size_t)(j + u) * ld2);
#pragma unroll
    for (int u = 0; u < 8; u += 2) {
      a0 += g[u].x * sv[j + u];
      b0 += g[u].y * sv[j + u];
      a1 += g[u + 1].x * sv[j + u + 1];
      b1 += g[u + 1].y * sv[j + u + 1];
    }
  }
  for (; j < F.np; ++j) {
    const double2 g = __ldcs(G + (size_t)j * ld2);
    a0 += g.x * sv[j];
    b0 += g.y * sv[j];
  }
  emit_update(F, t, cbuf, row, a0 + a1);
  if (row + 1 < F.nb) emit_update(F, t, cbuf, row + 1, b0 + b1);
}

template <bool STAGED>
__global__ void __launch_bounds__(kSolveThreads) backward_rows_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                      const double* __restrict__ w_fin, double* x_perm) {
  extern __shared__ __align__(16) double sv[];
  const SolveTile tile = tiles[blockIdx.x];
  const DeviceFront F = t.fronts[tile.front];
  const int np = F.np, m = F.np + F.nb;
  const int* bd = t.bd_index + F.bd_off;
  const double* wp = w_fin + F.p0;
  if (STAGED) {
    for (int l = threadIdx.x; l < m; l += blockDim.x) sv[l] = l < np ? wp[l] : x_perm[bd[l - np]];
    if (threadIdx.x == 0 && (m & 1)) sv[m] = 0.0;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int n2 = F.ld_bwd >> 1;
  const double2* B = reinterpret_cast<const double2*>(t.bwd + F.bwd_off);
  const double2* S2 = reinterpret_cast<const double2*>(sv);
  auto element = [&](int e) -> double { return e < np ? wp[e] : (e < m ? x_perm[bd[e - np]] : 0.0); };
  for (int r = 2 * warp; r < tile.nrows; r += 2 * n_warps) {
    const int rowA = tile.row0 + r;
    const bool hasB = r + 1 < tile.nrows;
    const int rowB = hasB ? rowA + 1 : rowA;
    double dA, dB;
    if (STAGED)
      dot2(B + (size_t)rowA * n2, B + (size_t)rowB * n2, n2, lane, [&](int j) { return S2[j]; }, dA, dB);
    else
      dot2(B + (size_t)rowA * n2, B + (size_t)rowB * n2, n2, lane,
           [&](int j) { return make_double2(element(2 * j), element(2 * j + 1)); }, dA, dB);
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
  cudaFuncSetAttribute(forward_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
  cudaFuncSetAttribute(backward_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
}

void launch_forward_rows(const SolveTables& t, const SolveTile* tiles, int n_tiles, int smem_doubles, const double* w_in,
                         double* w_fin, double* cbuf, cudaStream_t s) {
  if (n_tiles == 0) return;
  forward_rows_kernel<<<n_tiles, kSolveThreads, (size_t)smem_doubles * sizeof(double), s>>>(t, tiles, w_in, w_fin, cbuf);
}

void launch_forward_cols(const SolveTables& t, const SolveTile* tiles, int n_tiles, const double* w_in, double* w_fin,
                         double* cbuf, cudaStream_t s) {
  if (n_tiles == 0) return;
  forward_cols_kernel<<<n_tiles, kSolveThreads, 0, s>>>(t, tiles, w_in, w_fin, cbuf);
}

void launch_backward_rows(const SolveTables& t, const SolveTile* tiles, int n_tiles, int smem_doubles, const double* w_fin,
                          double* x_perm, cudaStream_t s) {
  if (n_tiles == 0) return;
  if (smem_doubles > 0)
    backward_rows_kernel<true><<<n_tiles, kSolveThreads, (size_t)smem_doubles * sizeof(double), s>>>(t, tiles, w_fin, x_perm);
  else
    backward_rows_kernel<false><<<n_tiles, kSolveThreads, 0, s>>>(t, tiles, w_fin, x_perm);
}

void launch_gather(int n, const int* index, const double* in, double* out, cudaStream_t s) {
  if (n == 0) return;
  gather_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, index, in, out);
}

} // namespace pecs
