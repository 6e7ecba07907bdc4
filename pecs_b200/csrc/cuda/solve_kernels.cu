// solve_kernels.cu -- the two sweeps of the multifrontal solve as level-batched dense mat-vecs (sm_100a).
//
// Replaces SparseDirectUMFPACK::vmult, i.e. the sequential sparse triangular substitutions the reference runs for
// every Carrier::solve / PoissonData::solve (reference source/Carrier.cpp:34-40, source/Poisson.cpp:98-105).
// With the explicit front operators built at setup (host/SparseDirect.hpp) one solve is
//     forward level kernels  (deepest level -> root):  w_P = b_P - children,  t = children + G w_P  -> parent
//     backward level kernels (root -> deepest level):  x_P = [Inv | -H] [w_P ; x_B]
// The only traffic that matters is the factor tables: every entry is streamed from HBM exactly once per solve
// (8 B per stored entry, nothing is re-read).  The kernels are built around that stream:
//   * tables are stored as contiguous PANELS (P rows, column-major, host/SparseDirect.hpp); one warp owns a panel;
//   * every warp runs its own pipeline of bulk asynchronous copies (cp.async.bulk global -> shared, completion on an
//     mbarrier): lane 0 keeps `stages` 2 KB chunks of the warp's panels in flight, independent of what the warp's
//     arithmetic is doing, so the bytes in flight per SM are set by shared memory, not by registers or occupancy;
//   * 32 consecutive doubles of a chunk are 32/P columns of the panel's P rows: lane l multiplies entry 32 s + l with
//     vector element (column) and accumulates for row l % P -- no shuffles until the end of the panel;
//   * the front's input vector is staged once per thread block in shared memory while the first chunks are in flight;
//   * a front's update goes to a dense buffer in its parent's local numbering: the parent reads it with unit stride;
//   * the permutations in and out of the elimination order are fused into the staging / the final store;
//   * every output row has exactly one owner and a fixed summation order: no atomics, bit-reproducible solves.
#include "solve_kernels.cuh"

namespace pecs {

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP); the table is read once per solve: L2 evict-first
__device__ __forceinline__ void bulk_copy(double* dst, const double* src, uint32_t bytes, unsigned long long* bar,
                                          unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_addr(dst)),
      "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ unsigned long long evict_first_policy() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// One warp streams its share of the tile's panels (panel0 + warp, + n_warps, ...) through its private ring and hands
// every finished row to emit(row, value).  sv: the front's vector in shared memory, zero beyond the logical columns.
template <class Emit>
__device__ __forceinline__ void stream_panels(const double* __restrict__ table, int log2P, int cols_pad, int rows,
                                              const SolveTile& tile, const double* sv, double* ring,
                                              unsigned long long* bars, int stages, Emit emit) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int P = 1 << log2P;
  const int cg = 32 >> log2P;                               // columns covered by 32 consecutive doubles
  const int panel_doubles = cols_pad << log2P;
  const int cpp = (panel_doubles + kChunkDoubles - 1) / kChunkDoubles; // chunks per panel
  const int n_my = warp < tile.npanels ? (tile.npanels - warp + n_warps - 1) / n_warps : 0;
  const int total = n_my * cpp;
  double* my_ring = ring + (size_t)warp * stages * kChunkDoubles;
  unsigned long long* my_bars = bars + warp * stages;
  const unsigned long long policy = evict_first_policy();

  // producer state (lane 0 only): next chunk to issue
  int ik = 0, ic = 0, issued = 0, islot = 0;
  auto issue = [&]() {
    const int panel = tile.panel0 + warp + ik * n_warps;
    const int e0 = ic * kChunkDoubles;
    const int elems = min(kChunkDoubles, panel_doubles - e0);
    mbar_expect_tx(my_bars + islot, (uint32_t)elems * 8u);
    bulk_copy(my_ring + islot * kChunkDoubles, table + (size_t)panel * panel_doubles + e0, (uint32_t)elems * 8u,
              my_bars + islot, policy);
    ++issued;
    if (++islot == stages) islot = 0;
    if (++ic == cpp) {
      ic = 0;
      ++ik;
    }
  };
  if (lane == 0)
    for (int q = 0; q < stages && q < total; ++q) issue();

  // the vector is staged by the whole block while the first chunks fly
  __syncthreads();

  const int row_in_panel = lane & (P - 1);
  const int col_of_lane = lane >> log2P;
  int slot = 0;
  uint32_t phase = 0;
  for (int k = 0; k < n_my; ++k) {
    const int panel = tile.panel0 + warp + k * n_warps;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int c = 0; c < cpp; ++c) {
      mbar_wait(my_bars + slot, phase);
      const double* ch = my_ring + slot * kChunkDoubles + lane;
      const int e0 = c * kChunkDoubles;
      const int elems = min(kChunkDoubles, panel_doubles - e0);
      const double* v = sv + (e0 >> log2P) + col_of_lane;
      if (elems == kChunkDoubles) {
        a0 += ch[0] * v[0];
        a1 += ch[32] * v[cg];
        a2 += ch[64] * v[2 * cg];
        a3 += ch[96] * v[3 * cg];
        a0 += ch[128] * v[4 * cg];
        a1 += ch[160] * v[5 * cg];
        a2 += ch[192] * v[6 * cg];
        a3 += ch[224] * v[7 * cg];
      } else {
        for (int s = 0; s < elems; s += 32) a0 += ch[s] * v[(s >> 5) * cg];
      }
      __syncwarp();
      if (lane == 0 && issued < total) issue();
      if (++slot == stages) {
        slot = 0;
        phase ^= 1u;
      }
    }
    double sum = (a0 + a1) + (a2 + a3);
    for (int o = P; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const int row = (panel << log2P) + row_in_panel;
    if (lane < P && row < rows) emit(row, sum);
  }
}

__device__ __forceinline__ void init_pipeline(unsigned long long* bars, int stages) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    for (int q = 0; q < stages; ++q) mbar_init(bars + warp * stages + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kSolveWarps * 32) forward_level_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                        int vec_doubles, int stages,
                                                                        const double* __restrict__ rhs,
                                                                        double* __restrict__ w_fin, double* cbuf) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sv = reinterpret_cast<double*>(smem_raw);
  double* ring = sv + vec_doubles;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring + (size_t)(blockDim.x >> 5) * stages * kChunkDoubles);
  const SolveTile tile = tiles[blockIdx.x];
  const DeviceFront F = t.fronts[tile.front];
  init_pipeline(bars, stages);
  // finalised pivot right-hand side: w_P = b_P - what the children eliminated into it
  {
    const double* c0 = F.cbuf_off[0] >= 0 ? cbuf + F.cbuf_off[0] : nullptr;
    const double* c1 = F.cbuf_off[1] >= 0 ? cbuf + F.cbuf_off[1] : nullptr;
    const int count = max(F.np, F.fwd_cols_pad);
    for (int l = threadIdx.x; l < count; l += blockDim.x) {
      double v = 0.0;
      if (l < F.np) {
        v = rhs[t.iperm[F.p0 + l]];
        if (c0) v -= c0[l];
        if (c1) v -= c1[l];
        if (tile.first) w_fin[F.p0 + l] = v;
      }
      if (l < F.fwd_cols_pad) sv[l] = v;
    }
  }
  const double* carry0 = F.cbuf_off[0] >= 0 ? cbuf + F.cbuf_off[0] + F.np : nullptr;
  const double* carry1 = F.cbuf_off[1] >= 0 ? cbuf + F.cbuf_off[1] + F.np : nullptr;
  const int* omap = t.out_map + F.bd_off;
  double* out = cbuf + F.out_off;
  if (F.np == 0) {
    // a front without pivots (its region fell apart into unconnected pieces) only hands its children's updates on
    for (int row = threadIdx.x; row < F.nb; row += blockDim.x) {
      double carry = 0.0;
      if (carry0) carry += carry0[row];
      if (carry1) carry += carry1[row];
      out[omap[row]] = carry;
    }
  }
  stream_panels(t.fwd + F.fwd_off, F.fwd_log2P, F.fwd_cols_pad, F.nb, tile, sv, ring, bars, stages, [&](int row, double dot) {
    double carry = 0.0;
    if (carry0) carry += carry0[row];
    if (carry1) carry += carry1[row];
    out[omap[row]] = carry + dot;
  });
}

__global__ void __launch_bounds__(kSolveWarps * 32) backward_level_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                         int vec_doubles, int stages,
                                                                         const double* __restrict__ w_fin, double* x_perm,
                                                                         double* __restrict__ solution) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sv = reinterpret_cast<double*>(smem_raw);
  double* ring = sv + vec_doubles;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring + (size_t)(blockDim.x >> 5) * stages * kChunkDoubles);
  const SolveTile tile = tiles[blockIdx.x];
  const DeviceFront F = t.fronts[tile.front];
  init_pipeline(bars, stages);
  {
    const int np = F.np, m = F.np + F.nb;
    const int* bd = t.bd_index + F.bd_off;
    const double* wp = w_fin + F.p0;
    for (int l = threadIdx.x; l < F.bwd_cols_pad; l += blockDim.x) sv[l] = l < np ? wp[l] : (l < m ? x_perm[bd[l - np]] : 0.0);
  }
  stream_panels(t.bwd + F.bwd_off, F.bwd_log2P, F.bwd_cols_pad, F.np, tile, sv, ring, bars, stages, [&](int row, double x) {
    x_perm[F.p0 + row] = x;
    solution[t.iperm[F.p0 + row]] = x;
  });
}

} // namespace

void configure_solve_kernels(int max_smem_bytes) {
  cudaFuncSetAttribute(forward_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
  cudaFuncSetAttribute(backward_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
}

void launch_forward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int vec_doubles, int warps, int stages,
                          const double* rhs, double* w_fin, double* cbuf, cudaStream_t s) {
  if (n_tiles == 0) return;
  const int vec = (vec_doubles + 15) / 16 * 16;
  forward_level_kernel<<<n_tiles, warps * 32, solve_smem_bytes(vec_doubles, warps, stages), s>>>(t, tiles, vec, stages, rhs,
                                                                                                w_fin, cbuf);
}

void launch_backward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int vec_doubles, int warps, int stages,
                           const double* w_fin, double* x_perm, double* solution, cudaStream_t s) {
  if (n_tiles == 0) return;
  const int vec = (vec_doubles + 15) / 16 * 16;
  backward_level_kernel<<<n_tiles, warps * 32, solve_smem_bytes(vec_doubles, warps, stages), s>>>(t, tiles, vec, stages, w_fin,
                                                                                                 x_perm, solution);
}

} // namespace pecs
