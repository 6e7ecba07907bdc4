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
//   * solves are done in increment form (x += A^-1 (b - A x), cuda/context.cu): the residual arrives in elimination
//     order from the fused ELL kernel and the backward sweep adds the correction to the caller's vector in place;
//   * every output row has exactly one owner and a fixed summation order: no atomics, bit-reproducible solves.
#include "solve_kernels.cuh"

#include "device_util.cuh"

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

// One warp streams panels panel0 + rank, panel0 + rank + n_ranks, ... of a tile through its private ring of bulk
// copies.  start() puts the first chunks in flight -- it touches only the static table, so it runs BEFORE the kernel
// waits for its predecessor (programmatic dependent launch) and before the vector is staged; run() consumes the
// chunks against the vector sv in shared memory (zero beyond the logical columns), keeps the ring full and hands
// every finished row to emit(row, value).
template <int CHUNK>
struct PanelStream {
  const double* table;
  double* ring;
  unsigned long long* bars;
  unsigned long long policy;
  int lane, log2P, panel_doubles, cpp, n_my, total, stages;
  int panel0, rank, n_ranks;
  int ik = 0, ic = 0, issued = 0, islot = 0; // producer state (lane 0 only): next chunk to issue

  __device__ __forceinline__ PanelStream(const double* table_, const SolveTile& tile, int rank_, int n_ranks_, double* my_ring,
                                         unsigned long long* my_bars, int stages_)
      : table(table_), ring(my_ring), bars(my_bars), policy(evict_first_policy()), lane(threadIdx.x & 31), log2P(tile.log2P),
        panel_doubles(tile.cols_pad << tile.log2P), stages(stages_), panel0(tile.panel0), rank(rank_), n_ranks(n_ranks_) {
    cpp = (panel_doubles + CHUNK - 1) / CHUNK; // chunks per panel
    n_my = rank < tile.npanels ? (tile.npanels - rank + n_ranks - 1) / n_ranks : 0;
    total = n_my * cpp;
  }
  __device__ __forceinline__ void issue() {
    const int panel = panel0 + rank + ik * n_ranks;
    const int e0 = ic * CHUNK;
    const int elems = min(CHUNK, panel_doubles - e0);
    mbar_expect_tx(bars + islot, (uint32_t)elems * 8u);
    bulk_copy(ring + islot * CHUNK, table + (size_t)panel * panel_doubles + e0, (uint32_t)elems * 8u, bars + islot, policy);
    ++issued;
    if (++islot == stages) islot = 0;
    if (++ic == cpp) {
      ic = 0;
      ++ik;
    }
  }
  __device__ __forceinline__ void start() {
    if (lane == 0)
      for (int q = 0; q < stages && q < total; ++q) issue();
  }
  template <class Emit>
  __device__ __forceinline__ void run(const double* sv, int rows, Emit emit) {
    const int P = 1 << log2P;
    const int cg = 32 >> log2P; // columns covered by 32 consecutive doubles
    const int row_in_panel = lane & (P - 1);
    const int col_of_lane = lane >> log2P;
    int slot = 0;
    uint32_t phase = 0;
    for (int k = 0; k < n_my; ++k) {
      const int panel = panel0 + rank + k * n_ranks;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      for (int c = 0; c < cpp; ++c) {
        mbar_wait(bars + slot, phase);
        const double* ch = ring + slot * CHUNK + lane;
        const int e0 = c * CHUNK;
        const int elems = min(CHUNK, panel_doubles - e0);
        const double* v = sv + (e0 >> log2P) + col_of_lane;
        if (elems == CHUNK) {
#pragma unroll
          for (int s = 0; s < CHUNK / 32; s += 4) {
            a0 += ch[32 * s] * v[s * cg];
            a1 += ch[32 * s + 32] * v[(s + 1) * cg];
            a2 += ch[32 * s + 64] * v[(s + 2) * cg];
            a3 += ch[32 * s + 96] * v[(s + 3) * cg];
          }
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
};

__device__ __forceinline__ void init_pipeline(unsigned long long* my_bars, int stages) {
  if ((threadIdx.x & 31) == 0) {
    for (int q = 0; q < stages; ++q) mbar_init(my_bars + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

// shared memory carve-up common to both sweeps
struct BlockSmem {
  double* sv;                  // this thread's vector (the block's, or its warp's)
  double* my_ring;
  unsigned long long* my_bars;
};
template <bool PER_WARP, int CHUNK>
__device__ __forceinline__ BlockSmem carve(unsigned char* raw, int vec_doubles, int stages) {
  const int warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  double* base = reinterpret_cast<double*>(raw);
  double* ring = base + (size_t)vec_doubles * (PER_WARP ? n_warps : 1);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring + (size_t)n_warps * stages * CHUNK);
  return BlockSmem{base + (PER_WARP ? (size_t)warp * vec_doubles : 0), ring + (size_t)warp * stages * CHUNK,
                   bars + warp * stages};
}

template <bool PER_WARP, int CHUNK>
__global__ void __launch_bounds__(kSolveWarps * 32) forward_level_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                        int n_tiles, int vec_doubles, int stages,
                                                                        const double* __restrict__ w_in,
                                                                        double* __restrict__ w_fin, double* cbuf) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const BlockSmem sm = carve<PER_WARP, CHUNK>(smem_raw, vec_doubles, stages);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int tile_index = PER_WARP ? blockIdx.x * n_warps + warp : blockIdx.x;
  if (PER_WARP && tile_index >= n_tiles) return;
  const SolveTile tile = tiles[tile_index];
  init_pipeline(sm.my_bars, stages);
  PanelStream<CHUNK> stream(t.fwd + tile.table_off, tile, PER_WARP ? 0 : warp, PER_WARP ? 1 : n_warps, sm.my_ring, sm.my_bars,
                            stages);
  stream.start();
  wait_for_predecessor();
  const int first_thread = PER_WARP ? lane : threadIdx.x, n_threads = PER_WARP ? 32 : blockDim.x;
  const double* c0 = tile.cbuf_off[0] >= 0 ? cbuf + tile.cbuf_off[0] : nullptr;
  const double* c1 = tile.cbuf_off[1] >= 0 ? cbuf + tile.cbuf_off[1] : nullptr;
  // finalised pivot right-hand side: w_P = b_P - what the children eliminated into it
  {
    const int count = max(tile.np, tile.cols_pad);
    for (int l = first_thread; l < count; l += n_threads) {
      double v = 0.0;
      if (l < tile.np) {
        v = w_in[tile.p0 + l];
        if (c0) v -= c0[l];
        if (c1) v -= c1[l];
        if (tile.first) w_fin[tile.p0 + l] = v;
      }
      if (l < tile.cols_pad) sm.sv[l] = v;
    }
  }
  const double* carry0 = c0 ? c0 + tile.np : nullptr;
  const double* carry1 = c1 ? c1 + tile.np : nullptr;
  const int* omap = t.out_map + tile.bd_off;
  double* out = cbuf + tile.out_off;
  if (tile.np == 0) {
    // a front without pivots (its region fell apart into unconnected pieces) only hands its children's updates on
    for (int row = first_thread; row < tile.nb; row += n_threads) {
      double carry = 0.0;
      if (carry0) carry += carry0[row];
      if (carry1) carry += carry1[row];
      out[omap[row]] = carry;
    }
  }
  if (PER_WARP)
    __syncwarp();
  else
    __syncthreads();
  stream.run(sm.sv, tile.nb, [&](int row, double dot) {
    double carry = 0.0;
    if (carry0) carry += carry0[row];
    if (carry1) carry += carry1[row];
    out[omap[row]] = carry + dot;
  });
}

template <bool PER_WARP, int CHUNK>
__global__ void __launch_bounds__(kSolveWarps * 32) backward_level_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                         int n_tiles, int vec_doubles, int stages,
                                                                         const double* __restrict__ w_in,
                                                                         const double* __restrict__ cbuf,
                                                                         const double* __restrict__ w_fin, double* x_perm,
                                                                         double* solution) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const BlockSmem sm = carve<PER_WARP, CHUNK>(smem_raw, vec_doubles, stages);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int tile_index = PER_WARP ? blockIdx.x * n_warps + warp : blockIdx.x;
  if (PER_WARP && tile_index >= n_tiles) return;
  const SolveTile tile = tiles[tile_index];
  init_pipeline(sm.my_bars, stages);
  PanelStream<CHUNK> stream(t.bwd + tile.table_off, tile, PER_WARP ? 0 : warp, PER_WARP ? 1 : n_warps, sm.my_ring, sm.my_bars,
                            stages);
  stream.start();
  wait_for_predecessor();
  {
    const int first_thread = PER_WARP ? lane : threadIdx.x, n_threads = PER_WARP ? 32 : blockDim.x;
    const int np = tile.np, m = tile.np + tile.nb;
    const int* bd = t.bd_index + tile.bd_off;
    if (tile.first) {
      // no boundary, hence no forward tile: w_P = b_P - what the children eliminated into it, finalised here
      const double* c0 = tile.cbuf_off[0] >= 0 ? cbuf + tile.cbuf_off[0] : nullptr;
      const double* c1 = tile.cbuf_off[1] >= 0 ? cbuf + tile.cbuf_off[1] : nullptr;
      for (int l = first_thread; l < tile.cols_pad; l += n_threads) {
        double v = 0.0;
        if (l < np) {
          v = w_in[tile.p0 + l];
          if (c0) v -= c0[l];
          if (c1) v -= c1[l];
        }
        sm.sv[l] = v;
      }
    } else {
      const double* wp = w_fin + tile.p0;
      for (int l = first_thread; l < tile.cols_pad; l += n_threads)
        sm.sv[l] = l < np ? wp[l] : (l < m ? x_perm[bd[l - np]] : 0.0);
    }
  }
  if (PER_WARP)
    __syncwarp();
  else
    __syncthreads();
  stream.run(sm.sv, tile.np, [&](int row, double x) {
    x_perm[tile.p0 + row] = x;
    const int i = t.iperm[tile.p0 + row];
    const double v = solution[i] + x; // increment form: the right-hand side was the residual of `solution`
    solution[i] = v;
    for (int m = 0; m < t.n_mirror; ++m) t.mirror[m][i] = v; // peer copies (sharded step)
  });
}

__global__ void gather_kernel(int n, const int* __restrict__ index, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[index[i]];
}

} // namespace

template <bool PER_WARP, int CHUNK>
void configure_one(int max_smem_bytes) {
  cudaFuncSetAttribute(forward_level_kernel<PER_WARP, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
  cudaFuncSetAttribute(backward_level_kernel<PER_WARP, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_bytes);
}
void configure_solve_kernels(int max_smem_bytes) {
  configure_one<false, 256>(max_smem_bytes);
  configure_one<true, 256>(max_smem_bytes);
  configure_one<false, 512>(max_smem_bytes);
  configure_one<true, 512>(max_smem_bytes);
}

void launch_forward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, bool per_warp, int vec_doubles, int warps,
                          int stages, int chunk, const double* w_in, double* w_fin, double* cbuf, cudaStream_t s) {
  if (n_tiles == 0) return;
  const int vec = (vec_doubles + 15) / 16 * 16;
  const size_t smem = solve_smem_bytes(vec_doubles, per_warp, warps, stages, chunk);
  const int grid = per_warp ? (n_tiles + warps - 1) / warps : n_tiles;
#define PECS_LAUNCH(PW, CH) \
  launch_pdl(forward_level_kernel<PW, CH>, grid, warps * 32, smem, s, t, tiles, n_tiles, vec, stages, w_in, w_fin, cbuf)
  if (chunk == 512) {
    if (per_warp) PECS_LAUNCH(true, 512); else PECS_LAUNCH(false, 512);
  } else {
    if (per_warp) PECS_LAUNCH(true, 256); else PECS_LAUNCH(false, 256);
  }
#undef PECS_LAUNCH
}

void launch_backward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, bool per_warp, int vec_doubles, int warps,
                           int stages, int chunk, const double* w_in, const double* cbuf, const double* w_fin, double* x_perm,
                           double* solution, cudaStream_t s) {
  if (n_tiles == 0) return;
  const int vec = (vec_doubles + 15) / 16 * 16;
  const size_t smem = solve_smem_bytes(vec_doubles, per_warp, warps, stages, chunk);
  const int grid = per_warp ? (n_tiles + warps - 1) / warps : n_tiles;
#define PECS_LAUNCH(PW, CH)                                                                                           \
  launch_pdl(backward_level_kernel<PW, CH>, grid, warps * 32, smem, s, t, tiles, n_tiles, vec, stages, w_in, cbuf, w_fin, \
             x_perm, solution)
  if (chunk == 512) {
    if (per_warp) PECS_LAUNCH(true, 512); else PECS_LAUNCH(false, 512);
  } else {
    if (per_warp) PECS_LAUNCH(true, 256); else PECS_LAUNCH(false, 256);
  }
#undef PECS_LAUNCH
}

void launch_gather(int n, const int* index, const double* in, double* out, cudaStream_t s) {
  if (n == 0) return;
  gather_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, index, in, out);
}

} // namespace pecs
