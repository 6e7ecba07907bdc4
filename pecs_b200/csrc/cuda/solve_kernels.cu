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
//   * every output row has exactly one owner and a fixed summation order: no atomics, bit-reproducible solves;
//   * DATAFLOW between levels.  A sweep is a chain of programmatically launched level kernels.  Round 1 made every
//     kernel wait for its whole predecessor grid (griddepcontrol.wait) before touching a vector: 15-30 grid drains per
//     sweep, the reason a lone solve reached half the HBM peak and the Poisson solve a quarter.  Now a kernel releases
//     its successor at once (griddepcontrol.launch_dependents first thing), so the blocks of the next levels become
//     resident and prefetch their tables while this level still computes, and a tile waits only for the FRONTS it
//     reads: per-front completion counters (forward: the two children; backward: the nearest ancestor with tiles and
//     the front's own forward tiles), released with __threadfence + atomicAdd by the producing tile, acquired with
//     ld.acquire.gpu by the consumer, data read from L2 (ld_step).  No deadlock: blocks of kernel k+1 are scheduled only
//     after ALL blocks of kernel k have started, so whatever a spinning block waits for is already running.  Every
//     block executes griddepcontrol.wait as its LAST instruction, so a kernel completes only after its predecessor has:
//     completion stays transitive along the chain, and whatever follows the chain (an event, a plainly launched kernel,
//     a kernel that waits for its predecessor grid) sees the whole sweep finished.  PECS_B200_DATAFLOW=0 restores the
//     grid waits (A/B; bit-identical results either way);
#include "solve_kernels.cuh"

#include <algorithm>
#include <cstdlib>

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
  // the warp's ring: lives for the whole kernel, across all the tiles the warp works on (mbarrier phases carry over)
  double* ring;
  unsigned long long* bars;
  unsigned long long policy;
  int lane, stages;
  int slot = 0, islot = 0; // consumer / producer position in the ring
  uint32_t phase = 0;
  // the tile being streamed
  const double* table = nullptr;
  int log2P = 0, panel_doubles = 0, cpp = 0, n_my = 0, total = 0;
  int panel0 = 0, rank = 0, n_ranks = 1;
  int ik = 0, ic = 0, issued = 0; // producer state (lane 0 only): next chunk to issue

  __device__ __forceinline__ PanelStream(double* my_ring, unsigned long long* my_bars, int stages_)
      : ring(my_ring), bars(my_bars), policy(evict_first_policy()), lane(threadIdx.x & 31), stages(stages_) {}
  // next tile; the ring is empty here (every issued chunk of the previous tile has been consumed)
  __device__ __forceinline__ void begin(const double* table_, const SolveTile& tile, int rank_, int n_ranks_) {
    table = table_;
    log2P = tile.log2P;
    panel_doubles = tile.cols_pad << tile.log2P;
    panel0 = tile.panel0;
    rank = rank_;
    n_ranks = n_ranks_;
    cpp = (panel_doubles + CHUNK - 1) / CHUNK; // chunks per panel
    n_my = rank < tile.npanels ? (tile.npanels - rank + n_ranks - 1) / n_ranks : 0;
    total = n_my * cpp;
    ik = ic = issued = 0;
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
  // NRHS right-hand sides share ONE pass over the table (the two redox carriers of the default input have identical
  // matrices, reference source/LDG.cpp:624-678 + equal mobilities): sv holds NRHS vectors `vec_stride` apart; every
  // right-hand side keeps the accumulation order of the single-vector kernel, so results are bit-identical to two solves
  template <int NRHS, class Emit>
  __device__ __forceinline__ void run(const double* sv, int vec_stride, int rows, Emit emit) {
    const int P = 1 << log2P;
    const int cg = 32 >> log2P; // columns covered by 32 consecutive doubles
    const int row_in_panel = lane & (P - 1);
    const int col_of_lane = lane >> log2P;
    for (int k = 0; k < n_my; ++k) {
      const int panel = panel0 + rank + k * n_ranks;
      double a[NRHS][4];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) a[r][0] = a[r][1] = a[r][2] = a[r][3] = 0.0;
      for (int c = 0; c < cpp; ++c) {
        mbar_wait(bars + slot, phase);
        const double* ch = ring + slot * CHUNK + lane;
        const int e0 = c * CHUNK;
        const int elems = min(CHUNK, panel_doubles - e0);
        const double* v = sv + (e0 >> log2P) + col_of_lane;
        if (elems == CHUNK) {
#pragma unroll
          for (int s = 0; s < CHUNK / 32; s += 4) {
            const double t0 = ch[32 * s], t1 = ch[32 * s + 32], t2 = ch[32 * s + 64], t3 = ch[32 * s + 96];
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
              const double* vr = v + r * vec_stride;
              a[r][0] += t0 * vr[s * cg];
              a[r][1] += t1 * vr[(s + 1) * cg];
              a[r][2] += t2 * vr[(s + 2) * cg];
              a[r][3] += t3 * vr[(s + 3) * cg];
            }
          }
        } else {
          for (int s = 0; s < elems; s += 32) {
            const double t0 = ch[s];
#pragma unroll
            for (int r = 0; r < NRHS; ++r) a[r][0] += t0 * v[r * vec_stride + (s >> 5) * cg];
          }
        }
        __syncwarp();
        if (lane == 0 && issued < total) issue();
        if (++slot == stages) {
          slot = 0;
          phase ^= 1u;
        }
      }
      double sum[NRHS];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        sum[r] = (a[r][0] + a[r][1]) + (a[r][2] + a[r][3]);
        for (int o = P; o < 32; o <<= 1) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
      }
      const int row = (panel << log2P) + row_in_panel;
      if (lane < P && row < rows) emit(row, sum);
    }
  }
};

// ---- per-front completion counters
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// one thread: spin until *counter >= need.  Bounded (about a second): a logic error flags `error` instead of hanging
__device__ __forceinline__ void wait_count(const int* counter, int need, int* error) {
  if (need <= 0) return;
  unsigned spins = 0;
  while (ld_acquire(counter) < need) {
    __nanosleep(32);
    ++spins;
    if ((spins & 1023u) == 0 && ld_acquire(error) != 0) break; // someone already gave up: do not queue up behind it
    if (spins > (1u << 20)) {
      atomicExch(error, 1);
      break;
    }
  }
}
// one thread, after the barrier that follows the tile's last store: publish "one more tile of this front is complete".
// One release-reduction: the stores of the whole block (ordered before it by the barrier) become visible before the count
__device__ __forceinline__ void signal_count(int* counter) {
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(counter) : "memory");
}
// ---- trace build only (python -m pecs_b200.build --variant trace): per-block timestamps of the level kernels, the ground
// truth about what overlaps with what (scripts/trace_step.py).  Record = {tag << 32 | block, start, dependencies met, end}
#ifndef PECS_B200_TRACE
#define PECS_B200_TRACE 0
#endif
#if PECS_B200_TRACE
__device__ unsigned long long* g_trace = nullptr;
__device__ unsigned int g_trace_cap = 0;
__device__ unsigned int g_trace_n = 0;
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_record(int tag, unsigned long long t0, unsigned long long t1, unsigned long long t2) {
  if (!g_trace) return;
  const unsigned int k = atomicAdd(&g_trace_n, 1u);
  if (k >= g_trace_cap) return;
  g_trace[4 * (size_t)k] = ((unsigned long long)(unsigned)tag << 32) | blockIdx.x;
  g_trace[4 * (size_t)k + 1] = t0;
  g_trace[4 * (size_t)k + 2] = t1;
  g_trace[4 * (size_t)k + 3] = t2;
}
#define PECS_TRACE_T(name) const unsigned long long name = trace_now()
#define PECS_TRACE_REC(tag, a, b, c) if (first_thread == 0) trace_record(tag, a, b, c)
#else
#define PECS_TRACE_T(name)
#define PECS_TRACE_REC(tag, a, b, c)
#endif

__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void release_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
// vec_doubles: all NRHS vectors of one front (NRHS * padded length)
template <bool PER_WARP, int CHUNK>
__device__ __forceinline__ BlockSmem carve(unsigned char* raw, int vec_doubles, int stages) {
  const int warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  double* base = reinterpret_cast<double*>(raw);
  double* ring = base + (size_t)vec_doubles * (PER_WARP ? n_warps : 1);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring + (size_t)n_warps * stages * CHUNK);
  return BlockSmem{base + (PER_WARP ? (size_t)warp * vec_doubles : 0), ring + (size_t)warp * stages * CHUNK,
                   bars + warp * stages};
}

// Tiles map to thread blocks statically (block tiles: tile = blockIdx; one-warp tiles: one per warp), a launch normally has
// a block per tile and the loop below runs once.  The loop exists for launches with FEWER blocks than tiles
// (PECS_B200_LEVEL_WAVES: grids capped at so many resident waves).  Measured with the per-block-timestamp trace
// (profiles/r02_trace_*.log): a single resident wave per level that takes tiles off a counter is SLOWER (electron solve
// 776 us against 651 us): the blocks of level k+1 then occupy the SM slots for the whole level while they spin on
// counters of level k, whose remaining tiles wait for slots -- the hardware's block scheduler, which recycles a slot per
// tile, is the better dynamic scheduler here.

template <bool PER_WARP, int CHUNK, int NRHS>
__global__ void __launch_bounds__(kSolveWarps * 32) forward_level_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                        int n_tiles, int vec_doubles, int stages,
                                                                        SolveVectors io) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const BlockSmem sm = carve<PER_WARP, CHUNK>(smem_raw, vec_doubles * NRHS, stages);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int first_thread = PER_WARP ? lane : threadIdx.x, n_threads = PER_WARP ? 32 : blockDim.x;
  init_pipeline(sm.my_bars, stages);
  PanelStream<CHUNK> stream(sm.my_ring, sm.my_bars, stages);
  if (io.grid_wait) grid_dependency_wait();
  release_dependents();
  const int tile_stride = (int)gridDim.x * (PER_WARP ? n_warps : 1);
  for (int tile_index = PER_WARP ? blockIdx.x * n_warps + warp : blockIdx.x; tile_index < n_tiles; tile_index += tile_stride) {
    PECS_TRACE_T(trace_t0);
    const SolveTile tile = tiles[tile_index];
    stream.begin(t.fwd + tile.table_off, tile, PER_WARP ? 0 : warp, PER_WARP ? 1 : n_warps);
    stream.start();
    // the children's updates (their forward tiles) must be complete: two threads poll the two counters side by side
    if (io.use_counters) {
      if (first_thread < 2 && tile.dep[first_thread] >= 0)
        wait_count(io.done_fwd + tile.dep[first_thread], tile.need[first_thread], io.error);
      if (PER_WARP)
        __syncwarp();
      else
        __syncthreads();
    }
    PECS_TRACE_T(trace_t1);
    const int* omap = t.out_map + tile.bd_off;
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
      const double* cb = io.cbuf + (size_t)r * io.cbuf_stride;
      const double* c0 = tile.cbuf_off[0] >= 0 ? cb + tile.cbuf_off[0] : nullptr;
      const double* c1 = tile.cbuf_off[1] >= 0 ? cb + tile.cbuf_off[1] : nullptr;
      const double* w_in = io.w_in + (size_t)r * io.n_stride;
      double* w_fin = io.w_fin + (size_t)r * io.n_stride;
      double* sv = sm.sv + r * vec_doubles;
      // finalised pivot right-hand side: w_P = b_P - what the children eliminated into it
      const int count = max(tile.np, tile.cols_pad);
      for (int l = first_thread; l < count; l += n_threads) {
        double v = 0.0;
        if (l < tile.np) {
          v = ld_step(w_in + tile.p0 + l);
          if (c0) v -= ld_step(c0 + l);
          if (c1) v -= ld_step(c1 + l);
          if (tile.first) w_fin[tile.p0 + l] = v;
        }
        if (l < tile.cols_pad) sv[l] = v;
      }
      if (tile.np == 0) {
        // a front without pivots (its region fell apart into unconnected pieces) only hands its children's updates on
        double* out = io.cbuf + (size_t)r * io.cbuf_stride + tile.out_off;
        for (int row = first_thread; row < tile.nb; row += n_threads) {
          double carry = 0.0;
          if (c0) carry += ld_step(c0 + tile.np + row);
          if (c1) carry += ld_step(c1 + tile.np + row);
          out[omap[row]] = carry;
        }
      }
    }
    if (PER_WARP)
      __syncwarp();
    else
      __syncthreads();
    stream.template run<NRHS>(sm.sv, vec_doubles, tile.nb, [&](int row, const double (&dot)[NRHS]) {
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        double* cb = io.cbuf + (size_t)r * io.cbuf_stride;
        double carry = 0.0;
        if (tile.cbuf_off[0] >= 0) carry += ld_step(cb + tile.cbuf_off[0] + tile.np + row);
        if (tile.cbuf_off[1] >= 0) carry += ld_step(cb + tile.cbuf_off[1] + tile.np + row);
        cb[tile.out_off + omap[row]] = carry + dot[r];
      }
    });
    if (PER_WARP)
      __syncwarp();
    else
      __syncthreads();
    if (io.use_counters && first_thread == 0 && tile.out_off >= 0) signal_count(io.done_fwd + tile.front);
    PECS_TRACE_T(trace_t2);
    PECS_TRACE_REC(io.tag, trace_t0, trace_t1, trace_t2);
  }
  if (!io.grid_wait) grid_dependency_wait(); // completion stays transitive along the kernel chain
}

template <bool PER_WARP, int CHUNK, int NRHS>
__global__ void __launch_bounds__(kSolveWarps * 32) backward_level_kernel(SolveTables t, const SolveTile* __restrict__ tiles,
                                                                         int n_tiles, int vec_doubles, int stages,
                                                                         SolveVectors io) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const BlockSmem sm = carve<PER_WARP, CHUNK>(smem_raw, vec_doubles * NRHS, stages);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int first_thread = PER_WARP ? lane : threadIdx.x, n_threads = PER_WARP ? 32 : blockDim.x;
  init_pipeline(sm.my_bars, stages);
  PanelStream<CHUNK> stream(sm.my_ring, sm.my_bars, stages);
  if (io.grid_wait) grid_dependency_wait();
  release_dependents();
  const int tile_stride = (int)gridDim.x * (PER_WARP ? n_warps : 1);
  for (int tile_index = PER_WARP ? blockIdx.x * n_warps + warp : blockIdx.x; tile_index < n_tiles; tile_index += tile_stride) {
    PECS_TRACE_T(trace_t0);
    const SolveTile tile = tiles[tile_index];
    stream.begin(t.bwd + tile.table_off, tile, PER_WARP ? 0 : warp, PER_WARP ? 1 : n_warps);
    stream.start();
    // x of every ancestor (the nearest one with tiles waited for its own ancestors), this front's finalised right-hand
    // side (its forward tiles), and -- a front without boundary reads its children's updates itself -- the children
    // (four threads poll the up-to-four counters side by side: one L2 round trip instead of four)
    if (io.use_counters) {
      if (first_thread == 0 && tile.up >= 0) wait_count(io.done_bwd + tile.up, tile.need_up, io.error);
      if (first_thread == 1) wait_count(io.done_fwd + tile.front, tile.need_self, io.error);
      if (tile.first && (first_thread == 2 || first_thread == 3) && tile.dep[first_thread - 2] >= 0)
        wait_count(io.done_fwd + tile.dep[first_thread - 2], tile.need[first_thread - 2], io.error);
      if (PER_WARP)
        __syncwarp();
      else
        __syncthreads();
    }
    PECS_TRACE_T(trace_t1);
    {
      const int np = tile.np, m = tile.np + tile.nb;
      const int* bd = t.bd_index + tile.bd_off;
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        double* sv = sm.sv + r * vec_doubles;
        if (tile.first) {
          // no boundary, hence no forward tile: w_P = b_P - what the children eliminated into it, finalised here
          const double* cb = io.cbuf + (size_t)r * io.cbuf_stride;
          const double* c0 = tile.cbuf_off[0] >= 0 ? cb + tile.cbuf_off[0] : nullptr;
          const double* c1 = tile.cbuf_off[1] >= 0 ? cb + tile.cbuf_off[1] : nullptr;
          const double* w_in = io.w_in + (size_t)r * io.n_stride;
          for (int l = first_thread; l < tile.cols_pad; l += n_threads) {
            double v = 0.0;
            if (l < np) {
              v = ld_step(w_in + tile.p0 + l);
              if (c0) v -= ld_step(c0 + l);
              if (c1) v -= ld_step(c1 + l);
            }
            sv[l] = v;
          }
        } else {
          const double* wp = io.w_fin + (size_t)r * io.n_stride + tile.p0;
          const double* xp = io.x_perm + (size_t)r * io.n_stride;
          for (int l = first_thread; l < tile.cols_pad; l += n_threads)
            sv[l] = l < np ? ld_step(wp + l) : (l < m ? ld_step(xp + bd[l - np]) : 0.0);
        }
      }
    }
    if (PER_WARP)
      __syncwarp();
    else
      __syncthreads();
    stream.template run<NRHS>(sm.sv, vec_doubles, tile.np, [&](int row, const double (&x)[NRHS]) {
      const int i = t.iperm[tile.p0 + row];
#pragma unroll
      for (int r = 0; r < NRHS; ++r) {
        io.x_perm[(size_t)r * io.n_stride + tile.p0 + row] = x[r];
        double* solution = io.solution[r];
        const double v = ld_step(solution + i) + x[r]; // increment form: the right-hand side was the residual of `solution`
        solution[i] = v;
        for (int m = 0; m < io.n_mirror[r]; ++m) io.mirror[r][m][i] = v; // peer copies (sharded step)
      }
    });
    if (PER_WARP)
      __syncwarp();
    else
      __syncthreads();
    // only fronts some deeper front waits for publish their completion (leaf fronts, half of all, do not)
    if (io.use_counters && first_thread == 0 && tile.signal_bwd) signal_count(io.done_bwd + tile.front);
    PECS_TRACE_T(trace_t2);
    PECS_TRACE_REC(io.tag, trace_t0, trace_t1, trace_t2);
  }
  if (!io.grid_wait) grid_dependency_wait(); // completion stays transitive along the kernel chain
}

} // namespace

template <bool PER_WARP, int NRHS>
void configure_one(int max_smem_bytes) {
  cudaFuncSetAttribute(forward_level_kernel<PER_WARP, kChunkDoubles, NRHS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       max_smem_bytes);
  cudaFuncSetAttribute(backward_level_kernel<PER_WARP, kChunkDoubles, NRHS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       max_smem_bytes);
}
void configure_solve_kernels(int max_smem_bytes) {
  configure_one<false, 1>(max_smem_bytes);
  configure_one<true, 1>(max_smem_bytes);
  configure_one<false, 2>(max_smem_bytes);
  configure_one<true, 2>(max_smem_bytes);
}

int level_grid(bool forward, bool per_warp, int n_rhs, int n_tiles, int vec_doubles, int warps, int stages) {
  if (n_tiles == 0) return 0;
  const size_t smem = solve_smem_bytes(vec_doubles, per_warp, warps, stages, n_rhs);
  const void* kernel = nullptr;
#define PECS_PICK(PW, NR) \
  kernel = forward ? (const void*)forward_level_kernel<PW, kChunkDoubles, NR> : (const void*)backward_level_kernel<PW, kChunkDoubles, NR>
  if (n_rhs == 2) {
    if (per_warp) PECS_PICK(true, 2); else PECS_PICK(false, 2);
  } else {
    if (per_warp) PECS_PICK(true, 1); else PECS_PICK(false, 1);
  }
#undef PECS_PICK
  int per_sm = 0, dev = 0, sms = 0;
  PECS_CUDA(cudaGetDevice(&dev));
  PECS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PECS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
  const int units = per_warp ? (n_tiles + warps - 1) / warps : n_tiles;
  static const int waves = [] {
    const char* e = std::getenv("PECS_B200_LEVEL_WAVES");
    return e ? std::atoi(e) : 0;
  }();
  if (waves <= 0) return units; // default: a block per tile (one-warp tiles: per `warps` tiles)
  return std::max(1, std::min(units, waves * std::max(per_sm, 1) * sms));
}

void launch_forward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int grid, bool per_warp, int vec_doubles,
                          int warps, int stages, const SolveVectors& io, cudaStream_t s) {
  if (n_tiles == 0) return;
  const int vec = (vec_doubles + 15) / 16 * 16;
  const size_t smem = solve_smem_bytes(vec_doubles, per_warp, warps, stages, io.n_rhs);
#define PECS_LAUNCH(PW, NR) \
  launch_pdl(forward_level_kernel<PW, kChunkDoubles, NR>, grid, warps * 32, smem, s, t, tiles, n_tiles, vec, stages, io)
  if (io.n_rhs == 2) {
    if (per_warp) PECS_LAUNCH(true, 2); else PECS_LAUNCH(false, 2);
  } else {
    if (per_warp) PECS_LAUNCH(true, 1); else PECS_LAUNCH(false, 1);
  }
#undef PECS_LAUNCH
}

void launch_backward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int grid, bool per_warp, int vec_doubles,
                           int warps, int stages, const SolveVectors& io, cudaStream_t s) {
  if (n_tiles == 0) return;
  const int vec = (vec_doubles + 15) / 16 * 16;
  const size_t smem = solve_smem_bytes(vec_doubles, per_warp, warps, stages, io.n_rhs);
#define PECS_LAUNCH(PW, NR) \
  launch_pdl(backward_level_kernel<PW, kChunkDoubles, NR>, grid, warps * 32, smem, s, t, tiles, n_tiles, vec, stages, io)
  if (io.n_rhs == 2) {
    if (per_warp) PECS_LAUNCH(true, 2); else PECS_LAUNCH(false, 2);
  } else {
    if (per_warp) PECS_LAUNCH(true, 1); else PECS_LAUNCH(false, 1);
  }
#undef PECS_LAUNCH
}

#if PECS_B200_TRACE
} // namespace pecs
extern "C" int pecs_trace_start(int capacity) {
  static unsigned long long* buffer = nullptr;
  if (buffer) cudaFree(buffer);
  buffer = nullptr;
  if (capacity > 0 && cudaMalloc(&buffer, (size_t)capacity * 4 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  const unsigned int cap = capacity > 0 ? (unsigned)capacity : 0, zero = 0;
  cudaMemcpyToSymbol(pecs::g_trace, &buffer, sizeof(buffer));
  cudaMemcpyToSymbol(pecs::g_trace_cap, &cap, sizeof(cap));
  cudaMemcpyToSymbol(pecs::g_trace_n, &zero, sizeof(zero));
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}
extern "C" int pecs_trace_read(unsigned long long* out, int max_records) {
  cudaDeviceSynchronize();
  unsigned int n = 0, cap = 0;
  unsigned long long* buffer = nullptr;
  cudaMemcpyFromSymbol(&n, pecs::g_trace_n, sizeof(n));
  cudaMemcpyFromSymbol(&cap, pecs::g_trace_cap, sizeof(cap));
  cudaMemcpyFromSymbol(&buffer, pecs::g_trace, sizeof(buffer));
  const unsigned int count = n < cap ? n : cap;
  const int take = (int)count < max_records ? (int)count : max_records;
  if (take > 0) cudaMemcpy(out, buffer, (size_t)take * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return take;
}
namespace pecs {
#endif

} // namespace pecs
