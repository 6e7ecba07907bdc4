// solve_kernels.cuh -- device data layout and launch interface of the multifrontal solve sweeps.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pecs {

// A unit of work of one level kernel, self-contained (one 96-byte load, no second look-up): panels
// [panel0, panel0 + npanels) of one front's table (host/SparseDirect.hpp for the maths and the panel layout).
// Large fronts are cut into several tiles, one thread block each, the block's warps sharing the front's vector;
// small fronts are one tile each and a thread block takes kSolveWarps of them, one per warp.
struct SolveTile {
  int np, nb, p0;          // the front
  int log2P, cols_pad;     // panel shape of the table this tile belongs to (forward: G, backward: [Inv | -H])
  int panel0, npanels;
  int first;               // forward: this tile publishes the finalised pivot right-hand side of its front;
                           // backward: the front has no boundary (root), its right-hand side is finalised here
  long long table_off;     // first panel of the front's table
  long long bd_off;        // boundary index list (positions in the elimination order); the out map shares the offset
  long long cbuf_off[2];   // dense update buffers written by the two children (np + nb entries each), -1: no child
  long long out_off;       // the parent's buffer this front scatters its own update into, -1: root
  // dataflow between the level kernels (solve_kernels.cu): per-front completion counters instead of grid-wide barriers
  int front;               // this front's counter slot
  int dep[2], need[2];     // children whose forward tiles must all be complete before this tile reads their updates
  int up, need_up;         // backward: nearest ancestor that has backward tiles (-1: none) and how many it has
  int need_self;           // backward: forward tiles of this front (they publish its finalised right-hand side)
  int signal_bwd;          // backward: some descendant's tile waits for this front (else nothing needs its counter)
};

struct SolveTables {
  const int* bd_index;
  const int* out_map;
  const int* iperm;    // iperm[position in the elimination order] = unknown
  const double* fwd;
  const double* bwd;
};

constexpr int kMaxRhs = 2;
// The step-varying vectors of a solve.  Up to kMaxRhs right-hand sides share ONE pass over the factor tables (two
// carriers whose constant matrices are identical -- reductants and oxidants at equal mobility); right-hand side r uses
// w_in + r * n_stride etc. and adds its solution into solution[r].
struct SolveVectors {
  int n_rhs;
  long long n_stride;     // between the right-hand sides in w_in, w_fin, x_perm
  long long cbuf_stride;  // ... in the child-update buffers
  const double* w_in;     // residual in elimination order
  double* w_fin;          // finalised pivot right-hand sides (written by the `first` tile of every front)
  double* cbuf;           // child-update buffers
  double* x_perm;         // solution increment in elimination order (read by the deeper levels)
  double* solution[kMaxRhs];
  // sharded step over several GPUs: the backward sweep also stores every finished solution entry into the other
  // ranks' copies of the vector (peer memory over NVLink), so the exchange rides on the solve itself
  double* mirror[kMaxRhs][3];
  int n_mirror[kMaxRhs];
  // completion counters of the fronts (zeroed at the start of every solve): forward tiles done / backward tiles done
  int* done_fwd;
  int* done_bwd;
  int* error;      // set to 1 if a counter wait ever gave up (bounded spin: a logic error must not hang the GPU)
  int use_counters; // 0: every kernel waits for its whole predecessor grid, the counters are neither read nor written
  int tag;         // identifies the launch in the trace build (-DPECS_B200_TRACE=1, scripts/trace_step.py); unused otherwise
  int grid_wait;   // 1: wait for the whole predecessor grid BEFORE reading anything (first kernel behind a producer
                   // that signals no counters, or dataflow switched off); 0: counters only, grid wait at the very end
};

constexpr int kChunkDoubles = 256; // one bulk asynchronous copy: 2 KB of a panel (4 KB chunks lost in tuning and are gone)
constexpr int kSolveWarps = 16;

// shared memory of one thread block: n_rhs vectors (per block, or per warp), one ring of `stages` chunks per warp, one
// mbarrier per slot
inline size_t solve_smem_bytes(int vec_doubles, bool per_warp, int warps, int stages, int n_rhs = 1) {
  const size_t vec = ((size_t)vec_doubles + 15) / 16 * 16 * (per_warp ? warps : 1) * n_rhs;
  return (vec + (size_t)warps * stages * kChunkDoubles) * sizeof(double) + (size_t)warps * stages * sizeof(unsigned long long);
}

// forward sweep of one level.  per_warp: one small front per warp.  grid: thread blocks of the launch (level_grid() below)
void launch_forward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int grid, bool per_warp, int vec_doubles,
                          int warps, int stages, const SolveVectors& io, cudaStream_t s);
// backward sweep of one level; writes x_perm and ADDS the result to the caller's solution vectors (increment form)
void launch_backward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int grid, bool per_warp, int vec_doubles,
                           int warps, int stages, const SolveVectors& io, cudaStream_t s);
// thread blocks of a level launch: one per tile (one-warp tiles: one per `warps` tiles); with PECS_B200_LEVEL_WAVES=k
// at most k resident waves (occupancy of the kernel variant at this block shape x number of SMs), the blocks then loop
int level_grid(bool forward, bool per_warp, int n_rhs, int n_tiles, int vec_doubles, int warps, int stages);
// opt in to large dynamic shared memory once per process
void configure_solve_kernels(int max_smem_bytes);

} // namespace pecs
