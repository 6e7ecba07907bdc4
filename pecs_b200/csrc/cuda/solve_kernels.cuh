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
};

struct SolveTables {
  const int* bd_index;
  const int* out_map;
  const int* iperm;    // iperm[position in the elimination order] = unknown
  const double* fwd;
  const double* bwd;
  // sharded step over several GPUs: the backward sweep also stores every finished solution entry into the other
  // ranks' copies of the vector (peer memory over NVLink), so the exchange rides on the solve itself
  double* mirror[3];
  int n_mirror;
};

constexpr int kChunkDoubles = 256; // one bulk asynchronous copy: 2 KB of a panel (512 = 4 KB is the other compiled variant)
constexpr int kSolveWarps = 8;

// shared memory of one thread block: the vector (one per block, or one per warp), one ring of `stages` chunks per
// warp, one mbarrier per slot
inline size_t solve_smem_bytes(int vec_doubles, bool per_warp, int warps, int stages, int chunk = kChunkDoubles) {
  const size_t vec = ((size_t)vec_doubles + 15) / 16 * 16 * (per_warp ? warps : 1);
  return (vec + (size_t)warps * stages * chunk) * sizeof(double) + (size_t)warps * stages * sizeof(unsigned long long);
}

// forward sweep of one level.  w_in: right-hand side in elimination order; w_fin: finalised pivot right-hand sides
// (written by the `first` tile of every front); cbuf: child-update buffers.  per_warp: one small front per warp.
void launch_forward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, bool per_warp, int vec_doubles, int warps,
                          int stages, int chunk, const double* w_in, double* w_fin, double* cbuf, cudaStream_t s);
// backward sweep of one level; writes x_perm (elimination order, read by the deeper levels) and ADDS the result to the
// caller's solution vector (increment form)
void launch_backward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, bool per_warp, int vec_doubles, int warps,
                           int stages, int chunk, const double* w_in, const double* cbuf, const double* w_fin, double* x_perm,
                           double* solution, cudaStream_t s);
// out[i] = in[index[i]]
void launch_gather(int n, const int* index, const double* in, double* out, cudaStream_t s);
// opt in to large dynamic shared memory once per process
void configure_solve_kernels(int max_smem_bytes);

} // namespace pecs
