// solve_kernels.cuh -- device data layout and launch interface of the multifrontal solve sweeps.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pecs {

// one front as the kernels see it (see host/SparseDirect.hpp for the maths and the panel layout of the tables)
struct DeviceFront {
  int np, nb, p0;
  int fwd_log2P, fwd_cols_pad; // G          : nb rows, np columns
  int bwd_log2P, bwd_cols_pad; // [Inv | -H] : np rows, np + nb columns
  long long bd_off;      // boundary index list (positions in the permuted vector); the out map shares the offset
  long long fwd_off;     // first panel of G
  long long bwd_off;     // first panel of [Inv | -H]
  long long cbuf_off[2]; // dense update buffers written by the two children (np+nb entries each), -1: no child
  long long out_off;     // the parent's buffer this front scatters its own update into, -1: root
};

// a unit of work of one level kernel (one thread block): panels [panel0, panel0+npanels) of one front's table
struct SolveTile {
  int front, panel0, npanels, first; // first != 0: this tile also publishes the finalised pivot right-hand side
};

struct SolveTables {
  const DeviceFront* fronts;
  const int* bd_index;
  const int* out_map;
  const int* iperm;    // iperm[position in the elimination order] = unknown
  const double* fwd;
  const double* bwd;
};

constexpr int kChunkDoubles = 256; // one bulk asynchronous copy: 2 KB of a panel
constexpr int kSolveWarps = 8;

// shared memory of one thread block: the front's vector, one ring of `stages` chunks per warp, one mbarrier per slot
inline size_t solve_smem_bytes(int vec_doubles, int warps, int stages) {
  const size_t vec = ((size_t)vec_doubles + 15) / 16 * 16;
  return (vec + (size_t)warps * stages * kChunkDoubles) * sizeof(double) + (size_t)warps * stages * sizeof(unsigned long long);
}

// forward sweep of one level.  rhs: right-hand side in the caller's numbering (read through iperm); w_fin: finalised
// pivot right-hand sides in elimination order (written by the `first` tile of every front); cbuf: child-update buffers.
void launch_forward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int vec_doubles, int warps, int stages,
                          const double* rhs, double* w_fin, double* cbuf, cudaStream_t s);
// backward sweep of one level; writes x_perm (elimination order, read by the deeper levels) and the caller's solution
void launch_backward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int vec_doubles, int warps, int stages,
                           const double* w_fin, double* x_perm, double* solution, cudaStream_t s);
// opt in to large dynamic shared memory once per process
void configure_solve_kernels(int max_smem_bytes);

} // namespace pecs
