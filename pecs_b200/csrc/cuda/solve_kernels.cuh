// solve_kernels.cuh -- device data layout and launch interface of the multifrontal solve sweeps.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pecs {

// one front as the kernels see it (see host/SparseDirect.hpp for the maths and the table layouts)
struct DeviceFront {
  int np, nb, p0;
  int ld_fwd, ld_bwd, fwd_colmajor;
  long long bd_off;      // boundary index list (positions in the permuted vector); the out map shares the offset
  long long fwd_off;     // G
  long long bwd_off;     // [Inv | -H]
  long long cbuf_off[2]; // dense update buffers written by the two children (np+nb entries each), -1: no child
  long long out_off;     // the parent's buffer this front scatters its own update into, -1: root
};

// a unit of work of one level kernel: rows [row0, row0+nrows) of one front's table
struct SolveTile {
  int front, row0, nrows, first; // first != 0: this tile also publishes the finalised pivot right-hand side
};

struct SolveTables {
  const DeviceFront* fronts;
  const int* bd_index;
  const int* out_map;
  const double* fwd;
  const double* bwd;
};

constexpr int kSolveThreads = 128;          // 4 warps, 2 rows per warp and pass
constexpr int kColTileRows = 2 * kSolveThreads; // column-major forward kernel: 2 rows per thread
constexpr int kBackwardStageMax = 6144;     // backward vectors up to this length are staged in shared memory

// forward sweep of one level.  w_in: permuted right-hand side (read only); w_fin: finalised pivot right-hand sides
// (written by the `first` tile of every front); cbuf: all child-update buffers.
void launch_forward_rows(const SolveTables& t, const SolveTile* tiles, int n_tiles, int smem_doubles, const double* w_in,
                         double* w_fin, double* cbuf, cudaStream_t s);
void launch_forward_cols(const SolveTables& t, const SolveTile* tiles, int n_tiles, const double* w_in, double* w_fin,
                         double* cbuf, cudaStream_t s);
// backward sweep of one level; smem_doubles == 0 selects the variant that gathers the vector on the fly
void launch_backward_rows(const SolveTables& t, const SolveTile* tiles, int n_tiles, int smem_doubles, const double* w_fin,
                          double* x_perm, cudaStream_t s);
// out[i] = in[index[i]]
void launch_gather(int n, const int* index, const double* in, double* out, cudaStream_t s);
// opt in to large dynamic shared memory once per process
void configure_solve_kernels(int max_smem_bytes);

} // namespace pecs
