// solve_kernels.cuh -- device data layout and launch interface of the multifrontal solve sweeps.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pecs {

// one front as the kernels see it (see host/SparseDirect.hpp for the maths)
struct DeviceFront {
  int np, nb, p0;
  int has_child[2];
  long long bd_off;       // boundary index list (positions in the permuted vector)
  long long fwd_off;      // G          nb x np       row-major
  long long bwd_off;      // [Inv | -H] np x (np+nb)  row-major
  long long upd_off;      // this front's update vector
  long long cmap_off[2];  // inverse child maps, np+nb ints each
  long long child_upd_off[2];
};

// a unit of work of one level kernel: rows [row0, row0+nrows) of one front's table
struct SolveTile {
  int front, row0, nrows, first; // first != 0: this tile also publishes the finalised pivot rhs
};

struct SolveTables {
  const DeviceFront* fronts;
  const int* bd_index;
  const int* child_map;
  const double* fwd;
  const double* bwd;
};

constexpr int kSolveThreads = 256;
constexpr int kSolveRowsPerTile = 64;

// w_in: permuted right-hand side (read only); w_fin: finalised pivot right-hand sides (written);
// upd: update vectors of all fronts.
// vec2: every row of every front of the level starts on a 16-byte boundary and has an even length
void launch_forward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int smem_doubles, bool vec2,
                          const double* w_in, double* w_fin, double* upd, cudaStream_t s);
void launch_backward_level(const SolveTables& t, const SolveTile* tiles, int n_tiles, int smem_doubles, bool vec2,
                           const double* w_fin, double* x_perm, cudaStream_t s);
// out[p] = in[iperm[p]]  /  out[i] = in[perm[i]]
void launch_gather(int n, const int* index, const double* in, double* out, cudaStream_t s);
// opt in to large dynamic shared memory once per process
void configure_solve_kernels(int max_smem_bytes);

} // namespace pecs
