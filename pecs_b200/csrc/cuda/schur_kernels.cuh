// schur_kernels.cuh -- sparse mat-vecs around the reduced (density-only) carrier solve, see host/SchurReduction.hpp.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../host/Csr.hpp"
#include "device_util.cuh"

namespace pecs {

// ELLPACK copy of a CSR matrix: `width` slots per row, stored column-major (slot k of row i at k*n + i) so that
// one thread per row reads with unit stride across the warp.  Padding entries have value 0 and column 0.
// block == 1: a slot is one entry (8 B value + 4 B column).  block == 4: a slot is four entries in four consecutive,
// 4-aligned columns (32 B of values + ONE 4 B column): the LDG matrices couple whole cells (4 nodal values), so their
// entries come in such groups and the column indices shrink to a quarter -- 9 instead of 12 B per entry.  upload()
// picks whichever is smaller for the matrix at hand.
struct DeviceEll {
  int n = 0, width = 0, block = 1;
  DeviceBuffer<int> col;
  DeviceBuffer<double> val;
  // row_order (optional): ELL row i holds row (*row_order)[i] of A
  void upload(const CsrMatrix& A, const std::vector<int>* row_order = nullptr);
  void upload(const HostEll& table); // built ahead (the preparation threads of pecs_ctx_create)
  size_t bytes() const { return col.bytes() + val.bytes(); }
};

// one sparse term of a combination: sign * A x (A == nullptr: term absent)
struct EllTerm {
  const DeviceEll* A = nullptr;
  const double* x = nullptr;
  double sign = 1.0;
  const double* x2 = nullptr; // second vector of launch_ell_combine2
};
struct EllBase {
  const double* p[2];
};
struct EllOut {
  double* p[2];
};

// y[i] = (base ? base[base_index ? base_index[i] : i] : 0) + sum over the terms of sign * (A x)[i]
void launch_ell_combine(int n_rows, const double* base, const int* base_index, EllTerm t0, EllTerm t1, EllTerm t2, double* y,
                        cudaStream_t s);

// the same for two vectors at once: y = base + sum sign * A x and y2 = base2 + sum sign * A x2 with ONE pass over every
// table (two carriers with identical matrices); per vector the arithmetic and its order are those of launch_ell_combine
void launch_ell_combine2(int n_rows, const double* base, const double* base2, const int* base_index, EllTerm t0, EllTerm t1,
                         EllTerm t2, double* y, double* y2, cudaStream_t s);

} // namespace pecs
