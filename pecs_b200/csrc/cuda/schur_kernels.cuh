// schur_kernels.cuh -- sparse mat-vecs around the reduced (density-only) carrier solve, see host/SchurReduction.hpp.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../host/Csr.hpp"
#include "device_util.cuh"

namespace pecs {

// ELLPACK copy of a CSR matrix: `width` entries per row, stored column-major (entry k of row i at k*n + i) so that
// one thread per row reads with unit stride across the warp.  Padding entries have value 0 and column 0.
struct DeviceEll {
  int n = 0, width = 0;
  DeviceBuffer<int> col;
  DeviceBuffer<double> val;
  // row_order (optional): ELL row i holds row (*row_order)[i] of A
  void upload(const CsrMatrix& A, const std::vector<int>* row_order = nullptr);
  size_t bytes() const { return col.bytes() + val.bytes(); }
};

// one sparse term of a combination: sign * A x (A == nullptr: term absent)
struct EllTerm {
  const DeviceEll* A = nullptr;
  const double* x = nullptr;
  double sign = 1.0;
};

// y[i] = (base ? base[base_index ? base_index[i] : i] : 0) + sum over the terms of sign * (A x)[i]
void launch_ell_combine(int n_rows, const double* base, const int* base_index, EllTerm t0, EllTerm t1, EllTerm t2, double* y,
                        cudaStream_t s);

} // namespace pecs
