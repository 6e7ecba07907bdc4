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
  void upload(const CsrMatrix& A);
  size_t bytes() const { return col.bytes() + val.bytes(); }
};

// y[i] = (base ? base[i] : 0) + sum_k A1(i,k) x1[k] - sum_k A2(i,k) x2[k]     (A1 may be null)
void launch_ell_combine(int n_rows, const double* base, const DeviceEll* A1, const double* x1, const DeviceEll& A2,
                        const double* x2, double* y, cudaStream_t s);

} // namespace pecs
