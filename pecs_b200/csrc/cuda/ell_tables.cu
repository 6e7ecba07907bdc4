// ell_tables.cu -- host side of schur_kernels.cuh: the ELL layout of a fixed CSR matrix (host/EllTable.cpp) goes to the
// device.  (Kept apart from the kernels: profiles/r02_solve_traffic.json is tied to a hash of the kernel sources.)
#include <omp.h>

#include "schur_kernels.cuh"

namespace pecs {

void DeviceEll::upload(const HostEll& E) {
  n = E.n;
  width = E.width;
  block = E.block;
  col.upload(E.col);
  val.upload(E.val);
}

void DeviceEll::upload(const CsrMatrix& A, const std::vector<int>* row_order) {
  upload(build_ell(A, row_order, omp_get_max_threads()));
}

} // namespace pecs
