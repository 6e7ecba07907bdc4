// ell_tables.cu -- host side of schur_kernels.cuh: the ELL layout of a fixed CSR matrix, built once at setup and uploaded.
// (Kept apart from the kernels: profiles/r02_solve_traffic.json is tied to a hash of the kernel sources.)
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "schur_kernels.cuh"

namespace pecs {

void DeviceEll::upload(const CsrMatrix& A, const std::vector<int>* row_order) {
  n = A.n;
  // slots per row in both formats
  int w1 = 0, w4 = 0;
#pragma omp parallel for schedule(static) reduction(max : w1, w4)
  for (int i = 0; i < n; ++i) {
    w1 = std::max(w1, A.row_ptr[i + 1] - A.row_ptr[i]);
    int groups = 0, last = -1;
    for (int k = A.row_ptr[i]; k < A.row_ptr[i + 1]; ++k)
      if (A.col[k] / 4 != last) {
        last = A.col[k] / 4;
        ++groups;
      }
    w4 = std::max(w4, groups);
  }
  const bool forced_scalar = std::getenv("PECS_B200_ELL_SCALAR") != nullptr;
  block = (!forced_scalar && A.n % 4 == 0 && (size_t)w4 * 36 < (size_t)w1 * 12) ? 4 : 1;
  width = block == 4 ? w4 : w1;
  std::vector<int> c((size_t)n * width, 0);
  std::vector<double> v((size_t)n * width * block, 0.0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) { // row i owns entry i of every slot: no two rows write the same place
    const int r = row_order ? (*row_order)[i] : i;
    if (block == 1) {
      for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; ++k) {
        const size_t slot = (size_t)(k - A.row_ptr[r]) * n + i;
        c[slot] = A.col[k];
        v[slot] = A.val[k];
      }
    } else {
      int g = -1, last = -1;
      for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; ++k) { // columns are sorted within a row
        if (A.col[k] / 4 != last) {
          last = A.col[k] / 4;
          ++g;
          c[(size_t)g * n + i] = 4 * last;
        }
        v[((size_t)g * 4 + A.col[k] % 4) * n + i] = A.val[k];
      }
    }
  }
  col.upload(c);
  val.upload(v);
}

} // namespace pecs
