// EllTable.cpp -- see HostEll in Csr.hpp.
#include <algorithm>
#include <cstdlib>

#include "Csr.hpp"

namespace pecs {

HostEll build_ell(const CsrMatrix& A, const std::vector<int>* row_order, int threads) {
  HostEll E;
  const int n = E.n = A.n;
  threads = std::max(1, threads);
  // slots per row in both formats
  int w1 = 0, w4 = 0;
#pragma omp parallel for schedule(static) reduction(max : w1, w4) num_threads(threads)
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
  E.block = (!forced_scalar && A.n % 4 == 0 && (size_t)w4 * 36 < (size_t)w1 * 12) ? 4 : 1;
  E.width = E.block == 4 ? w4 : w1;
  E.col.assign((size_t)n * E.width, 0);
  E.val.assign((size_t)n * E.width * E.block, 0.0);
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int i = 0; i < n; ++i) { // row i owns entry i of every slot: no two rows write the same place
    const int r = row_order ? (*row_order)[i] : i;
    if (E.block == 1) {
      for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; ++k) {
        const size_t slot = (size_t)(k - A.row_ptr[r]) * n + i;
        E.col[slot] = A.col[k];
        E.val[slot] = A.val[k];
      }
    } else {
      int g = -1, last = -1;
      for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; ++k) { // columns are sorted within a row
        if (A.col[k] / 4 != last) {
          last = A.col[k] / 4;
          ++g;
          E.col[(size_t)g * n + i] = 4 * last;
        }
        E.val[((size_t)g * 4 + A.col[k] % 4) * n + i] = A.val[k];
      }
    }
  }
  return E;
}

} // namespace pecs
