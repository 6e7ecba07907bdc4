// factor_device.cu -- see factor_device.cuh.  (device implementation follows; host factorisation until then)
#include "factor_device.cuh"

#include <cstdlib>

#include "../error.hpp"

namespace pecs {

bool device_factorization_enabled() { return false; }

void factorize_device(const SolvePlan&, const CsrMatrix&, const DeviceFront*, const int*, const int*, double*, double*) {
  throw StatusError(PECS_ERR_INTERNAL, "factorize_device: not built");
}

} // namespace pecs
