// factor_device.cuh -- numeric multifrontal factorisation on the device (one-time setup).
#pragma once
#include "../host/Csr.hpp"
#include "../host/SparseDirect.hpp"
#include "solve_kernels.cuh"

namespace pecs {

// The device factorisation is the default; PECS_B200_HOST_FACTOR=1 selects the host reference implementation
// (host/SparseDirect.cpp::factorize_host) for debugging -- setup only, never the per-step path.
bool device_factorization_enabled();

// Fills the forward / backward tables (already allocated on the device, zero-initialised inside) of `plan`.
void factorize_device(const SolvePlan& plan, const CsrMatrix& A, double* d_fwd, double* d_bwd);

} // namespace pecs
