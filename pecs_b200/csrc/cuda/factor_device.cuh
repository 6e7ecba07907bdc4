// factor_device.cuh -- numeric multifrontal factorisation on the device (one-time setup).
#pragma once
#include "../host/Csr.hpp"
#include "../host/SparseDirect.hpp"
#include "solve_kernels.cuh"

namespace pecs {

// The device factorisation is the default; PECS_B200_HOST_FACTOR=1 selects the host reference implementation
// (host/SparseDirect.cpp::factorize_host) for debugging -- setup only, never the per-step path.
bool device_factorization_enabled();
// creates the current device's cuSOLVER / cuBLAS handles of factorize_device now (about a second, once per process and device)
void warm_factor_handles();

// Fills the forward / backward tables (already allocated on the device, zero-initialised inside) of `plan`.
void factorize_device(const SolvePlan& plan, const CsrMatrix& A, double* d_fwd, double* d_bwd);
// the same with P A P^T and its transpose (P = plan.perm) already formed: the two host permutations cost about as much
// as the device work, so pecs_ctx_create forms them on the concurrent preparation threads
void factorize_device(const SolvePlan& plan, const CsrMatrix& A_permuted, const CsrMatrix& A_permuted_transposed, double* d_fwd,
                      double* d_bwd);

} // namespace pecs
