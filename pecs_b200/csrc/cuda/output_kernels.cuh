// output_kernels.cuh -- device side of the output path (output_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace pecs {

// doubles per cell of the patch buffers: 4 patch vertices x (3-vector + scalar) per field
constexpr int kCarrierPatchDoubles = 32; // [current_1 12][density_1 4][current_2 12][density_2 4] blocks, see below
constexpr int kPoissonPatchDoubles = 16; // [field 12][potential 4]

// One DataOut patch per cell (4 vertices, deal.II lexicographic order), fields already rescaled like the reference's
// PostProcessor (source/PostProcessor.cpp:80-123): currents x scale_current, densities unscaled.  Block layout of
// `out` (n = cells): current_1 [4n][3] | density_1 [4n] | current_2 [4n][3] | density_2 [4n]  -- each block is exactly
// one VTU DataArray, so the host writes them without touching the numbers.
void launch_carrier_patches(int n_cells, const double* u1, const double* u2, double scale_current, double* out,
                            cudaStream_t s);
// RT0 x DGQ0: field(x_a) = J(x_a) [Xf0 (1-xi) + Xf1 xi, Xf2 (1-eta) + Xf3 eta] / det J(x_a) at the four vertices,
// x scale_field; potential x scale_potential.  Layout: field [4n][3] | potential [4n].
void launch_poisson_patches(int n_cells, const double* vx, const double* vy, const int* face_dof, int n_rt,
                            const double* X, double scale_field, double scale_potential, double* out, cudaStream_t s);

} // namespace pecs
