// output_kernels.cu -- the output path on the device (sm_100a).
//
// Replaces what the reference's print_results does per time stamp on the host (source/SolarCell.cpp:1826-1858):
// DataOut::build_patches evaluating every solution at the patch vertices through FEValues, and PostProcessor's
// compute_derived_quantities_vector rescaling them to physical units (source/PostProcessor.cpp:80-123).  Here one
// kernel per vector writes the rescaled patch values straight in the layout of the VTU data arrays; the context then
// copies them to pinned host memory on its output stream while the time loop goes on (cuda/context.cu,
// pecs_output_snapshot), and a host thread writes the files (host/Output.cpp).
#include "output_kernels.cuh"

#include "../fe.hpp"
#include "../rhs_math.hpp"

namespace pecs {

namespace {

constexpr int kThreads = 128;

// DGQ1 support points are the cell vertices in lexicographic order: the patch value at vertex a IS nodal value a.
__global__ void carrier_patch_kernel(int n, const double* u1, const double* u2, double scale_current,
                                     double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const size_t N = (size_t)n;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double* u = k == 0 ? u1 : u2;
    double* current = out + (size_t)k * 16 * N;
    double* density = current + 12 * N;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const size_t dof = 4 * (size_t)c + a;
      current[3 * dof + 0] = scale_current * u[dof];
      current[3 * dof + 1] = scale_current * u[4 * N + dof];
      current[3 * dof + 2] = 0.0;
      density[dof] = u[8 * N + dof];
    }
  }
}

__global__ void poisson_patch_kernel(int n, const double* __restrict__ vx, const double* __restrict__ vy,
                                     const int* __restrict__ face_dof, int n_rt, const double* X,
                                     double scale_field, double scale_potential, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const size_t N = (size_t)n;
  fe::CellVerts v;
  double Xf[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    v.x[a] = vx[(size_t)a * N + c];
    v.y[a] = vy[(size_t)a * N + c];
    Xf[a] = X[face_dof[4 * (size_t)c + a]];
  }
  const double potential = scale_potential * X[n_rt + c];
  double* field = out;
  double* pot = out + 12 * N;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const size_t p = 4 * (size_t)c + a;
    double fx, fy;
    rhsmath::rt0_field_at_vertex(v, Xf, a, scale_field, fx, fy);
    field[3 * p + 0] = fx;
    field[3 * p + 1] = fy;
    field[3 * p + 2] = 0.0;
    pot[p] = potential;
  }
}

} // namespace

void launch_carrier_patches(int n_cells, const double* u1, const double* u2, double scale_current, double* out,
                            cudaStream_t s) {
  if (n_cells == 0) return;
  carrier_patch_kernel<<<(n_cells + kThreads - 1) / kThreads, kThreads, 0, s>>>(n_cells, u1, u2, scale_current, out);
}

void launch_poisson_patches(int n_cells, const double* vx, const double* vy, const int* face_dof, int n_rt,
                            const double* X, double scale_field, double scale_potential, double* out, cudaStream_t s) {
  if (n_cells == 0) return;
  poisson_patch_kernel<<<(n_cells + kThreads - 1) / kThreads, kThreads, 0, s>>>(n_cells, vx, vy, face_dof, n_rt, X,
                                                                              scale_field, scale_potential, out);
}

} // namespace pecs
