// rhs_kernels.cuh -- launch interface of the right-hand-side assembly kernels (rhs_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../rhs_math.hpp"

namespace pecs {

// one carrier subdomain as the kernels see it (structure of arrays, all device pointers)
struct DomainView {
  int n_cells;
  const double* vx;      // [4][n_cells] vertex x, lexicographic vertex a at vx[a*n + c]
  const double* vy;      // [4][n_cells]
  const int* rt_dof;     // [4][n_cells] Poisson flux dof of face f of the matched Poisson cell
  const int* phi_dof;    // [n_cells]    Poisson potential dof of the matched Poisson cell
  // boundary cells (cells with at least one boundary face)
  int n_bcells;
  const int* bcell;      // [n_bcells] cell index
  const int* bface_id;   // [n_bcells][4] boundary id of face f, -1 if the face is interior
  const int* bnb_cell;   // [n_bcells] matched cell of the other subdomain across the interface (or -1)
  const int* bnb_face;   // [n_bcells] its face number
  const int* brecord;    // [n_cells] index of the cell's boundary record, -1 for interior cells
  const double* bgeom;   // [n_bcells][4][4] {n_x, n_y, ds, tau/h} per face of a boundary cell (launch_boundary_geometry)
  // static cell integrals (production; launch_static_cell_integrals), NULL when absent
  const double* nodal_int; // [n_cells][4] int N_a
  const double* gen_int;   // [n_cells][4] int N_a G   (semiconductor under illumination only)
};

// everything one subdomain contributes to a fused launch; n_cells == 0: pass absent
struct CarrierPass {
  DomainView d;
  RhsParams p;
  int other_n_cells;       // cells of the other subdomain (offset of its density block)
  const double *u1, *u2;   // this subdomain's carriers (state)
  const double *o1, *o2;   // the other subdomain's carriers: traces across the interface
  double *rhs1, *rhs2;
};

// one-time: the static per-cell integrals the production kernels read (gen_int may be NULL: dark, or electrolyte)
void launch_static_cell_integrals(const DomainView& d, const RhsParams& p, double* nodal_int, double* gen_int,
                                  cudaStream_t s);
// one-time: normals, surface elements and penalty weights of the faces of the boundary cells
void launch_boundary_geometry(const DomainView& d, double tau, double* out, cudaStream_t s);
// which production carrier kernel launch_carrier_rhs uses: 0 point-by-point (v7), 1 (default) sum-factorised, one thread
// per cell, 2 sum-factorised streaming kernel (cp.async ring; measured slower at 2.2 cells per thread), 11/14/15 launch
// shapes of 1; PECS_B200_RHS_KERNEL overrides
int carrier_rhs_variant();

// rhs_c = M u_c + cell terms + boundary / interface / Schottky face terms for both carriers of both passes, one launch
// (SURVEY K1-K3); X = Poisson solution (electric field)
void launch_carrier_rhs(const CarrierPass& a, const CarrierPass& b, int kind, const double* X, cudaStream_t s);
// potential rows of the Poisson rhs: -int (doping + z1 rho1 + z2 rho2) per matched cell, both passes, one launch
void launch_poisson_cell_rhs(const CarrierPass& a, const CarrierPass& b, int kind, const double* static_rows, int n_static,
                             double* poisson_rhs, cudaStream_t s);

// Poisson boundary faces: time-independent, evaluated once into a static vector
struct PoissonFaceView {
  int n_faces;
  const int* cell;
  const int* face;
  const int* id;
  const int* cell_is_semiconductor; // [n_poisson_cells]
  const double* vx;                 // [4][n_poisson_cells]
  const double* vy;
  int n_poisson_cells;
  const int* face_dof;              // [n_poisson_cells][4]
  const int* constraint_master;     // [n_dofs] -2 unconstrained, -1 pinned to zero, else master dof
  const double* constraint_weight;  // [n_dofs]
};
struct PoissonFaceParams {
  int kind;
  double phi_bi, phi_app, phi_sch, sch_location;
};
void launch_poisson_face_rhs(const PoissonFaceView& v, const PoissonFaceParams& p, double* static_rhs, cudaStream_t s);

// ConstraintMatrix::distribute on the device
// I-V post-processing: per boundary record of the semiconductor {int k_et (rho_n - rho_n^e) rho_o, int k_ht (rho_p - rho_p^e) rho_r}
// over its interface face (zero without one) -> partial[2 n_bcells]
void launch_interface_currents(const CarrierPass& semiconductor, double* partial, cudaStream_t s);
void launch_distribute(int n_constraints, const int* dof, const int* master, const double* weight, double* x,
                       cudaStream_t s);

// algorithmic bytes per cell of the carrier kernel (both carriers): SURVEY section 8(d)
constexpr int64_t kCarrierRhsBytesPerCell = 368;
constexpr int64_t kPoissonRhsBytesPerCell = 140;

} // namespace pecs
