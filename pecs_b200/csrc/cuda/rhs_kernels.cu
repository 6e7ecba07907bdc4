// rhs_kernels.cu -- per-step right-hand-side assembly on the device (sm_100a).
//
// Replaces the WorkStream passes of the reference's time loop:
//   carrier_rhs           <- ONE launch for both subdomains:
//       cell terms        <- mass_matrix.vmult x2 + the cell q-loop of assemble_local_{semiconductor,electrolyte}_rhs
//                            (reference source/SolarCell.cpp:1043-1047, 1146-1193, 1423-1427, 1519-1553)
//       boundary terms    <- the boundary-face branches Dirichlet / Interface / Schottky
//                            (reference source/SolarCell.cpp:1197-1412, 1557-1725), by the thread that owns the cell
//   poisson_cell_rhs      <- the cell loops of assemble_local_Poisson_rhs_for_{semiconductor,electrolyte}
//                            (reference source/SolarCell.cpp:551-578, 741-762)
//   poisson_face_rhs      <- their Dirichlet / Schottky face loops (reference source/SolarCell.cpp:583-683, 768-814),
//                            time independent, hence evaluated once
// plus the manufactured-solution variants (reference source/LDG.cpp:681-982, source/SolarCell.cpp:2105-2330).
//
// Design: one thread per cell, everything in registers.  Nothing per-quadrature-point is loaded from memory:
// the reference-cell basis values are compile-time constants after unrolling and the Jacobian is recomputed from
// the four vertices (8 doubles per cell, stored as structure of arrays so that a warp reads 256 contiguous bytes
// per vertex coordinate).  M u is fused into the cell integral (sum_q N_a rho(x_q) JxW / dt) so no mass matrix is
// read, and det J cancels in the drift term (JxW * D_h = w_q J Dhat), so the hot loop has no division.
// DG test functions live on one cell: every thread owns its 24 output rows, no atomics, fixed summation order.
// Algorithmic traffic 368 B per cell (both carriers); the kernel is HBM-bound.
#include "rhs_kernels.cuh"

#include "../../../include/pecs_b200.h"
#include "../fe.hpp"
#include "../test_functions.hpp"

namespace pecs {

namespace {

constexpr int kThreads = 128;

struct CarrierPassPair {
  CarrierPass pass[2];
};

__device__ __forceinline__ fe::CellVerts load_verts(const DomainView& d, int c) {
  fe::CellVerts v;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    v.x[a] = __ldg(d.vx + (size_t)a * d.n_cells + c);
    v.y[a] = __ldg(d.vy + (size_t)a * d.n_cells + c);
  }
  return v;
}

__device__ __forceinline__ void load4(const double* p, double out[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  out[0] = a.x;
  out[1] = a.y;
  out[2] = b.x;
  out[3] = b.y;
}
__device__ __forceinline__ void store4(double* p, const double v[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
__device__ __forceinline__ void add4(double* p, const double v[4]) {
  double t[4];
  load4(p, t);
#pragma unroll
  for (int a = 0; a < 4; ++a) t[a] += v[a];
  store4(p, t);
}

// ------------------------------------------------------------------------------------------ carrier cell terms
template <int KIND>
__device__ __forceinline__ void carrier_cell_terms(const DomainView& d, const RhsParams& p, int c,
                                                   const double* __restrict__ u1, const double* __restrict__ u2,
                                                   const double* __restrict__ X, double* __restrict__ rhs1,
                                                   double* __restrict__ rhs2) {
  const size_t n = (size_t)d.n_cells;
  constexpr bool kProduction = KIND == PECS_KIND_PRODUCTION;
  constexpr bool kDrift = KIND != PECS_KIND_TEST_STEADY;
  constexpr bool kPoissonField = KIND == PECS_KIND_PRODUCTION || KIND == PECS_KIND_TEST_DD_POISSON;

  const fe::CellVerts v = load_verts(d, c);
  double r1[4] = {0, 0, 0, 0}, r2[4] = {0, 0, 0, 0}, Xf[4] = {0, 0, 0, 0};
  if (KIND != PECS_KIND_TEST_STEADY) load4(u1 + 8 * n + 4 * (size_t)c, r1);
  if (kProduction) load4(u2 + 8 * n + 4 * (size_t)c, r2);
  if (kPoissonField) {
#pragma unroll
    for (int f = 0; f < 4; ++f) Xf[f] = __ldg(X + __ldg(d.rt_dof + (size_t)f * n + c));
  }

  double jx1[4] = {0, 0, 0, 0}, jy1[4] = {0, 0, 0, 0}, rh1[4] = {0, 0, 0, 0};
  double jx2[4] = {0, 0, 0, 0}, jy2[4] = {0, 0, 0, 0}, rh2[4] = {0, 0, 0, 0};

#pragma unroll
  for (int qy = 0; qy < 3; ++qy) {
    const double eta = fe::gauss_x(qy), wy = fe::gauss_w(qy);
#pragma unroll
    for (int qx = 0; qx < 3; ++qx) {
      const double xi = fe::gauss_x(qx), w = fe::gauss_w(qx) * wy;
      const fe::Jac j = fe::jacobian(v, xi, eta);
      double N[4];
      fe::shape(xi, eta, N);
      const double JxW = j.det * w;
      const double rho1 = N[0] * r1[0] + N[1] * r1[1] + N[2] * r1[2] + N[3] * r1[3];
      double rho2 = 0.0;
      if (kProduction) rho2 = N[0] * r2[0] + N[1] * r2[1] + N[2] * r2[2] + N[3] * r2[3];

      // source on the density rows; the mass term of M u^{k-1} rides along
      double src1, src2 = 0.0;
      if (kProduction) {
        double gen = 0.0;
        if (p.gen_scale != 0.0) { // Generation::value, reference Generation.cpp:29-44 (semiconductor only)
          const double y = v.y[0] * N[0] + v.y[1] * N[1] + v.y[2] * N[2] + v.y[3] * N[3];
          gen = p.gen_scale * exp(p.gen_alpha * (y - p.gen_location));
        }
        // SRH_Recombination == 0.0 (reference SolarCell.hpp:86-98)
        src1 = (rho1 * p.inv_dt + gen) * JxW;
        src2 = (rho2 * p.inv_dt + gen) * JxW;
      } else {
        double x, y;
        fe::map_point(v, xi, eta, x, y);
        double f;
        if (KIND == PECS_KIND_TEST_STEADY)
          f = testfn::poisson_rhs(x, y);
        else if (KIND == PECS_KIND_TEST_TRANSIENT)
          f = testfn::ldg_rhs(x, y, p.time) + rho1 * p.inv_dt;
        else
          f = testfn::dd_rhs(x, y, p.time) + rho1 * p.inv_dt;
        src1 = f * JxW;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        rh1[a] += N[a] * src1;
        if (kProduction) rh2[a] += N[a] * src2;
      }

      if (kDrift) {
        // JxW * E = w * J * Dhat / eps  (det J cancels: contravariant Piola map of RT0)
        double Dx, Dy;
        if (kPoissonField) {
          const double dhx = Xf[0] * (1.0 - xi) + Xf[1] * xi;
          const double dhy = Xf[2] * (1.0 - eta) + Xf[3] * eta;
          Dx = (j.xxi * dhx + j.xeta * dhy) * w;
          Dy = (j.yxi * dhx + j.yeta * dhy) * w;
        } else { // test_transient: fixed field (1, 0)
          Dx = JxW;
          Dy = 0.0;
        }
        const double s1 = (kProduction ? p.charge1 * p.inv_eps : -1.0) * rho1;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          jx1[a] += N[a] * (s1 * Dx);
          jy1[a] += N[a] * (s1 * Dy);
        }
        if (kProduction) {
          const double s2 = p.charge2 * p.inv_eps * rho2;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            jx2[a] += N[a] * (s2 * Dx);
            jy2[a] += N[a] * (s2 * Dy);
          }
        }
      }
    }
  }
  store4(rhs1 + 4 * (size_t)c, jx1);
  store4(rhs1 + 4 * n + 4 * (size_t)c, jy1);
  store4(rhs1 + 8 * n + 4 * (size_t)c, rh1);
  if (kProduction) {
    store4(rhs2 + 4 * (size_t)c, jx2);
    store4(rhs2 + 4 * n + 4 * (size_t)c, jy2);
    store4(rhs2 + 8 * n + 4 * (size_t)c, rh2);
  }
}

// ------------------------------------------------------------------------------------------ boundary faces
// trace of the density of a cell at a face point
__device__ __forceinline__ double trace(const double N[4], const double r[4]) {
  return N[0] * r[0] + N[1] * r[1] + N[2] * r[2] + N[3] * r[3];
}

// Face terms of boundary cell record r, added to what the cell terms of the same thread have just stored.  Kept out
// of line: 1.5 % of the cells take this path and its registers must not burden the other 98.5 %.
template <int KIND>
__device__ __noinline__ void carrier_boundary_terms(const DomainView& d, int other_n_cells, const RhsParams& p, int r,
                                                    const double* __restrict__ u1, const double* __restrict__ u2,
                                                    const double* __restrict__ o1, const double* __restrict__ o2, double* rhs1,
                                                    double* rhs2) {
  constexpr bool kProduction = KIND == PECS_KIND_PRODUCTION;
  const int c = d.bcell[r];
  const size_t n = (size_t)d.n_cells;
  const fe::CellVerts v = load_verts(d, c);
  const double h = fe::cell_diameter(v);
  double r1[4], r2[4] = {0, 0, 0, 0};
  load4(u1 + 8 * n + 4 * (size_t)c, r1);
  if (kProduction) load4(u2 + 8 * n + 4 * (size_t)c, r2);

  double jx1[4] = {0, 0, 0, 0}, jy1[4] = {0, 0, 0, 0}, rh1[4] = {0, 0, 0, 0};
  double jx2[4] = {0, 0, 0, 0}, jy2[4] = {0, 0, 0, 0}, rh2[4] = {0, 0, 0, 0};

  for (int f = 0; f < 4; ++f) {
    const int id = d.bface_id[4 * r + f];
    if (id < 0 || id == PECS_NEUMANN) continue; // interior face, or insulating: nothing to do
    // the other subdomain's traces on an interface face (same q index on both sides)
    double q1[4] = {0, 0, 0, 0}, q2[4] = {0, 0, 0, 0};
    int nb_face = 0;
    if (kProduction && id == PECS_INTERFACE) {
      const int nc = d.bnb_cell[r];
      nb_face = d.bnb_face[r];
      load4(o1 + 8 * (size_t)other_n_cells + 4 * (size_t)nc, q1);
      load4(o2 + 8 * (size_t)other_n_cells + 4 * (size_t)nc, q2);
    }
    for (int q = 0; q < 3; ++q) {
      const double t = fe::gauss_x(q);
      double xi, eta, nx, ny, ds, N[4];
      fe::face_point(f, t, xi, eta);
      const fe::Jac j = fe::jacobian(v, xi, eta);
      fe::face_normal_ds(j, f, nx, ny, ds);
      fe::shape(xi, eta, N);
      const double W = ds * fe::gauss_w(q);
      if (id == PECS_DIRICHLET) {
        // int ( -p.n + (tau/h) v ) u_D
        double bc1, bc2 = 0.0;
        if (kProduction) {
          bc1 = p.rho1_e;
          bc2 = p.rho2_e;
        } else {
          double x, y;
          fe::map_point(v, xi, eta, x, y);
          bc1 = (KIND == PECS_KIND_TEST_STEADY) ? testfn::poisson_bc(x, y) : testfn::density(x, y, p.time);
        }
        const double pen = p.tau / h;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          jx1[a] += -N[a] * nx * bc1 * W;
          jy1[a] += -N[a] * ny * bc1 * W;
          rh1[a] += pen * N[a] * bc1 * W;
          if (kProduction) {
            jx2[a] += -N[a] * nx * bc2 * W;
            jy2[a] += -N[a] * ny * bc2 * W;
            rh2[a] += pen * N[a] * bc2 * W;
          }
        }
      } else if (id == PECS_INTERFACE) {
        if (kProduction) {
          double xin, etan, Nn[4];
          fe::face_point(nb_face, t, xin, etan);
          fe::shape(xin, etan, Nn);
          if (p.is_semiconductor) {
            // -v k_et (rho_n - rho_n^e) rho_o -> electrons ; +v k_ht (rho_p - rho_p^e) rho_r -> holes
            const double e = -p.k_et * (trace(N, r1) - p.rho1_e) * trace(Nn, q2) * W;
            const double hl = p.k_ht * (trace(N, r2) - p.rho2_e) * trace(Nn, q1) * W;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              rh1[a] += N[a] * e;
              rh2[a] += N[a] * hl;
            }
          } else {
            // current = -k_et (rho_n - rho_n^e) rho_o + k_ht (rho_p - rho_p^e) rho_r ; reductants += , oxidants -=
            const double cur = (-p.k_et * (trace(Nn, q1) - p.other1_e) * trace(N, r2) +
                                p.k_ht * (trace(Nn, q2) - p.other2_e) * trace(N, r1)) * W;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              rh1[a] += N[a] * cur;
              rh2[a] -= N[a] * cur;
            }
          }
        } else if (KIND == PECS_KIND_TEST_TRANSIENT) {
          double x, y;
          fe::map_point(v, xi, eta, x, y);
          const double g = -testfn::ldg_interface(x, y, p.time) * W;
#pragma unroll
          for (int a = 0; a < 4; ++a) rh1[a] += N[a] * g;
        }
      } else if (id == PECS_SCHOTTKY) {
        if (kProduction && p.is_semiconductor) {
          const double e = -p.v_n * (trace(N, r1) - p.rho1_e) * W;
          const double hl = p.v_p * (trace(N, r2) - p.rho2_e) * W;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            rh1[a] += N[a] * e;
            rh2[a] += N[a] * hl;
          }
        }
      }
    }
  }
  add4(rhs1 + 4 * (size_t)c, jx1);
  add4(rhs1 + 4 * n + 4 * (size_t)c, jy1);
  add4(rhs1 + 8 * n + 4 * (size_t)c, rh1);
  if (kProduction) {
    add4(rhs2 + 4 * (size_t)c, jx2);
    add4(rhs2 + 4 * n + 4 * (size_t)c, jy2);
    add4(rhs2 + 8 * n + 4 * (size_t)c, rh2);
  }
}

// ONE launch assembles the carrier right-hand sides of BOTH subdomains: blocks [0, blocks_a) work on pass a, the rest
// on pass b (a pass with n_cells == 0 is absent); every thread does the cell terms of its cell and, if the cell has
// boundary faces, their terms right after.
template <int KIND>
__global__ void __launch_bounds__(kThreads, 4) carrier_rhs_kernel(const __grid_constant__ CarrierPassPair pp, int blocks_a,
                                                               const double* __restrict__ X) {
  const bool first = (int)blockIdx.x < blocks_a;
  const CarrierPass& w = pp.pass[first ? 0 : 1]; // stays in the constant bank: no local copy
  const int c = ((int)blockIdx.x - (first ? 0 : blocks_a)) * blockDim.x + threadIdx.x;
  if (c >= w.d.n_cells) return;
  carrier_cell_terms<KIND>(w.d, w.p, c, w.u1, w.u2, X, w.rhs1, w.rhs2);
  const int r = w.d.brecord ? w.d.brecord[c] : -1;
  if (r >= 0) carrier_boundary_terms<KIND>(w.d, w.other_n_cells, w.p, r, w.u1, w.u2, w.o1, w.o2, w.rhs1, w.rhs2);
}

// ------------------------------------------------------------------------------------------ Poisson cells
template <int KIND>
__global__ void __launch_bounds__(kThreads) poisson_cell_rhs_kernel(const __grid_constant__ CarrierPassPair pp, int blocks_a,
                                                                    double* __restrict__ poisson_rhs) {
  const bool first = (int)blockIdx.x < blocks_a;
  const CarrierPass& w = pp.pass[first ? 0 : 1];
  const DomainView& d = w.d;
  const RhsParams& p = w.p;
  const double* __restrict__ u1 = w.u1;
  const double* __restrict__ u2 = w.u2;
  const int c = ((int)blockIdx.x - (first ? 0 : blocks_a)) * blockDim.x + threadIdx.x;
  if (c >= d.n_cells) return;
  const size_t n = (size_t)d.n_cells;
  const fe::CellVerts v = load_verts(d, c);
  double r1[4] = {0, 0, 0, 0}, r2[4] = {0, 0, 0, 0};
  if (KIND != PECS_KIND_TEST_STEADY) load4(u1 + 8 * n + 4 * (size_t)c, r1);
  if (KIND == PECS_KIND_PRODUCTION) load4(u2 + 8 * n + 4 * (size_t)c, r2);
  double acc = 0.0;
#pragma unroll
  for (int qy = 0; qy < 3; ++qy)
#pragma unroll
    for (int qx = 0; qx < 3; ++qx) {
      const double xi = fe::gauss_x(qx), eta = fe::gauss_x(qy);
      const double w = fe::gauss_w(qx) * fe::gauss_w(qy);
      const fe::Jac j = fe::jacobian(v, xi, eta);
      double N[4];
      fe::shape(xi, eta, N);
      double f;
      if (KIND == PECS_KIND_PRODUCTION) {
        f = p.doping + (p.charge1 * trace(N, r1) + p.charge2 * trace(N, r2));
      } else {
        double x, y;
        fe::map_point(v, xi, eta, x, y);
        f = (KIND == PECS_KIND_TEST_STEADY) ? testfn::poisson_rhs(x, y)
                                            : (testfn::dd_poisson_rhs(x, y, p.time) - trace(N, r1));
      }
      acc += -f * (j.det * w);
    }
  poisson_rhs[d.phi_dof[c]] = acc;
}

// ------------------------------------------------------------------------------------------ Poisson faces (static)
__global__ void poisson_face_rhs_kernel(PoissonFaceView fv, PoissonFaceParams p, double* static_rhs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= fv.n_faces) return;
  const int id = fv.id[k], c = fv.cell[k], f = fv.face[k];
  const bool semi = fv.cell_is_semiconductor[c] != 0;
  const bool production = p.kind == PECS_KIND_PRODUCTION;
  // which faces carry Dirichlet data: reference SolarCell.cpp:595-681 (semiconductor: Dirichlet and Schottky),
  // :778-813 (electrolyte: Dirichlet with Bulk_Bias == 0); tests: Dirichlet with the manufactured potential
  if (!(id == PECS_DIRICHLET || (production && semi && id == PECS_SCHOTTKY))) return;
  fe::CellVerts v;
  for (int a = 0; a < 4; ++a) {
    v.x[a] = fv.vx[(size_t)a * fv.n_poisson_cells + c];
    v.y[a] = fv.vy[(size_t)a * fv.n_poisson_cells + c];
  }
  // (psi_f . n) ds = (psihat_f . nhat) dt = +-1 dt on the face's own edge, 0 for the other three functions
  const double sign = (f & 1) ? 1.0 : -1.0;
  double acc = 0.0;
  for (int q = 0; q < 3; ++q) {
    double xi, eta, x, y;
    fe::face_point(f, fe::gauss_x(q), xi, eta);
    fe::map_point(v, xi, eta, x, y);
    double value;
    if (production) {
      if (semi) {
        const double bi = (x == 0.0) ? p.phi_bi : 0.0; // Built_In_Bias, reference BiasValues.cpp:11-31
        const double bc = (id == PECS_DIRICHLET) ? ((x == 0.0) ? p.phi_app : 0.0)          // Applied_Bias :75-94
                                                 : ((y == p.sch_location) ? p.phi_sch : 0.0); // Schottky_Bias :49-65
        value = bi - bc;
      } else {
        value = 0.0; // Bulk_Bias, reference BiasValues.cpp:99-115
      }
    } else {
      value = testfn::poisson_bc(x, y);
    }
    acc += -sign * value * fe::gauss_w(q);
  }
  int dof = fv.face_dof[4 * c + f];
  const int master = fv.constraint_master[dof];
  if (master == -1) return; // pinned to zero: contribution dropped
  if (master >= 0) {
    acc *= fv.constraint_weight[dof];
    dof = master;
  }
  atomicAdd(static_rhs + dof, acc);
}

__global__ void distribute_kernel(int n, const int* __restrict__ dof, const int* __restrict__ master,
                                  const double* __restrict__ weight, double* x) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int m = master[k];
  x[dof[k]] = m >= 0 ? weight[k] * x[m] : 0.0;
}

inline int blocks_for(int n) { return (n + kThreads - 1) / kThreads; }

} // namespace

#define PECS_DISPATCH_KIND(kind, CALL)                                   \
  switch (kind) {                                                        \
    case PECS_KIND_PRODUCTION: { CALL(PECS_KIND_PRODUCTION); break; }    \
    case PECS_KIND_TEST_STEADY: { CALL(PECS_KIND_TEST_STEADY); break; }  \
    case PECS_KIND_TEST_TRANSIENT: { CALL(PECS_KIND_TEST_TRANSIENT); break; } \
    default: { CALL(PECS_KIND_TEST_DD_POISSON); break; }                 \
  }

void launch_carrier_rhs(const CarrierPass& a, const CarrierPass& b, int kind, const double* X, cudaStream_t s) {
  const int blocks_a = blocks_for(a.d.n_cells), blocks_b = blocks_for(b.d.n_cells);
  if (blocks_a + blocks_b == 0) return;
  const CarrierPassPair pp{{a, b}};
#define CALL(K) carrier_rhs_kernel<K><<<blocks_a + blocks_b, kThreads, 0, s>>>(pp, blocks_a, X)
  PECS_DISPATCH_KIND(kind, CALL)
#undef CALL
}

void launch_poisson_cell_rhs(const CarrierPass& a, const CarrierPass& b, int kind, double* poisson_rhs, cudaStream_t s) {
  const int blocks_a = blocks_for(a.d.n_cells), blocks_b = blocks_for(b.d.n_cells);
  if (blocks_a + blocks_b == 0) return;
  const CarrierPassPair pp{{a, b}};
#define CALL(K) poisson_cell_rhs_kernel<K><<<blocks_a + blocks_b, kThreads, 0, s>>>(pp, blocks_a, poisson_rhs)
  PECS_DISPATCH_KIND(kind, CALL)
#undef CALL
}

void launch_poisson_face_rhs(const PoissonFaceView& v, const PoissonFaceParams& p, double* static_rhs, cudaStream_t s) {
  if (v.n_faces == 0) return;
  poisson_face_rhs_kernel<<<blocks_for(v.n_faces), kThreads, 0, s>>>(v, p, static_rhs);
}

void launch_distribute(int n_constraints, const int* dof, const int* master, const double* weight, double* x,
                       cudaStream_t s) {
  if (n_constraints == 0) return;
  distribute_kernel<<<blocks_for(n_constraints), kThreads, 0, s>>>(n_constraints, dof, master, weight, x);
}

} // namespace pecs
