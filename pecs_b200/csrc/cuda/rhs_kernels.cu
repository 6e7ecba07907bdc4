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

#include <algorithm>
#include <cstdlib>

#include "device_util.cuh"

#include "../../../include/pecs_b200.h"
#include "../fe.hpp"
#include "../rhs_math.hpp"
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
// 4-vector of a step-varying vector (state): coherent loads, see device_util.cuh: ld_step
__device__ __forceinline__ void load4_step(const double* p, double out[4]) {
  const double2 a = ld_vec(reinterpret_cast<const double2*>(p));
  const double2 b = ld_vec(reinterpret_cast<const double2*>(p + 2));
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
                                                   const double* u1, const double* u2, const double* X, double* rhs1,
                                                   double* rhs2) {
  const size_t n = (size_t)d.n_cells;
  constexpr bool kProduction = KIND == PECS_KIND_PRODUCTION;
  constexpr bool kDrift = KIND != PECS_KIND_TEST_STEADY;
  constexpr bool kPoissonField = KIND == PECS_KIND_PRODUCTION || KIND == PECS_KIND_TEST_DD_POISSON;

  const fe::CellVerts v = load_verts(d, c);
  double r1[4] = {0, 0, 0, 0}, r2[4] = {0, 0, 0, 0}, Xf[4] = {0, 0, 0, 0};
  if (KIND != PECS_KIND_TEST_STEADY) load4_step(u1 + 8 * n + 4 * (size_t)c, r1);
  if (kProduction) load4_step(u2 + 8 * n + 4 * (size_t)c, r2);
  if (kPoissonField) {
#pragma unroll
    for (int f = 0; f < 4; ++f) Xf[f] = ld_vec(X + __ldg(d.rt_dof + (size_t)f * n + c));
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
        // SRH_Recombination (reference SolarCell.hpp:86-98): 0.0 unless switched on
        const double R = p.srh ? rhsmath::srh_recombination(rho1, rho2, p.n_i, p.tau_n, p.tau_p) : 0.0;
        src1 = (rho1 * p.inv_dt + gen + R) * JxW;
        src2 = (rho2 * p.inv_dt + gen + R) * JxW;
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
// Face terms of boundary cell record r, added to what the cell terms of the same thread have just stored.  Kept out
// of line: 1.5 % of the cells take this path and its registers must not burden the other 98.5 %.
__device__ __forceinline__ rhsmath::BoundaryRecord load_record(const DomainView& d, int r) {
  rhsmath::BoundaryRecord b;
  const int4 ids = __ldg(reinterpret_cast<const int4*>(d.bface_id) + r);
  b.id[0] = ids.x, b.id[1] = ids.y, b.id[2] = ids.z, b.id[3] = ids.w;
  b.nb_cell = __ldg(d.bnb_cell + r);
  b.nb_face = __ldg(d.bnb_face + r);
  return b;
}

__device__ __forceinline__ void load_geometry(const DomainView& d, int r, double geom[4][4]) {
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(d.bgeom + 16 * (size_t)r + 4 * f));
    const double2 b = __ldg(reinterpret_cast<const double2*>(d.bgeom + 16 * (size_t)r + 4 * f + 2));
    geom[f][0] = a.x, geom[f][1] = a.y, geom[f][2] = b.x, geom[f][3] = b.y;
  }
}

// one-time: {n_x, n_y, ds, tau/h} of the four faces of every boundary cell (launch_boundary_geometry)
__global__ void boundary_geometry_kernel(DomainView d, double tau, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.n_bcells) return;
  const fe::CellVerts v = load_verts(d, d.bcell[r]);
  rhsmath::boundary_geometry(v, tau, out + 16 * (size_t)r);
}

// face terms of boundary record r added to the rows the cell terms have just been stored to (point-by-point kernel)
template <int KIND>
__device__ __noinline__ void carrier_boundary_terms(const DomainView& d, int other_n_cells, const RhsParams& p, int r,
                                                    const double* u1, const double* u2, const double* o1, const double* o2,
                                                    double* rhs1, double* rhs2) {
  constexpr bool kProduction = KIND == PECS_KIND_PRODUCTION;
  const int c = d.bcell[r];
  const size_t n = (size_t)d.n_cells;
  const fe::CellVerts v = load_verts(d, c);
  double r1[4], r2[4] = {0, 0, 0, 0};
  load4_step(u1 + 8 * n + 4 * (size_t)c, r1);
  if (kProduction) load4_step(u2 + 8 * n + 4 * (size_t)c, r2);
  double jx1[4] = {0, 0, 0, 0}, jy1[4] = {0, 0, 0, 0}, rh1[4] = {0, 0, 0, 0};
  double jx2[4] = {0, 0, 0, 0}, jy2[4] = {0, 0, 0, 0}, rh2[4] = {0, 0, 0, 0};
  const rhsmath::BoundaryRecord rec = load_record(d, r);
  double q1[4] = {0, 0, 0, 0}, q2[4] = {0, 0, 0, 0};
  if (kProduction && rec.nb_cell >= 0) {
    load4_step(o1 + 8 * (size_t)other_n_cells + 4 * (size_t)rec.nb_cell, q1);
    load4_step(o2 + 8 * (size_t)other_n_cells + 4 * (size_t)rec.nb_cell, q2);
  }
  double geom[4][4];
  load_geometry(d, r, geom);
  rhsmath::boundary_terms_accumulate<KIND>(p, rec, geom, v, r1, r2, q1, q2, jx1, jy1, rh1, jx2, jy2, rh2);
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
                                                               const double* X) {
  const bool first = (int)blockIdx.x < blocks_a;
  const CarrierPass& w = pp.pass[first ? 0 : 1]; // stays in the constant bank: no local copy
  const int c = ((int)blockIdx.x - (first ? 0 : blocks_a)) * blockDim.x + threadIdx.x;
  if (c >= w.d.n_cells) return;
  carrier_cell_terms<KIND>(w.d, w.p, c, w.u1, w.u2, X, w.rhs1, w.rhs2);
  const int r = w.d.brecord ? w.d.brecord[c] : -1;
  if (r >= 0) carrier_boundary_terms<KIND>(w.d, w.other_n_cells, w.p, r, w.u1, w.u2, w.o1, w.o2, w.rhs1, w.rhs2);
}

// ------------------------------------------------------------------------------------------ static cell integrals
// Time-independent per-cell tables, evaluated once at context creation (launch_static_cell_integrals):
//   nodal_int[c][a] = sum_q N_a(x_q) JxW_q               (the Poisson charge integral becomes a 4-term dot product)
//   gen_int[c][a]   = sum_q N_a(x_q) G(x_q) JxW_q        (Generation::value, reference Generation.cpp:29-44: the nine
//                                                          exp() per cell and step of SolarCell.cpp:1160-1165 leave the
//                                                          hot loop; same quadrature, summed ahead of time)
__global__ void static_cell_integrals_kernel(DomainView d, RhsParams p, double* __restrict__ nodal_int,
                                             double* __restrict__ gen_int) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.n_cells) return;
  const fe::CellVerts v = load_verts(d, c);
  double m[4], g[4];
  rhsmath::static_cell_integrals(v, gen_int != nullptr, p.gen_scale, p.gen_alpha, p.gen_location, m, g);
  store4(nodal_int + 4 * (size_t)c, m);
  if (gen_int) store4(gen_int + 4 * (size_t)c, g);
}

// ------------------------------------------------------------------------------------------ production carrier kernel
// 256-bit global accesses (sm_100a): one instruction per 4-vector of nodal values
__device__ __forceinline__ void store4_256(double* p, const double v[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
// static tables only (read-only for the lifetime of every grid); states go through ld_step4
__device__ __forceinline__ void load4_256(const double* p, double v[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Streaming production kernel.  One resident wave of blocks; a block owns tiles of kThreads consecutive cells (both
// subdomains form one tile sequence) and walks them with stride gridDim.x.  Every thread keeps the inputs of its NEXT
// cell in flight while it computes the current one: vertices, the two density 4-vectors, the generation integrals and
// the four gathered Poisson fluxes go global -> shared with cp.async (LDGSTS) into the thread's own slot of a two-stage
// ring -- no barrier, no registers.  The flux indices and the boundary record of the tile after next travel the same
// way one group earlier, so the dependent gather never sits on the critical path and nothing loaded is carried in
// registers across the arithmetic.  Cells with boundary faces (1.5 %) are skipped by the cell tiles and done -- cell
// terms + face terms, one store -- by dense BOUNDARY tiles (thread = boundary record) that the first blocks work off
// while their first cell tile is in flight: no warp waits on a divergent lane.  Output: six 256-bit stores per cell.
constexpr int kStageDoubles = 24;                        // per thread and stage: 8 vertices + 8 densities + 4 generation + 4 fluxes
constexpr int kRingDoubles = 2 * kStageDoubles * kThreads;
constexpr int kIndexInts = 5;                            // per thread and stage: 4 flux dofs + boundary record
constexpr int kStreamSmemBytes = kRingDoubles * (int)sizeof(double) + 2 * kIndexInts * kThreads * (int)sizeof(int);

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

struct TileCell {
  int sel; // pass
  int c;   // cell, -1: none
};
__device__ __forceinline__ TileCell locate(const CarrierPassPair& pp, int tile, int tiles_a, int tiles_total) {
  TileCell t;
  t.sel = tile < tiles_a ? 0 : 1;
  const int c = (tile - (t.sel ? tiles_a : 0)) * kThreads + (int)threadIdx.x;
  t.c = (tile < tiles_total && c < pp.pass[t.sel].d.n_cells) ? c : -1;
  return t;
}
// flux dofs and boundary record of a cell -> index stage (ints, [5][T])
__device__ __forceinline__ void issue_index(const CarrierPassPair& pp, TileCell t, int* istage) {
  if (t.c < 0) return;
  const DomainView& d = pp.pass[t.sel].d;
  const int tid = threadIdx.x;
#pragma unroll
  for (int f = 0; f < 4; ++f) cp_async4(istage + f * kThreads + tid, d.rt_dof + (size_t)f * d.n_cells + t.c);
  cp_async4(istage + 4 * kThreads + tid, d.brecord + t.c);
}
// stage layout (doubles, T = kThreads): [0,8T) vertices SoA | [8T,16T) densities as 16-byte pieces [4][T] |
// [16T,20T) generation [2][T] 16-byte pieces | [20T,24T) fluxes SoA; the flux dofs are read from the (landed) index stage
__device__ __forceinline__ void issue_cell(const CarrierPassPair& pp, TileCell t, const int* istage, const double* X,
                                           double* stage) {
  if (t.c < 0) return;
  const CarrierPass& w = pp.pass[t.sel];
  const size_t n = (size_t)w.d.n_cells;
  const int tid = threadIdx.x;
#pragma unroll
  for (int f = 0; f < 4; ++f) cp_async8(stage + (20 + f) * kThreads + tid, X + istage[f * kThreads + tid]);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    cp_async8(stage + a * kThreads + tid, w.d.vx + (size_t)a * n + t.c);
    cp_async8(stage + (4 + a) * kThreads + tid, w.d.vy + (size_t)a * n + t.c);
  }
  const double* u1 = w.u1 + 8 * n + 4 * (size_t)t.c;
  const double* u2 = w.u2 + 8 * n + 4 * (size_t)t.c;
  double* rs = stage + 8 * kThreads;
  cp_async16(rs + 2 * (0 * kThreads + tid), u1);
  cp_async16(rs + 2 * (1 * kThreads + tid), u1 + 2);
  cp_async16(rs + 2 * (2 * kThreads + tid), u2);
  cp_async16(rs + 2 * (3 * kThreads + tid), u2 + 2);
  if (w.d.gen_int) {
    double* gs = stage + 16 * kThreads;
    cp_async16(gs + 2 * (0 * kThreads + tid), w.d.gen_int + 4 * (size_t)t.c);
    cp_async16(gs + 2 * (1 * kThreads + tid), w.d.gen_int + 4 * (size_t)t.c + 2);
  }
}

// WIDE: six 256-bit stores.  The out-of-line boundary routine uses the 128-bit form: inside a non-inlined function
// ptxas 12.9 lowers the inline-PTX st.global.v4.f64 to a 64-bit store of the first element only (seen in the SASS and
// as unwritten rows in tests/test_gpu_extra.py::test_production_rhs_kernels_agree).
template <bool WIDE>
__device__ __forceinline__ void store_cell(const CarrierPass& w, int c, const double jx1[4], const double jy1[4],
                                           const double rh1[4], const double jx2[4], const double jy2[4],
                                           const double rh2[4]) {
  const size_t n = (size_t)w.d.n_cells, o = 4 * (size_t)c;
  if (WIDE) {
    store4_256(w.rhs1 + o, jx1);
    store4_256(w.rhs1 + 4 * n + o, jy1);
    store4_256(w.rhs1 + 8 * n + o, rh1);
    store4_256(w.rhs2 + o, jx2);
    store4_256(w.rhs2 + 4 * n + o, jy2);
    store4_256(w.rhs2 + 8 * n + o, rh2);
  } else {
    store4(w.rhs1 + o, jx1);
    store4(w.rhs1 + 4 * n + o, jy1);
    store4(w.rhs1 + 8 * n + o, rh1);
    store4(w.rhs2 + o, jx2);
    store4(w.rhs2 + 4 * n + o, jy2);
    store4(w.rhs2 + 8 * n + o, rh2);
  }
}

// Shockley-Read-Hall terms of one cell.  The production kernels are compiled TWICE (template parameter SRH): the
// instantiation the reference's configuration runs (SRH_Recombination == 0.0) contains nothing of this, so its registers
// and code are exactly those of the kernel without the switch; the other one calls this out of line
__device__ __noinline__ void srh_terms(const double vx[4], const double vy[4], const double r1[4], const double r2[4],
                                       const RhsParams& p, double rh1[4], double rh2[4]) {
  rhsmath::srh_cell_terms(vx, vy, r1, r2, p.n_i, p.tau_n, p.tau_p, rh1, rh2);
}

// one boundary record: cell terms + face terms of its cell, single writer of the cell's 24 rows.  Three dependent
// round trips to memory at most: {record, cell index} -> {vertices, densities, flux dofs, neighbour densities} -> fluxes
template <bool SRH>
__device__ __noinline__ void boundary_record(const CarrierPass& w, int r, const double* X) {
  const DomainView& d = w.d;
  const size_t n = (size_t)d.n_cells;
  const int c = __ldg(d.bcell + r);
  const rhsmath::BoundaryRecord rec = load_record(d, r);
  double geom[4][4];
  load_geometry(d, r, geom);
  fe::CellVerts v;
  double r1[4], r2[4], Xf[4], gen[4] = {0, 0, 0, 0}, q1[4] = {0, 0, 0, 0}, q2[4] = {0, 0, 0, 0};
  int dof[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    dof[a] = __ldg(d.rt_dof + (size_t)a * n + c);
    v.x[a] = __ldg(d.vx + (size_t)a * n + c);
    v.y[a] = __ldg(d.vy + (size_t)a * n + c);
  }
  ld_vec4(w.u1 + 8 * n + 4 * (size_t)c, r1);
  ld_vec4(w.u2 + 8 * n + 4 * (size_t)c, r2);
  if (d.gen_int) load4_256(d.gen_int + 4 * (size_t)c, gen);
  if (rec.nb_cell >= 0) {
    ld_vec4(w.o1 + 8 * (size_t)w.other_n_cells + 4 * (size_t)rec.nb_cell, q1);
    ld_vec4(w.o2 + 8 * (size_t)w.other_n_cells + 4 * (size_t)rec.nb_cell, q2);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) Xf[a] = ld_vec(X + dof[a]);
  double jx1[4], jy1[4], rh1[4], jx2[4], jy2[4], rh2[4];
  rhsmath::production_cell_terms(v.x, v.y, r1, r2, Xf, gen, w.p.inv_dt, w.p.charge1 * w.p.inv_eps, w.p.charge2 * w.p.inv_eps, jx1,
                        jy1, rh1, jx2, jy2, rh2);
  if (SRH && w.p.srh) srh_terms(v.x, v.y, r1, r2, w.p, rh1, rh2); // w.p.srh: the semiconductor pass only
  double bx1[4] = {0, 0, 0, 0}, by1[4] = {0, 0, 0, 0}, bh1[4] = {0, 0, 0, 0};
  double bx2[4] = {0, 0, 0, 0}, by2[4] = {0, 0, 0, 0}, bh2[4] = {0, 0, 0, 0};
  rhsmath::boundary_terms_accumulate<PECS_KIND_PRODUCTION>(w.p, rec, geom, v, r1, r2, q1, q2, bx1, by1, bh1, bx2, by2, bh2);
#pragma unroll
  for (int a = 0; a < 4; ++a) { // same order of additions as "store the cell terms, then add the face terms"
    jx1[a] += bx1[a];
    jy1[a] += by1[a];
    rh1[a] += bh1[a];
    jx2[a] += bx2[a];
    jy2[a] += by2[a];
    rh2[a] += bh2[a];
  }
  store_cell<false>(w, c, jx1, jy1, rh1, jx2, jy2, rh2);
}

__global__ void __launch_bounds__(kThreads, 4)
    carrier_rhs_stream_kernel(const __grid_constant__ CarrierPassPair pp, int tiles_a, int tiles_total, int btiles_a,
                              int btiles_total, const double* X) {
  extern __shared__ __align__(16) double ring[];
  const int tid = threadIdx.x;
  // the FIRST blocks of the grid do nothing but one boundary tile each: the face terms are a chain of dependent loads
  // and branchy arithmetic several microseconds long, which must start at once and share its SM with streaming blocks
  // instead of being the tail of one
  if ((int)blockIdx.x < btiles_total) {
    const int sel = (int)blockIdx.x < btiles_a ? 0 : 1;
    const int r = ((int)blockIdx.x - (sel ? btiles_a : 0)) * kThreads + tid;
    if (r < pp.pass[sel].d.n_bcells) boundary_record<false>(pp.pass[sel], r, X);
    return;
  }
  int* iring = reinterpret_cast<int*>(ring + kRingDoubles);
  const int grid = (int)gridDim.x - btiles_total, first = (int)blockIdx.x - btiles_total;
  issue_index(pp, locate(pp, first, tiles_a, tiles_total), iring);
  issue_index(pp, locate(pp, first + grid, tiles_a, tiles_total), iring + kIndexInts * kThreads);
  cp_async_commit();
  cp_async_wait<0>();
  issue_cell(pp, locate(pp, first, tiles_a, tiles_total), iring, X, ring);
  cp_async_commit();
  int s = 0;
  for (int tile = first; tile < tiles_total; tile += grid, s ^= 1) {
    cp_async_wait<0>(); // the current cell's values, the next cell's flux dofs and record have landed
    const TileCell cur = locate(pp, tile, tiles_a, tiles_total);
    const int record = cur.c >= 0 ? iring[(s * kIndexInts + 4) * kThreads + tid] : 0;
    issue_cell(pp, locate(pp, tile + grid, tiles_a, tiles_total), iring + (s ^ 1) * kIndexInts * kThreads, X,
               ring + (s ^ 1) * kStageDoubles * kThreads);
    issue_index(pp, locate(pp, tile + 2 * grid, tiles_a, tiles_total), iring + s * kIndexInts * kThreads);
    cp_async_commit();
    if (cur.c >= 0 && record < 0) {
      const double* stage = ring + s * kStageDoubles * kThreads;
      const CarrierPass& w = pp.pass[cur.sel];
      double vx[4], vy[4], r1[4], r2[4], Xf[4], gen[4] = {0, 0, 0, 0};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        vx[a] = stage[a * kThreads + tid];
        vy[a] = stage[(4 + a) * kThreads + tid];
        Xf[a] = stage[(20 + a) * kThreads + tid];
      }
      const double2* rs = reinterpret_cast<const double2*>(stage + 8 * kThreads);
      const double2 p0 = rs[0 * kThreads + tid], p1 = rs[1 * kThreads + tid], p2 = rs[2 * kThreads + tid],
                    p3 = rs[3 * kThreads + tid];
      r1[0] = p0.x, r1[1] = p0.y, r1[2] = p1.x, r1[3] = p1.y;
      r2[0] = p2.x, r2[1] = p2.y, r2[2] = p3.x, r2[3] = p3.y;
      if (w.d.gen_int) {
        const double2* gs = reinterpret_cast<const double2*>(stage + 16 * kThreads);
        const double2 g0 = gs[0 * kThreads + tid], g1 = gs[1 * kThreads + tid];
        gen[0] = g0.x, gen[1] = g0.y, gen[2] = g1.x, gen[3] = g1.y;
      }
      double jx1[4], jy1[4], rh1[4], jx2[4], jy2[4], rh2[4];
      rhsmath::production_cell_terms(vx, vy, r1, r2, Xf, gen, w.p.inv_dt, w.p.charge1 * w.p.inv_eps, w.p.charge2 * w.p.inv_eps,
                            jx1, jy1, rh1, jx2, jy2, rh2);
      store_cell<true>(w, cur.c, jx1, jy1, rh1, jx2, jy2, rh2);
    }
  }
  cp_async_wait<0>();
}

// One-thread-per-cell production kernel on the sum-factorised cell terms (no staging): the variant for meshes too
// small to fill a resident wave, and the A/B partner of the streaming kernel (PECS_B200_RHS_KERNEL=1).
template <int MIN_BLOCKS, int THREADS, bool SRH>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
    carrier_rhs_direct_kernel(const __grid_constant__ CarrierPassPair pp, int blocks_a, int btiles_a, int btiles_total,
                              const double* X) {
  if ((int)blockIdx.x < btiles_total) { // leading blocks: one boundary tile each (see the streaming kernel)
    const int sel = (int)blockIdx.x < btiles_a ? 0 : 1;
    const int r = ((int)blockIdx.x - (sel ? btiles_a : 0)) * THREADS + threadIdx.x;
    if (r < pp.pass[sel].d.n_bcells) boundary_record<SRH>(pp.pass[sel], r, X);
    return;
  }
  const int block = (int)blockIdx.x - btiles_total;
  const bool first = block < blocks_a;
  const CarrierPass& w = pp.pass[first ? 0 : 1];
  const int c = (block - (first ? 0 : blocks_a)) * blockDim.x + threadIdx.x;
  if (c >= w.d.n_cells) return;
  const size_t n = (size_t)w.d.n_cells;
  const int record = __ldg(w.d.brecord + c);
  double vx[4], vy[4], r1[4], r2[4], Xf[4], gen[4] = {0, 0, 0, 0};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    vx[a] = __ldg(w.d.vx + (size_t)a * n + c);
    vy[a] = __ldg(w.d.vy + (size_t)a * n + c);
    Xf[a] = ld_vec(X + __ldg(w.d.rt_dof + (size_t)a * n + c));
  }
  ld_vec4(w.u1 + 8 * n + 4 * (size_t)c, r1);
  ld_vec4(w.u2 + 8 * n + 4 * (size_t)c, r2);
  if (w.d.gen_int) load4_256(w.d.gen_int + 4 * (size_t)c, gen);
  if (record >= 0) return; // done by a boundary tile
  double jx1[4], jy1[4], rh1[4], jx2[4], jy2[4], rh2[4];
  rhsmath::production_cell_terms(vx, vy, r1, r2, Xf, gen, w.p.inv_dt, w.p.charge1 * w.p.inv_eps, w.p.charge2 * w.p.inv_eps, jx1,
                        jy1, rh1, jx2, jy2, rh2);
  if (SRH && w.p.srh) srh_terms(vx, vy, r1, r2, w.p, rh1, rh2); // w.p.srh: the semiconductor pass only
  store_cell<true>(w, c, jx1, jy1, rh1, jx2, jy2, rh2);
}

// ------------------------------------------------------------------------------------------ Poisson cells
template <int KIND>
__global__ void __launch_bounds__(kThreads) poisson_cell_rhs_kernel(const __grid_constant__ CarrierPassPair pp, int blocks_a,
                                                                    const double* __restrict__ static_rows, int n_static,
                                                                    double* __restrict__ poisson_rhs) {
  // the flux rows are time independent (Dirichlet / Schottky face data, evaluated once: poisson_face_rhs_kernel); the
  // reference zeroes and re-assembles the whole vector every step, so they are re-written here, by the same launch
  // (round 1 spent a separate device-to-device copy node on them, on the critical path between the carrier solves and
  // the Poisson solve)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_static; i += gridDim.x * blockDim.x) poisson_rhs[i] = static_rows[i];
  const bool first = (int)blockIdx.x < blocks_a;
  const CarrierPass& w = pp.pass[first ? 0 : 1];
  const DomainView& d = w.d;
  const RhsParams& p = w.p;
  const double* u1 = w.u1;
  const double* u2 = w.u2;
  const int c = ((int)blockIdx.x - (first ? 0 : blocks_a)) * blockDim.x + threadIdx.x;
  if (c >= d.n_cells) return;
  const size_t n = (size_t)d.n_cells;
  if (KIND == PECS_KIND_PRODUCTION && d.nodal_int) {
    // -int (doping + z1 rho1 + z2 rho2) = -sum_a m_a (doping + z1 r1_a + z2 r2_a) with the static m_a = int N_a:
    // the same quadrature sum, reordered; 108 B per cell
    double r1[4], r2[4], m[4];
    ld_vec4(u1 + 8 * n + 4 * (size_t)c, r1);
    ld_vec4(u2 + 8 * n + 4 * (size_t)c, r2);
    load4_256(d.nodal_int + 4 * (size_t)c, m);
    poisson_rhs[d.phi_dof[c]] = rhsmath::poisson_charge_row(p, m, r1, r2);
    return;
  }
  const fe::CellVerts v = load_verts(d, c);
  double r1[4] = {0, 0, 0, 0}, r2[4] = {0, 0, 0, 0};
  if (KIND != PECS_KIND_TEST_STEADY) load4_step(u1 + 8 * n + 4 * (size_t)c, r1);
  if (KIND == PECS_KIND_PRODUCTION) load4_step(u2 + 8 * n + 4 * (size_t)c, r2);
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
        f = p.doping + (p.charge1 * rhsmath::trace(N, r1) + p.charge2 * rhsmath::trace(N, r2));
      } else {
        double x, y;
        fe::map_point(v, xi, eta, x, y);
        f = (KIND == PECS_KIND_TEST_STEADY) ? testfn::poisson_rhs(x, y)
                                            : (testfn::dd_poisson_rhs(x, y, p.time) - rhsmath::trace(N, r1));
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

// I-V post-processing on the device: the two interface integrals of the step's charge-transfer terms
// (reference SolarCell.cpp:1265-1347), one thread per boundary record of the SEMICONDUCTOR, both values per record into
// partial[2 r], partial[2 r + 1] (zero for records without an interface face); the caller adds them up in record order
// on the host (a few hundred values: deterministic, no atomics).
__global__ void interface_current_kernel(CarrierPass w, double* partial) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const DomainView& d = w.d;
  if (r >= d.n_bcells) return;
  double i_et = 0.0, i_ht = 0.0;
  const rhsmath::BoundaryRecord rec = load_record(d, r);
  if (rec.nb_cell >= 0) {
    const size_t n = (size_t)d.n_cells, c = (size_t)d.bcell[r];
    double rn[4], rp[4], rr[4], ro[4], geom[4][4];
    load4_step(w.u1 + 8 * n + 4 * c, rn);
    load4_step(w.u2 + 8 * n + 4 * c, rp);
    load4_step(w.o1 + 8 * (size_t)w.other_n_cells + 4 * (size_t)rec.nb_cell, rr);
    load4_step(w.o2 + 8 * (size_t)w.other_n_cells + 4 * (size_t)rec.nb_cell, ro);
    load_geometry(d, r, geom);
    for (int f = 0; f < 4; ++f) {
      if (rec.id[f] != PECS_INTERFACE) continue;
      for (int q = 0; q < 3; ++q) {
        const double t = fe::gauss_x(q), W = geom[f][2] * fe::gauss_w(q);
        double xi, eta, N[4], Nn[4];
        fe::face_point(f, t, xi, eta);
        fe::shape(xi, eta, N);
        fe::face_point(rec.nb_face, t, xi, eta); // same quadrature index on both sides
        fe::shape(xi, eta, Nn);
        i_et += w.p.k_et * (rhsmath::trace(N, rn) - w.p.rho1_e) * rhsmath::trace(Nn, ro) * W;
        i_ht += w.p.k_ht * (rhsmath::trace(N, rp) - w.p.rho2_e) * rhsmath::trace(Nn, rr) * W;
      }
    }
  }
  partial[2 * r] = i_et;
  partial[2 * r + 1] = i_ht;
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

void launch_static_cell_integrals(const DomainView& d, const RhsParams& p, double* nodal_int, double* gen_int,
                                  cudaStream_t s) {
  if (d.n_cells == 0) return;
  static_cell_integrals_kernel<<<blocks_for(d.n_cells), kThreads, 0, s>>>(d, p, nodal_int, gen_int);
}

void launch_boundary_geometry(const DomainView& d, double tau, double* out, cudaStream_t s) {
  if (d.n_bcells == 0) return;
  boundary_geometry_kernel<<<blocks_for(d.n_bcells), kThreads, 0, s>>>(d, tau, out);
}

int carrier_rhs_variant() {
  // read at every launch (a graph freezes the choice at capture): tests flip it between calls of one process
  const char* e = std::getenv("PECS_B200_RHS_KERNEL");
  return e && *e ? std::atoi(e) : 1;
}

void launch_carrier_rhs(const CarrierPass& a, const CarrierPass& b, int kind, const double* X, cudaStream_t s) {
  const int blocks_a = blocks_for(a.d.n_cells), blocks_b = blocks_for(b.d.n_cells);
  if (blocks_a + blocks_b == 0) return;
  const CarrierPassPair pp{{a, b}};
  // production: the sum-factorised kernels on the static cell tables (variant 0 keeps the point-by-point kernel that
  // also serves the manufactured problems; it is the parity partner of the other two in tests/test_gpu_extra.py)
  int variant = (kind == PECS_KIND_PRODUCTION && a.d.nodal_int) ? carrier_rhs_variant() : 0;
  if (variant >= 2 && variant < 11 && (a.p.srh || b.p.srh)) variant = 1; // the streaming kernel has no SRH instantiation
  const int btiles_a = blocks_for(a.d.n_bcells), btiles_b = blocks_for(b.d.n_bcells), btiles = btiles_a + btiles_b;
  if (variant == 1 || (variant >= 11 && variant <= 15)) {
    // 1: 128 threads, 4 blocks per SM (no spills).  11: compiled for 5 blocks per SM; 14 / 15: 64- / 32-thread blocks
    // (finer tail).  Measured at cfg3 (bench.py variants_ms): 1 is the fastest.
    auto launch = [&](auto kernel, int threads) {
      auto nb = [threads](int n) { return (n + threads - 1) / threads; };
      const int ba = nb(a.d.n_cells), bb = nb(b.d.n_cells), ta = nb(a.d.n_bcells), tb = nb(b.d.n_bcells);
      kernel<<<ta + tb + ba + bb, threads, 0, s>>>(pp, ba, ta, ta + tb, X);
    };
    if (a.p.srh || b.p.srh) { // Shockley-Read-Hall recombination on: the second instantiation, one launch shape
      launch(carrier_rhs_direct_kernel<4, 128, true>, 128);
      return;
    }
    if (variant == 1) launch(carrier_rhs_direct_kernel<4, 128, false>, 128);
    if (variant == 11) launch(carrier_rhs_direct_kernel<5, 128, false>, 128);
    if (variant == 14) launch(carrier_rhs_direct_kernel<8, 64, false>, 64);
    if (variant == 15) launch(carrier_rhs_direct_kernel<16, 32, false>, 32);
    return;
  }
  if (variant >= 2) {
    static int wave_of[64] = {}; // blocks of one resident wave, per device
    int dev = 0;
    PECS_CUDA(cudaGetDevice(&dev));
    int& wave = wave_of[dev & 63];
    if (wave == 0) {
      PECS_CUDA(cudaFuncSetAttribute(carrier_rhs_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemBytes));
      int sms = 0, per_sm = 0;
      PECS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      PECS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, carrier_rhs_stream_kernel, kThreads, kStreamSmemBytes));
      wave = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int tiles = blocks_a + blocks_b;
    int cell_blocks = std::max(1, std::min(tiles, wave - btiles));
    if (const char* e = std::getenv("PECS_B200_RHS_GRID")) // tests: force several tiles per block on small meshes
      if (std::atoi(e) > 0 && std::atoi(e) < cell_blocks) cell_blocks = std::atoi(e);
    carrier_rhs_stream_kernel<<<btiles + cell_blocks, kThreads, kStreamSmemBytes, s>>>(pp, blocks_a, tiles, btiles_a, btiles, X);
    return;
  }
#define CALL(K) carrier_rhs_kernel<K><<<blocks_a + blocks_b, kThreads, 0, s>>>(pp, blocks_a, X)
  PECS_DISPATCH_KIND(kind, CALL)
#undef CALL
}

void launch_poisson_cell_rhs(const CarrierPass& a, const CarrierPass& b, int kind, const double* static_rows, int n_static,
                             double* poisson_rhs, cudaStream_t s) {
  const int blocks_a = blocks_for(a.d.n_cells), blocks_b = blocks_for(b.d.n_cells);
  if (blocks_a + blocks_b == 0) return;
  const CarrierPassPair pp{{a, b}};
#define CALL(K) \
  poisson_cell_rhs_kernel<K><<<blocks_a + blocks_b, kThreads, 0, s>>>(pp, blocks_a, static_rows, n_static, poisson_rhs)
  PECS_DISPATCH_KIND(kind, CALL)
#undef CALL
}

void launch_poisson_face_rhs(const PoissonFaceView& v, const PoissonFaceParams& p, double* static_rhs, cudaStream_t s) {
  if (v.n_faces == 0) return;
  poisson_face_rhs_kernel<<<blocks_for(v.n_faces), kThreads, 0, s>>>(v, p, static_rhs);
}

void launch_interface_currents(const CarrierPass& semiconductor, double* partial, cudaStream_t s) {
  if (semiconductor.d.n_bcells == 0) return;
  interface_current_kernel<<<blocks_for(semiconductor.d.n_bcells), kThreads, 0, s>>>(semiconductor, partial);
}

void launch_distribute(int n_constraints, const int* dof, const int* master, const double* weight, double* x,
                       cudaStream_t s) {
  if (n_constraints == 0) return;
  distribute_kernel<<<blocks_for(n_constraints), kThreads, 0, s>>>(n_constraints, dof, master, weight, x);
}

} // namespace pecs
