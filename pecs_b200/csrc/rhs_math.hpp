// rhs_math.hpp -- the arithmetic of the production right-hand-side kernels as host/device inline functions.
//
// cuda/rhs_kernels.cu calls these from its kernels; the host build calls the very same code from
// pecs_solarcell_selftest_cell_terms (host/capi_host.cpp), so that the CPU test-suite checks the formulas the GPU runs
// against the oracle (tests/test_host_tables.py::test_sum_factorised_cell_terms_match_oracle) -- the kernels themselves
// (loads, stores, launch shapes, boundary blocks) are covered by the -m gpu parity tests.
#pragma once
#include "../../include/pecs_b200.h"
#include "fe.hpp"
#include "test_functions.hpp"

namespace pecs {

// scalars of one subdomain pass (see include/pecs_b200.h PECS_P_*)
struct RhsParams {
  int kind;            // PECS_KIND_*
  int is_semiconductor;
  double inv_dt;       // 1 / delta_t (carried by the mass matrix in the reference, LDG.cpp:77-79)
  double tau;          // penalty
  double charge1, charge2;
  double inv_eps;
  double gen_scale, gen_alpha, gen_location; // alpha*G0, alpha, H (0 scale = dark)
  double rho1_e, rho2_e;                     // equilibrium / Dirichlet densities of this subdomain's carriers
  double other1_e, other2_e;                 // electrons_e, holes_e as seen from the electrolyte side
  double k_et, k_ht, v_n, v_p;
  double doping;       // N_D - N_A (semiconductor) or 0 (electrolyte)
  double time;         // manufactured right-hand sides
  int srh;             // Shockley-Read-Hall recombination on (semiconductor only; 0 = the reference's function body)
  double n_i, tau_n, tau_p;
};

// the scalars of subdomain w (0 semiconductor, 1 electrolyte) from the PECS_P_* parameter block of the problem
inline RhsParams make_rhs_params(const double* p, int kind, int w) {
  RhsParams r{};
  r.kind = kind;
  r.is_semiconductor = w == 0;
  r.inv_dt = 1.0 / p[PECS_P_DELTA_T];
  r.tau = p[PECS_P_PENALTY];
  r.charge1 = -1.0; // reference SolarCell.cpp:54,59,71,76
  r.charge2 = 1.0;
  r.inv_eps = 1.0 / (w == 0 ? p[PECS_P_EPS_S] : p[PECS_P_EPS_E]);
  r.gen_scale = w == 0 ? p[PECS_P_GEN_ALPHA] * p[PECS_P_GEN_FLUX] : 0.0;
  r.gen_alpha = p[PECS_P_GEN_ALPHA];
  r.gen_location = p[PECS_P_GEN_LOCATION];
  r.rho1_e = w == 0 ? p[PECS_P_RHO_N_E] : p[PECS_P_RHO_R_E];
  r.rho2_e = w == 0 ? p[PECS_P_RHO_P_E] : p[PECS_P_RHO_O_E];
  r.other1_e = p[PECS_P_RHO_N_E];
  r.other2_e = p[PECS_P_RHO_P_E];
  r.k_et = p[PECS_P_K_ET];
  r.k_ht = p[PECS_P_K_HT];
  r.v_n = p[PECS_P_V_N];
  r.v_p = p[PECS_P_V_P];
  r.doping = w == 0 ? p[PECS_P_RHO_N_E] - p[PECS_P_RHO_P_E] : 0.0; // N_D = electrons_e, N_A = holes_e (SolarCell.cpp:542-548)
  r.time = 0.0;
  r.srh = (w == 0 && kind == PECS_KIND_PRODUCTION && p[PECS_P_SRH] != 0.0) ? 1 : 0;
  r.n_i = p[PECS_P_N_INTRINSIC];
  r.tau_n = p[PECS_P_TAU_N];
  r.tau_p = p[PECS_P_TAU_P];
  return r;
}

namespace rhsmath {


// Time-independent per-cell integrals (evaluated once at context creation):
//   m[a] = sum_q N_a(x_q) JxW_q,   g[a] = sum_q N_a(x_q) G(x_q) JxW_q,   G(x) = gen_scale exp(gen_alpha (y - gen_location))
// (Generation::value, reference source/Generation.cpp:29-44; quadrature of source/SolarCell.cpp:1160-1165)
PECS_HD void static_cell_integrals(const fe::CellVerts& v, bool with_generation, double gen_scale, double gen_alpha,
                                   double gen_location, double m[4], double g[4]) {
  for (int a = 0; a < 4; ++a) m[a] = g[a] = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int qy = 0; qy < 3; ++qy)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int qx = 0; qx < 3; ++qx) {
      const double xi = fe::gauss_x(qx), eta = fe::gauss_x(qy), w = fe::gauss_w(qx) * fe::gauss_w(qy);
      const fe::Jac j = fe::jacobian(v, xi, eta);
      double N[4];
      fe::shape(xi, eta, N);
      const double JxW = j.det * w;
      double gen = 0.0;
      if (with_generation) {
        const double y = v.y[0] * N[0] + v.y[1] * N[1] + v.y[2] * N[2] + v.y[3] * N[3];
        gen = gen_scale * exp(gen_alpha * (y - gen_location));
      }
      for (int a = 0; a < 4; ++a) {
        m[a] += N[a] * JxW;
        g[a] += N[a] * (gen * JxW);
      }
    }
}

// reference include/SolarCell.hpp:86-98, the formula the reference carries as a comment (its function returns 0.0)
PECS_HD double srh_recombination(double electron_density, double hole_density, double n_i, double tau_n, double tau_p) {
  return (n_i * n_i - electron_density * hole_density) / (tau_n * (electron_density - n_i) + tau_p * (hole_density - n_i));
}
// int N_a R(rho_n, rho_p) over one cell, added to the density rows of BOTH carriers (reference SolarCell.cpp:1160-1187).
// R is not polynomial in the densities: it is evaluated point by point on the 3 x 3 Gauss rule (one division per
// point); the production kernels call this only when the switch is on, out of line of the sum-factorised terms.
PECS_HD void srh_cell_terms(const double vx[4], const double vy[4], const double r1[4], const double r2[4], double n_i,
                            double tau_n, double tau_p, double rh1[4], double rh2[4]) {
  fe::CellVerts v;
  for (int a = 0; a < 4; ++a) {
    v.x[a] = vx[a];
    v.y[a] = vy[a];
  }
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int qy = 0; qy < 3; ++qy)
    for (int qx = 0; qx < 3; ++qx) {
      const double xi = fe::gauss_x(qx), eta = fe::gauss_x(qy), w = fe::gauss_w(qx) * fe::gauss_w(qy);
      const fe::Jac j = fe::jacobian(v, xi, eta);
      double N[4];
      fe::shape(xi, eta, N);
      const double rho_n = N[0] * r1[0] + N[1] * r1[1] + N[2] * r1[2] + N[3] * r1[3];
      const double rho_p = N[0] * r2[0] + N[1] * r2[1] + N[2] * r2[2] + N[3] * r2[3];
      const double R = srh_recombination(rho_n, rho_p, n_i, tau_n, tau_p) * (j.det * w);
      for (int a = 0; a < 4; ++a) acc[a] += N[a] * R;
    }
  for (int a = 0; a < 4; ++a) {
    rh1[a] += acc[a];
    rh2[a] += acc[a];
  }
}

// Cell terms of both carriers of one production cell from values in registers, sum-factorised over the 3 x 3 tensor
// Gauss rule: with N_a = L_ax(xi) L_ay(eta), XA_j = w_j x_xi(eta_j), XB_i = w_i x_eta(xi_i), HX_i = w_i Dhat_x(xi_i),
// HY_j = w_j Dhat_y(eta_j) (and Y likewise)
//     JxW_ij = XA_j YB_i - XB_i YA_j,   JxW_ij eps E_x = XA_j HX_i + XB_i HY_j,   JxW_ij eps E_y = YA_j HX_i + YB_i HY_j
// are shared by both carriers; per carrier the three integrands rho {JxW, Ex, Ey} are contracted first along xi, then
// along eta: ~420 fp64 operations per cell instead of ~800 for the point-by-point form, no division, no exp.
PECS_HD void production_cell_terms(const double vx[4], const double vy[4], const double r1[4],
                                                      const double r2[4], const double Xf[4], const double gen[4],
                                                      double inv_dt, double s1, double s2, double jx1[4], double jy1[4],
                                                      double rh1[4], double jx2[4], double jy2[4], double rh2[4]) {
  const double ax = vx[1] - vx[0], bx = vx[3] - vx[2], cx = vx[2] - vx[0], dx = vx[3] - vx[1];
  const double ay = vy[1] - vy[0], by = vy[3] - vy[2], cy = vy[2] - vy[0], dy = vy[3] - vy[1];
  double XA[3], YA[3], XB[3], YB[3], HX[3], HY[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double l0w = (1.0 - fe::gauss_x(k)) * fe::gauss_w(k), l1w = fe::gauss_x(k) * fe::gauss_w(k);
    XA[k] = ax * l0w + bx * l1w;
    YA[k] = ay * l0w + by * l1w;
    XB[k] = cx * l0w + dx * l1w;
    YB[k] = cy * l0w + dy * l1w;
    HX[k] = Xf[0] * l0w + Xf[1] * l1w;
    HY[k] = Xf[2] * l0w + Xf[3] * l1w;
  }
  double JxW[3][3], Ex[3][3], Ey[3][3]; // [j][i]
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      JxW[j][i] = XA[j] * YB[i] - XB[i] * YA[j];
      Ex[j][i] = XA[j] * HX[i] + XB[i] * HY[j];
      Ey[j][i] = YA[j] * HX[i] + YB[i] * HY[j];
    }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double* r = k == 0 ? r1 : r2;
    double* jx = k == 0 ? jx1 : jx2;
    double* jy = k == 0 ? jy1 : jy2;
    double* rh = k == 0 ? rh1 : rh2;
    const double s = k == 0 ? s1 : s2;
    double bot[3], top[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double l0 = 1.0 - fe::gauss_x(i), l1 = fe::gauss_x(i);
      bot[i] = r[0] * l0 + r[1] * l1;
      top[i] = r[2] * l0 + r[3] * l1;
    }
    double ax_[4] = {0, 0, 0, 0}, ay_[4] = {0, 0, 0, 0}, am[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double m0 = 1.0 - fe::gauss_x(j), m1 = fe::gauss_x(j);
      double sm0 = 0, sm1 = 0, sx0 = 0, sx1 = 0, sy0 = 0, sy1 = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double rho = bot[i] * m0 + top[i] * m1;
        const double rho0 = rho * (1.0 - fe::gauss_x(i)), rho1 = rho * fe::gauss_x(i);
        sm0 += rho0 * JxW[j][i];
        sm1 += rho1 * JxW[j][i];
        sx0 += rho0 * Ex[j][i];
        sx1 += rho1 * Ex[j][i];
        sy0 += rho0 * Ey[j][i];
        sy1 += rho1 * Ey[j][i];
      }
      am[0] += m0 * sm0;
      am[1] += m0 * sm1;
      am[2] += m1 * sm0;
      am[3] += m1 * sm1;
      ax_[0] += m0 * sx0;
      ax_[1] += m0 * sx1;
      ax_[2] += m1 * sx0;
      ax_[3] += m1 * sx1;
      ay_[0] += m0 * sy0;
      ay_[1] += m0 * sy1;
      ay_[2] += m1 * sy0;
      ay_[3] += m1 * sy1;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      rh[a] = am[a] * inv_dt + gen[a];
      jx[a] = s * ax_[a];
      jy[a] = s * ay_[a];
    }
  }
}

// trace of the density of a cell at a face point
PECS_HD double trace(const double N[4], const double r[4]) {
  return N[0] * r[0] + N[1] * r[1] + N[2] * r[2] + N[3] * r[3];
}


// one boundary record as the face routine wants it: all loads are independent of each other
struct BoundaryRecord {
  int id[4];   // boundary id of face f, -1: interior face
  int nb_cell; // matched cell of the other subdomain across the interface face, -1: none
  int nb_face;
};

// {n_x, n_y, ds, tau/h} of the four faces of a cell into out[4][4] (static: evaluated once per boundary cell)
PECS_HD void boundary_geometry(const fe::CellVerts& v, double tau, double* out) {
  const double pen = tau / fe::cell_diameter(v);
  for (int f = 0; f < 4; ++f) {
    double xi, eta, nx, ny, ds;
    fe::face_point(f, 0.5, xi, eta);
    fe::face_normal_ds(fe::jacobian(v, xi, eta), f, nx, ny, ds);
    double* o = out + 4 * f;
    o[0] = nx, o[1] = ny, o[2] = ds, o[3] = pen;
  }
}

// geom[f] = {n_x, n_y, |dx/dt|, tau/h} of face f: the edges of a bilinear cell are straight, so normal and surface
// element are constant along a face and time independent -- evaluated once (boundary_geometry_kernel) instead of a
// square root and two divisions per quadrature point and step.  The face and point loops are unrolled: the traces
// N_a(x_q) become immediates (two of the four vanish on a face).
template <int KIND>
PECS_HD void boundary_terms_accumulate(const RhsParams& p, const BoundaryRecord& rec,
                                                          const double geom[4][4], const fe::CellVerts& v,
                                                          const double r1[4], const double r2[4], const double q1[4],
                                                          const double q2[4], double jx1[4], double jy1[4], double rh1[4],
                                                          double jx2[4], double jy2[4], double rh2[4]) {
  constexpr bool kProduction = KIND == PECS_KIND_PRODUCTION;
  const int nb_face = rec.nb_face;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int id = rec.id[f];
    if (id < 0 || id == PECS_NEUMANN) continue; // interior face, or insulating: nothing to do
    const double nx = geom[f][0], ny = geom[f][1], ds = geom[f][2], pen = geom[f][3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double t = fe::gauss_x(q);
      double xi, eta, N[4];
      fe::face_point(f, t, xi, eta);
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
}

// Potential row of the Poisson right-hand side of one carrier cell on the static table m_a = int N_a:
//   -int (doping + z1 rho1 + z2 rho2) = -sum_a m_a (doping + z1 r1_a + z2 r2_a)
// (reference source/SolarCell.cpp:551-578, 741-762: only the DG0 potential test function is non-zero)
PECS_HD double poisson_charge_row(const RhsParams& p, const double m[4], const double r1[4], const double r2[4]) {
  double acc = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int a = 0; a < 4; ++a) acc -= m[a] * (p.doping + (p.charge1 * r1[a] + p.charge2 * r2[a]));
  return acc;
}

// Output path: RT0 field J psihat / det J at patch vertex a (deal.II lexicographic) of a Poisson cell with face fluxes
// Xf, times scale (reference DataOut::build_patches + PostProcessor.cpp:105-118)
PECS_HD void rt0_field_at_vertex(const fe::CellVerts& v, const double Xf[4], int a, double scale, double& fx, double& fy) {
  const double xi = (double)(a & 1), eta = (double)(a >> 1);
  const fe::Jac j = fe::jacobian(v, xi, eta);
  const double dhx = Xf[0] * (1.0 - xi) + Xf[1] * xi, dhy = Xf[2] * (1.0 - eta) + Xf[3] * eta;
  const double s = scale / j.det;
  fx = s * (j.xxi * dhx + j.xeta * dhy);
  fy = s * (j.yxi * dhx + j.yeta * dhy);
}

} // namespace rhsmath
} // namespace pecs
