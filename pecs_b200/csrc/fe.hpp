// fe.hpp -- host/device finite-element primitives shared by the one-time host assembly and the CUDA kernels.
//
// What deal.II's FEValues / FEFaceValues / FESubfaceValues compute on the fly for the reference
// (update flags: reference source/Assembly.cpp:16-45) is restated here as closed-form inline functions of
// the four cell vertices, so the device never loads per-quadrature-point tables from memory:
//   * bilinear Q1 map x(xi,eta) = sum_a v_a N_a(xi,eta), vertices in deal.II lexicographic order
//   * DGQ1 nodal basis N_a (unmapped values, physical gradients J^-T grad N)
//   * RT0 basis psi_f = J psihat_f / det J with psihat = (1-xi,0),(xi,0),(0,1-eta),(0,eta)  
//   * QGauss(3) on [0,1] and its tensor product (x fastest), face rules along the face's own coordinate
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PECS_HD __host__ __device__ __forceinline__
#else
#define PECS_HD inline
#endif

namespace pecs {
namespace fe {

// QGauss<1>(3) mapped to [0,1]
#define PECS_GAUSS_X0 0.11270166537925831148
#define PECS_GAUSS_X1 0.5
#define PECS_GAUSS_X2 0.88729833462074168852
#define PECS_GAUSS_W0 0.27777777777777777778
#define PECS_GAUSS_W1 0.44444444444444444444
#define PECS_GAUSS_W2 0.27777777777777777778

PECS_HD double gauss_x(int q) { return q == 0 ? PECS_GAUSS_X0 : (q == 1 ? PECS_GAUSS_X1 : PECS_GAUSS_X2); }
PECS_HD double gauss_w(int q) { return q == 1 ? PECS_GAUSS_W1 : PECS_GAUSS_W0; }

// vertices of one cell: x[a], y[a], a = 0..3 lexicographic
struct CellVerts {
  double x[4], y[4];
};

// Jacobian of the bilinear map at (xi, eta):  [xxi xeta; yxi yeta]
struct Jac {
  double xxi, xeta, yxi, yeta, det;
};

PECS_HD Jac jacobian(const CellVerts& v, double xi, double eta) {
  Jac j;
  j.xxi = (v.x[1] - v.x[0]) * (1.0 - eta) + (v.x[3] - v.x[2]) * eta;
  j.yxi = (v.y[1] - v.y[0]) * (1.0 - eta) + (v.y[3] - v.y[2]) * eta;
  j.xeta = (v.x[2] - v.x[0]) * (1.0 - xi) + (v.x[3] - v.x[1]) * xi;
  j.yeta = (v.y[2] - v.y[0]) * (1.0 - xi) + (v.y[3] - v.y[1]) * xi;
  j.det = j.xxi * j.yeta - j.xeta * j.yxi;
  return j;
}

PECS_HD void shape(double xi, double eta, double N[4]) {
  N[0] = (1.0 - xi) * (1.0 - eta);
  N[1] = xi * (1.0 - eta);
  N[2] = (1.0 - xi) * eta;
  N[3] = xi * eta;
}

// reference-cell gradients dN/dxi, dN/deta
PECS_HD void shape_ref_grad(double xi, double eta, double dxi[4], double deta[4]) {
  dxi[0] = -(1.0 - eta);
  dxi[1] = (1.0 - eta);
  dxi[2] = -eta;
  dxi[3] = eta;
  deta[0] = -(1.0 - xi);
  deta[1] = -xi;
  deta[2] = (1.0 - xi);
  deta[3] = xi;
}

// physical gradients: grad N = J^-T (dN/dxi, dN/deta)
PECS_HD void shape_grad(const Jac& j, double xi, double eta, double gx[4], double gy[4]) {
  double dxi[4], deta[4];
  shape_ref_grad(xi, eta, dxi, deta);
  const double inv = 1.0 / j.det;
  for (int a = 0; a < 4; ++a) {
    gx[a] = (j.yeta * dxi[a] - j.yxi * deta[a]) * inv;
    gy[a] = (-j.xeta * dxi[a] + j.xxi * deta[a]) * inv;
  }
}

PECS_HD void map_point(const CellVerts& v, double xi, double eta, double& x, double& y) {
  double N[4];
  shape(xi, eta, N);
  x = v.x[0] * N[0] + v.x[1] * N[1] + v.x[2] * N[2] + v.x[3] * N[3];
  y = v.y[0] * N[0] + v.y[1] * N[1] + v.y[2] * N[2] + v.y[3] * N[3];
}

// RT0: det J * psi_f(x) = J psihat_f(xi,eta); returns the UNSCALED vectors (multiply by 1/det for psi_f)
PECS_HD void rt0_times_det(const Jac& j, double xi, double eta, double px[4], double py[4]) {
  px[0] = j.xxi * (1.0 - xi);
  py[0] = j.yxi * (1.0 - xi);
  px[1] = j.xxi * xi;
  py[1] = j.yxi * xi;
  px[2] = j.xeta * (1.0 - eta);
  py[2] = j.yeta * (1.0 - eta);
  px[3] = j.xeta * eta;
  py[3] = j.yeta * eta;
}
// reference divergence of psihat_f (div psi_f = this / det J)
PECS_HD double rt0_ref_div(int f) { return (f & 1) ? 1.0 : -1.0; }

// Face f of the reference cell (0: xi=0, 1: xi=1, 2: eta=0, 3: eta=1); t in [0,1] runs along the face.
PECS_HD void face_point(int f, double t, double& xi, double& eta) {
  if (f < 2) {
    xi = (double)f;
    eta = t;
  } else {
    xi = t;
    eta = (double)(f - 2);
  }
}

// Outward unit normal and surface element |dx/dt| of face f at reference point (xi, eta).
PECS_HD void face_normal_ds(const Jac& j, int f, double& nx, double& ny, double& ds) {
  // tangent along the face parameter
  const double tx = (f < 2) ? j.xeta : j.xxi;
  const double ty = (f < 2) ? j.yeta : j.yxi;
  ds = sqrt(tx * tx + ty * ty);
  // rotate so that the normal points out of the cell (det J > 0 on all meshes built here)
  // faces 0 (xi=0) and 3 (eta=1): n = (-ty, tx)/ds ; faces 1 (xi=1) and 2 (eta=0): n = (ty, -tx)/ds
  const double s = (f == 1 || f == 2) ? 1.0 : -1.0;
  nx = s * ty / ds;
  ny = -s * tx / ds;
}

PECS_HD double cell_diameter(const CellVerts& v) {
  const double d1 = sqrt((v.x[3] - v.x[0]) * (v.x[3] - v.x[0]) + (v.y[3] - v.y[0]) * (v.y[3] - v.y[0]));
  const double d2 = sqrt((v.x[2] - v.x[1]) * (v.x[2] - v.x[1]) + (v.y[2] - v.y[1]) * (v.y[2] - v.y[1]));
  return d1 > d2 ? d1 : d2;
}

} // namespace fe
} // namespace pecs
