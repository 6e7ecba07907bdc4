// oracle/fe_values.hpp -- TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// Restatement of what deal.II's FEValues / FEFaceValues / FESubfaceValues deliver to the reference's local
// assemblers (update flags reference source/Assembly.cpp:16-45) for the two element pairs of degree 1:
//   carrier FE  = FESystem(FESystem(FE_DGQ(1),2),1, FE_DGQ(1),1)   (reference source/SolarCell.cpp:30-31)
//   Poisson FE  = FESystem(FE_RaviartThomas(0),1, FE_DGQ(0),1)     (reference source/SolarCell.cpp:23-24)
// Conventions: Q1 mapping, QGauss(3) tensor rule with x fastest, DGQ1 nodal at the vertices in
// lexicographic order, RT0 = contravariant Piola image of (1-xi,0),(xi,0),(0,1-eta),(0,eta), local dof order
// [Jx0-3, Jy0-3, rho0-3] and [face0-3, Phi].  PARITY UNPINNED against deal.II itself (it cannot be built here);
// pinned only through the manufactured-solution convergence orders (see oracle/README.md).
#pragma once
#include <cmath>

namespace oracle {

static const double GAUSS_X[3] = {0.5 - 0.5 * 0.77459666924148337704, 0.5, 0.5 + 0.5 * 0.77459666924148337704};
static const double GAUSS_W[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};

struct Tensor1 {
  double c[2];
  double operator*(const Tensor1& o) const { return c[0] * o.c[0] + c[1] * o.c[1]; }
};

// Everything the local assemblers ask for at n_q points of one cell (n_q = 9 in the cell, 3 on a face).
struct FEValuesBase {
  int n_q = 0;
  double JxW_[9];
  Tensor1 point_[9];
  Tensor1 normal_[9];
  // DGQ1 nodal basis
  double N_[4][9];
  Tensor1 gradN_[4][9];
  // RT0 basis
  Tensor1 rt_[4][9];
  double rt_div_[4][9];

  double JxW(int q) const { return JxW_[q]; }
  const Tensor1& quadrature_point(int q) const { return point_[q]; }
  const Tensor1& normal_vector(int q) const { return normal_[q]; }

  // --- carrier FE, local dof i in 0..11: component i/4 (0 Jx, 1 Jy, 2 rho), node i%4 ---
  Tensor1 current_value(int i, int q) const {
    Tensor1 t{{0, 0}};
    if (i < 8) t.c[i / 4] = N_[i % 4][q];
    return t;
  }
  double current_divergence(int i, int q) const { return i < 8 ? gradN_[i % 4][q].c[i / 4] : 0.0; }
  double density_value(int i, int q) const { return i >= 8 ? N_[i - 8][q] : 0.0; }
  Tensor1 density_gradient(int i, int q) const { return i >= 8 ? gradN_[i - 8][q] : Tensor1{{0, 0}}; }
  // --- Poisson FE, local dof i in 0..4: 0..3 RT0 faces, 4 potential ---
  Tensor1 field_value(int i, int q) const { return i < 4 ? rt_[i][q] : Tensor1{{0, 0}}; }
  double field_divergence(int i, int q) const { return i < 4 ? rt_div_[i][q] : 0.0; }
  double potential_value(int i, int) const { return i == 4 ? 1.0 : 0.0; }

protected:
  // fill all tables at reference point (xi, eta) with weight w; face >= 0 also fills normal and surface JxW
  void fill(const double* vtx, int q, double xi, double eta, double w, int face) {
    const double x0 = vtx[0], y0 = vtx[1], x1 = vtx[2], y1 = vtx[3], x2 = vtx[4], y2 = vtx[5], x3 = vtx[6], y3 = vtx[7];
    const double Nv[4] = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
    const double dxi[4] = {-(1 - eta), (1 - eta), -eta, eta};
    const double deta[4] = {-(1 - xi), -xi, (1 - xi), xi};
    // J = d(x,y)/d(xi,eta)
    const double J00 = x0 * dxi[0] + x1 * dxi[1] + x2 * dxi[2] + x3 * dxi[3];
    const double J01 = x0 * deta[0] + x1 * deta[1] + x2 * deta[2] + x3 * deta[3];
    const double J10 = y0 * dxi[0] + y1 * dxi[1] + y2 * dxi[2] + y3 * dxi[3];
    const double J11 = y0 * deta[0] + y1 * deta[1] + y2 * deta[2] + y3 * deta[3];
    const double det = J00 * J11 - J01 * J10;
    // inverse Jacobian (covariant transformation of gradients: grad = J^-T grad_ref)
    const double I00 = J11 / det, I01 = -J01 / det, I10 = -J10 / det, I11 = J00 / det;
    point_[q].c[0] = x0 * Nv[0] + x1 * Nv[1] + x2 * Nv[2] + x3 * Nv[3];
    point_[q].c[1] = y0 * Nv[0] + y1 * Nv[1] + y2 * Nv[2] + y3 * Nv[3];
    for (int a = 0; a < 4; ++a) {
      N_[a][q] = Nv[a];
      gradN_[a][q].c[0] = I00 * dxi[a] + I10 * deta[a];
      gradN_[a][q].c[1] = I01 * dxi[a] + I11 * deta[a];
    }
    // RT0 reference functions and their contravariant Piola image  psi = J psihat / det
    const double ph[4][2] = {{1 - xi, 0}, {xi, 0}, {0, 1 - eta}, {0, eta}};
    const double dv[4] = {-1, 1, -1, 1};
    for (int f = 0; f < 4; ++f) {
      rt_[f][q].c[0] = (J00 * ph[f][0] + J01 * ph[f][1]) / det;
      rt_[f][q].c[1] = (J10 * ph[f][0] + J11 * ph[f][1]) / det;
      rt_div_[f][q] = dv[f] / det;
    }
    if (face < 0) {
      JxW_[q] = det * w;
      normal_[q] = Tensor1{{0, 0}};
    } else {
      // boundary form: J^-T nhat * det; its norm is the surface element, its direction the outward normal
      const double nh[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
      const double bx = (I00 * nh[face][0] + I10 * nh[face][1]) * det;
      const double by = (I01 * nh[face][0] + I11 * nh[face][1]) * det;
      const double len = std::sqrt(bx * bx + by * by);
      normal_[q].c[0] = bx / len;
      normal_[q].c[1] = by / len;
      JxW_[q] = len * w;
    }
  }
};

struct FEValues : FEValuesBase {
  void reinit(const double* vtx) {
    n_q = 9;
    for (int qy = 0; qy < 3; ++qy)
      for (int qx = 0; qx < 3; ++qx) fill(vtx, qx + 3 * qy, GAUSS_X[qx], GAUSS_X[qy], GAUSS_W[qx] * GAUSS_W[qy], -1);
  }
};

struct FEFaceValues : FEValuesBase {
  // whole face
  void reinit(const double* vtx, int face) { reinit_part(vtx, face, 0.0, 1.0); }

protected:
  void reinit_part(const double* vtx, int face, double t0, double len) {
    n_q = 3;
    for (int q = 0; q < 3; ++q) {
      const double t = t0 + len * GAUSS_X[q];
      const double xi = face < 2 ? (double)face : t;
      const double eta = face < 2 ? t : (double)(face - 2);
      fill(vtx, q, xi, eta, len * GAUSS_W[q], face);
    }
  }
};

struct FESubfaceValues : FEFaceValues {
  // sub-face s (0: first half, 1: second half of the face's own coordinate)
  void reinit(const double* vtx, int face, int subface) { reinit_part(vtx, face, 0.5 * subface, 0.5); }
};

// QIterated(QTrapez, 3) x QIterated(QTrapez, 3): the rule the reference integrates errors with
// (reference source/LDG.cpp:1000-1002, source/MixedFEM.cpp:271-273)
struct FEErrorValues : FEValuesBase {
  int n_points() const { return 16; }
  double JxW16[16];
  Tensor1 point16[16];
  double N16[4][16];
  Tensor1 rt16[4][16];
  void reinit(const double* vtx) {
    static const double T[4] = {0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0};
    static const double W[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
    for (int qy = 0; qy < 4; ++qy)
      for (int qx = 0; qx < 4; ++qx) {
        const int q = qx + 4 * qy;
        fill(vtx, 0, T[qx], T[qy], W[qx] * W[qy], -1);
        JxW16[q] = JxW_[0];
        point16[q] = point_[0];
        for (int a = 0; a < 4; ++a) {
          N16[a][q] = N_[a][0];
          rt16[a][q] = rt_[a][0];
        }
      }
  }
};

} // namespace oracle
