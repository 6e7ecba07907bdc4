// LDG.cpp -- see LDG.hpp.
#include "LDG.hpp"

#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../fe.hpp"

namespace LDG_System {

using pecs::CarrierDofs;
using pecs::CsrMatrix;
using pecs::MeshTables;
using pecs::TripletList;
using pecs::fe::CellVerts;
using pecs::fe::Jac;

namespace {

CellVerts load_verts(const MeshTables& mesh, int c) {
  CellVerts v;
  const double* p = mesh.vtx(c);
  for (int a = 0; a < 4; ++a) {
    v.x[a] = p[2 * a];
    v.y[a] = p[2 * a + 1];
  }
  return v;
}

// M_ab = int N_a N_b, Dx_ab = int d_x N_a N_b, Dy_ab = int d_y N_a N_b  (QGauss<2>(3); det J cancels in Dx, Dy)
void cell_tables(const CellVerts& v, double M[4][4], double Dx[4][4], double Dy[4][4]) {
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) M[a][b] = Dx[a][b] = Dy[a][b] = 0.0;
  for (int qy = 0; qy < 3; ++qy)
    for (int qx = 0; qx < 3; ++qx) {
      const double xi = pecs::fe::gauss_x(qx), eta = pecs::fe::gauss_x(qy);
      const double w = pecs::fe::gauss_w(qx) * pecs::fe::gauss_w(qy);
      const Jac j = pecs::fe::jacobian(v, xi, eta);
      double N[4], dxi[4], deta[4];
      pecs::fe::shape(xi, eta, N);
      pecs::fe::shape_ref_grad(xi, eta, dxi, deta);
      for (int a = 0; a < 4; ++a) {
        const double gxw = (j.yeta * dxi[a] - j.yxi * deta[a]) * w;
        const double gyw = (-j.xeta * dxi[a] + j.xxi * deta[a]) * w;
        const double mw = N[a] * j.det * w;
        for (int b = 0; b < 4; ++b) {
          M[a][b] += mw * N[b];
          Dx[a][b] += gxw * N[b];
          Dy[a][b] += gyw * N[b];
        }
      }
    }
}

// Trace tables on (part of) a face.  The minus cell sees the face as its face fm over the parameter range
// [t0, t0+len]; the plus cell (if any) sees the same points as its face fp over [0,1].
struct FaceTables {
  double Tmm[4][4], Tmp[4][4], Tpm[4][4], Tpp[4][4];
  double nx, ny;
};

FaceTables face_tables(const CellVerts& vm, int fm, double t0, double len, bool has_plus, int fp) {
  FaceTables F;
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) F.Tmm[a][b] = F.Tmp[a][b] = F.Tpm[a][b] = F.Tpp[a][b] = 0.0;
  F.nx = F.ny = 0.0;
  for (int q = 0; q < 3; ++q) {
    const double tq = pecs::fe::gauss_x(q);
    double xi, eta, ds;
    pecs::fe::face_point(fm, t0 + len * tq, xi, eta);
    const Jac j = pecs::fe::jacobian(vm, xi, eta);
    pecs::fe::face_normal_ds(j, fm, F.nx, F.ny, ds);
    const double W = ds * len * pecs::fe::gauss_w(q);
    double Nm[4], Np[4] = {0, 0, 0, 0};
    pecs::fe::shape(xi, eta, Nm);
    if (has_plus) {
      double xip, etap;
      pecs::fe::face_point(fp, tq, xip, etap);
      pecs::fe::shape(xip, etap, Np);
    }
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        F.Tmm[a][b] += Nm[a] * Nm[b] * W;
        if (has_plus) {
          F.Tmp[a][b] += Nm[a] * Np[b] * W;
          F.Tpm[a][b] += Np[a] * Nm[b] * W;
          F.Tpp[a][b] += Np[a] * Np[b] * W;
        }
      }
  }
  return F;
}

} // namespace

CsrMatrix LDG::assemble_mass_matrix(const MeshTables& mesh, double delta_t) const {
  CarrierDofs dofs{mesh.n_cells};
  TripletList tl(dofs.n_dofs());
  tl.reserve(16 * (size_t)mesh.n_cells);
  double M[4][4], Dx[4][4], Dy[4][4];
  for (int c = 0; c < mesh.n_cells; ++c) {
    cell_tables(load_verts(mesh, c), M, Dx, Dy);
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) tl.add(dofs.global(c, 8 + a), dofs.global(c, 8 + b), M[a][b] / delta_t);
  }
  return tl.compress();
}

void LDG::assemble_system_matrices(const MeshTables& mesh, int dirichlet_id, double mu1, double mu2, double delta_t,
                                   double transient_or_steady, double penalty, CsrMatrix& matrix_1,
                                   CsrMatrix& matrix_2) const {
  CarrierDofs dofs{mesh.n_cells};
  // one chunk of triplets per thread over a contiguous range of cells; compressed in chunk order, so the result does
  // not depend on the number of threads (reference LDG.cpp:283-426 assembles the flux terms sequentially)
  const double t_begin = omp_get_wtime();
  const int n_chunks = std::max(1, std::min(omp_get_max_threads(), mesh.n_cells / 64));
  pecs::PairedTripletChunks chunks(dofs.n_dofs(), n_chunks);
  const double beta[2] = {1.0 / std::sqrt(2.0), 1.0 / std::sqrt(2.0)};
  const double mass_scale = transient_or_steady / delta_t;

#pragma omp parallel for schedule(static, 1) num_threads(n_chunks)
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
  pecs::PairedTripletChunks::Chunk& out = chunks.chunk(chunk);
  const int c_begin = (int)((long long)mesh.n_cells * chunk / n_chunks);
  const int c_end = (int)((long long)mesh.n_cells * (chunk + 1) / n_chunks);
  out.reserve(250 * (size_t)(c_end - c_begin));
  auto both = [&](int i, int j, double v) { out.add(i, j, v, v); };
  double M[4][4], Dx[4][4], Dy[4][4];
  for (int c = c_begin; c < c_end; ++c) {
    const CellVerts v = load_verts(mesh, c);
    const double h = pecs::fe::cell_diameter(v);
    cell_tables(v, M, Dx, Dy);
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        const int ja = dofs.global(c, a), ka = dofs.global(c, 4 + a), ua = dofs.global(c, 8 + a);
        const int jb = dofs.global(c, b), kb = dofs.global(c, 4 + b), ub = dofs.global(c, 8 + b);
        out.add(ja, jb, M[a][b] / mu1, M[a][b] / mu2);
        out.add(ka, kb, M[a][b] / mu1, M[a][b] / mu2);
        both(ja, ub, -Dx[a][b]); // -(div p) u
        both(ka, ub, -Dy[a][b]);
        both(ua, jb, -Dx[a][b]); // -grad v . q
        both(ua, kb, -Dy[a][b]);
        both(ua, ub, mass_scale * M[a][b]);
      }

    for (int f = 0; f < 4; ++f) {
      const int kind = mesh.face_kind[4 * c + f];
      if (kind == pecs::FACE_BOUNDARY) {
        const FaceTables F = face_tables(v, f, 0.0, 1.0, false, 0);
        const double n[2] = {F.nx, F.ny};
        const bool dirichlet = mesh.boundary_id[4 * c + f] == dirichlet_id;
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) {
            const double T = F.Tmm[a][b];
            if (T == 0.0) continue;
            for (int d = 0; d < 2; ++d) {
              if (dirichlet)
                both(dofs.global(c, 8 + a), dofs.global(c, 4 * d + b), n[d] * T); // v n.q
              else
                both(dofs.global(c, 4 * d + a), dofs.global(c, 8 + b), n[d] * T); // (p.n) u
            }
            if (dirichlet) both(dofs.global(c, 8 + a), dofs.global(c, 8 + b), (penalty / h) * T);
          }
        continue;
      }
      // interior faces: same level -> lower index does the work; hanging -> coarse side, per sub-face
      int n_parts = 0, plus_cell[2] = {-1, -1};
      double sigma = 0.0;
      if (kind == pecs::FACE_SAME_LEVEL) {
        const int nb = mesh.neighbor[4 * c + f];
        if (nb < c) continue;
        n_parts = 1;
        plus_cell[0] = nb;
        sigma = penalty / std::min(h, mesh.diameter(nb));
      } else if (kind == pecs::FACE_HAS_CHILDREN) {
        n_parts = 2;
        plus_cell[0] = mesh.neighbor[4 * c + f];
        plus_cell[1] = mesh.neighbor2[4 * c + f];
        sigma = penalty / std::min(h, mesh.nb_parent_diameter[4 * c + f]);
      } else {
        continue; // FACE_COARSER: assembled from the coarse side
      }
      for (int s = 0; s < n_parts; ++s) {
        const int e = plus_cell[s];
        const double len = 1.0 / n_parts;
        const FaceTables F = face_tables(v, f, s * len, len, true, f ^ 1);
        const double n[2] = {F.nx, F.ny};
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) {
            for (int d = 0; d < 2; ++d) {
              const double hp = 0.5 * n[d] + beta[d], hm = 0.5 * n[d] - beta[d];
              // rows: current test functions, cols: densities
              both(dofs.global(c, 4 * d + a), dofs.global(c, 8 + b), hp * F.Tmm[a][b]);
              both(dofs.global(c, 4 * d + a), dofs.global(e, 8 + b), hm * F.Tmp[a][b]);
              both(dofs.global(e, 4 * d + a), dofs.global(c, 8 + b), -hp * F.Tpm[a][b]);
              both(dofs.global(e, 4 * d + a), dofs.global(e, 8 + b), -hm * F.Tpp[a][b]);
              // rows: density test functions, cols: currents
              both(dofs.global(c, 8 + a), dofs.global(c, 4 * d + b), hm * F.Tmm[a][b]);
              both(dofs.global(c, 8 + a), dofs.global(e, 4 * d + b), hp * F.Tmp[a][b]);
              both(dofs.global(e, 8 + a), dofs.global(c, 4 * d + b), -hm * F.Tpm[a][b]);
              both(dofs.global(e, 8 + a), dofs.global(e, 4 * d + b), -hp * F.Tpp[a][b]);
            }
            // penalty on the density jump
            both(dofs.global(c, 8 + a), dofs.global(c, 8 + b), sigma * F.Tmm[a][b]);
            both(dofs.global(c, 8 + a), dofs.global(e, 8 + b), -sigma * F.Tmp[a][b]);
            both(dofs.global(e, 8 + a), dofs.global(c, 8 + b), -sigma * F.Tpm[a][b]);
            both(dofs.global(e, 8 + a), dofs.global(e, 8 + b), sigma * F.Tpp[a][b]);
          }
      }
    }
  }
  } // chunk
  const double t_fill = omp_get_wtime();
  chunks.compress(true, matrix_1, matrix_2);
  if (std::getenv("PECS_B200_SETUP_TIMING"))
    std::fprintf(stderr, "LDG::assemble_system_matrices: %d cells, %d chunks: fill %.2f s, compress %.2f s\n", mesh.n_cells,
                 n_chunks, t_fill - t_begin, omp_get_wtime() - t_fill);
}

std::string int_to_string_3(unsigned int n) {
  std::string t = std::to_string(n);
  while (t.size() < 3) t = "0" + t;
  return t;
}

void LDG::output_rescaled_results(const pecs::VtuMesh& patches_mesh, const ChargeCarrierSpace::CarrierPair& carrier_pair,
                                  const ParameterSpace::Parameters& sim_params, const double* patches,
                                  const unsigned int time_step_number, const std::string& directory) const {
  const PostProcessor postprocessor_1(sim_params, true, carrier_pair.carrier_1.name);
  const PostProcessor postprocessor_2(sim_params, true, carrier_pair.carrier_2.name);
  const size_t n = (size_t)patches_mesh.n_cells();
  const std::vector<pecs::VtuField> fields = {
      {postprocessor_1.current_name, 3, patches},
      {postprocessor_1.density_name, 1, patches + 12 * n},
      {postprocessor_2.current_name, 3, patches + 16 * n},
      {postprocessor_2.density_name, 1, patches + 28 * n},
  };
  patches_mesh.write(directory + "/" + carrier_pair.material_name + int_to_string_3(time_step_number) + ".vtu", fields);
}

} // namespace LDG_System
