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

// keeps a value in a register as it is: the compiler cannot contract the product that made it with a later sum
#if defined(__x86_64__)
#define PECS_ROUNDED(x) __asm__("" : "+x"(x))
#elif defined(__aarch64__)
#define PECS_ROUNDED(x) __asm__("" : "+w"(x))
#else
#define PECS_ROUNDED(x) __asm__("" : "+m"(x))
#endif

namespace {

// The 12 rows of ONE cell of the two carrier matrices of a pair, gathered: every term that lands in these rows -- the
// cell integrals, the cell's boundary faces, the interior faces the cell works on, and the interior faces a NEIGHBOUR
// works on (same level: the lower cell index; hanging: the coarse side, per sub-face; reference LDG.cpp:283-426 visits
// every face once, from that side) -- added into dense 12 x 12 blocks, one per column cell.  Terms are added in the
// order a sequential cell loop inserts them (worker cell ascending; within the cell itself: cell part, then faces
// 0..3), so every entry is the same floating-point sum as in a sequential assembly, whatever runs in parallel.
struct CellRows {
  static constexpr int kMaxBlocks = 9; // the cell itself + at most two cells across each face
  int n_blocks = 0;
  int cell_of_block[kMaxBlocks];
  double v1[12][12 * kMaxBlocks], v2[12][12 * kMaxBlocks];

  int block(int cell) {
    for (int k = 0; k < n_blocks; ++k)
      if (cell_of_block[k] == cell) return k;
    cell_of_block[n_blocks] = cell;
    for (int r = 0; r < 12; ++r)
      for (int j = 0; j < 12; ++j) v1[r][12 * n_blocks + j] = v2[r][12 * n_blocks + j] = 0.0;
    return n_blocks++;
  }
  // a and b are ROUNDED terms (no fused multiply-add of the product that made them with this sum): an entry is then
  // the same sum of the same terms as in a sequential triplet assembly
  void add(int row, int blk, int col, double a, double b) {
    PECS_ROUNDED(a);
    PECS_ROUNDED(b);
    v1[row][12 * blk + col] += a;
    v2[row][12 * blk + col] += b;
  }
  void both(int row, int blk, int col, double v) { add(row, blk, col, v, v); }
};

struct LdgTerms {
  const MeshTables& mesh;
  int dirichlet_id;
  double mu1, mu2, mass_scale, penalty;

  // an interior face seen from its worker (minus) side
  struct Face {
    int worker, f, sub, n_parts, plus;
    double sigma;
  };

  void face_of_worker(int w, int f, int sub, Face& F) const {
    const int kind = mesh.face_kind[4 * w + f];
    const double h = pecs::fe::cell_diameter(load_verts(mesh, w));
    F.worker = w;
    F.f = f;
    F.sub = sub;
    if (kind == pecs::FACE_SAME_LEVEL) {
      F.n_parts = 1;
      F.plus = mesh.neighbor[4 * w + f];
      F.sigma = penalty / std::min(h, mesh.diameter(F.plus));
    } else { // FACE_HAS_CHILDREN
      F.n_parts = 2;
      F.plus = sub == 0 ? mesh.neighbor[4 * w + f] : mesh.neighbor2[4 * w + f];
      F.sigma = penalty / std::min(h, mesh.nb_parent_diameter[4 * w + f]);
    }
  }

  // rows of cell c from an interior face; c is the worker (minus side) or the plus side
  void interior(const Face& I, int c, CellRows& R) const {
    const double beta[2] = {1.0 / std::sqrt(2.0), 1.0 / std::sqrt(2.0)};
    const double len = 1.0 / I.n_parts;
    const FaceTables F = face_tables(load_verts(mesh, I.worker), I.f, I.sub * len, len, true, I.f ^ 1);
    const double n[2] = {F.nx, F.ny};
    const double sigma = I.sigma;
    if (c == I.worker) {
      const int own = R.block(c), other = R.block(I.plus);
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
          for (int d = 0; d < 2; ++d) {
            const double hp = 0.5 * n[d] + beta[d], hm = 0.5 * n[d] - beta[d];
            R.both(4 * d + a, own, 8 + b, hp * F.Tmm[a][b]);
            R.both(4 * d + a, other, 8 + b, hm * F.Tmp[a][b]);
            R.both(8 + a, own, 4 * d + b, hm * F.Tmm[a][b]);
            R.both(8 + a, other, 4 * d + b, hp * F.Tmp[a][b]);
          }
          R.both(8 + a, own, 8 + b, sigma * F.Tmm[a][b]);
          R.both(8 + a, other, 8 + b, -sigma * F.Tmp[a][b]);
        }
    } else {
      const int own = R.block(c), other = R.block(I.worker);
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
          for (int d = 0; d < 2; ++d) {
            const double hp = 0.5 * n[d] + beta[d], hm = 0.5 * n[d] - beta[d];
            R.both(4 * d + a, other, 8 + b, -hp * F.Tpm[a][b]);
            R.both(4 * d + a, own, 8 + b, -hm * F.Tpp[a][b]);
            R.both(8 + a, other, 4 * d + b, -hm * F.Tpm[a][b]);
            R.both(8 + a, own, 4 * d + b, -hp * F.Tpp[a][b]);
          }
          R.both(8 + a, other, 8 + b, -sigma * F.Tpm[a][b]);
          R.both(8 + a, own, 8 + b, sigma * F.Tpp[a][b]);
        }
    }
  }

  // what cell c itself inserts into its own rows: cell integrals, boundary faces, the interior faces it works on
  void own_terms(int c, CellRows& R) const {
    const CellVerts v = load_verts(mesh, c);
    const double h = pecs::fe::cell_diameter(v);
    double M[4][4], Dx[4][4], Dy[4][4];
    cell_tables(v, M, Dx, Dy);
    const int own = R.block(c);
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        R.add(a, own, b, M[a][b] / mu1, M[a][b] / mu2);
        R.add(4 + a, own, 4 + b, M[a][b] / mu1, M[a][b] / mu2);
        R.both(a, own, 8 + b, -Dx[a][b]); // -(div p) u
        R.both(4 + a, own, 8 + b, -Dy[a][b]);
        R.both(8 + a, own, b, -Dx[a][b]); // -grad v . q
        R.both(8 + a, own, 4 + b, -Dy[a][b]);
        R.both(8 + a, own, 8 + b, mass_scale * M[a][b]);
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
                R.both(8 + a, own, 4 * d + b, n[d] * T); // v n.q
              else
                R.both(4 * d + a, own, 8 + b, n[d] * T); // (p.n) u
            }
            if (dirichlet) R.both(8 + a, own, 8 + b, (penalty / h) * T);
          }
      } else if ((kind == pecs::FACE_SAME_LEVEL && mesh.neighbor[4 * c + f] > c) || kind == pecs::FACE_HAS_CHILDREN) {
        const int n_parts = kind == pecs::FACE_SAME_LEVEL ? 1 : 2;
        for (int sub = 0; sub < n_parts; ++sub) {
          Face I;
          face_of_worker(c, f, sub, I);
          interior(I, c, R);
        }
      }
    }
  }

  void rows_of_cell(int c, CellRows& R) const {
    R.n_blocks = 0;
    R.block(c);
    // faces a neighbour works on, in the order of the neighbours' cell indices
    Face foreign[4];
    int n_foreign = 0;
    for (int f = 0; f < 4; ++f) {
      const int kind = mesh.face_kind[4 * c + f];
      if (kind == pecs::FACE_SAME_LEVEL && mesh.neighbor[4 * c + f] < c)
        face_of_worker(mesh.neighbor[4 * c + f], f ^ 1, 0, foreign[n_foreign++]);
      else if (kind == pecs::FACE_COARSER)
        face_of_worker(mesh.neighbor[4 * c + f], f ^ 1, mesh.neighbor2[4 * c + f], foreign[n_foreign++]);
    }
    std::sort(foreign, foreign + n_foreign, [](const Face& x, const Face& y) {
      return x.worker != y.worker ? x.worker < y.worker : (x.f != y.f ? x.f < y.f : x.sub < y.sub);
    });
    int k = 0;
    for (; k < n_foreign && foreign[k].worker < c; ++k) interior(foreign[k], c, R);
    own_terms(c, R);
    for (; k < n_foreign; ++k) interior(foreign[k], c, R);
  }
};

} // namespace

void LDG::assemble_system_matrices(const MeshTables& mesh, int dirichlet_id, double mu1, double mu2, double delta_t,
                                   double transient_or_steady, double penalty, CsrMatrix& matrix_1,
                                   CsrMatrix& matrix_2) const {
  const double t_begin = omp_get_wtime();
  const int n = mesh.n_cells;
  const CarrierDofs dofs{n};
  const LdgTerms terms{mesh, dirichlet_id, mu1, mu2, transient_or_steady / delta_t, penalty};
  // rows are independent: pass 1 counts the entries of every row (entries whose two sums are both exactly zero are not
  // stored), pass 2 evaluates the rows again and writes them in place.  Columns of a row ascend: [Jx | Jy | rho], and
  // inside each component by cell.
  auto emit = [&](int c, const CellRows& R, bool write) {
    int order[CellRows::kMaxBlocks];
    for (int k = 0; k < R.n_blocks; ++k) order[k] = k;
    std::sort(order, order + R.n_blocks, [&](int x, int y) { return R.cell_of_block[x] < R.cell_of_block[y]; });
    for (int r = 0; r < 12; ++r) {
      const int row = dofs.global(c, r);
      int at = write ? matrix_1.row_ptr[row] : 0;
      for (int comp = 0; comp < 3; ++comp)
        for (int q = 0; q < R.n_blocks; ++q) {
          const int k = order[q];
          for (int j = 0; j < 4; ++j) {
            const double a = R.v1[r][12 * k + 4 * comp + j], b = R.v2[r][12 * k + 4 * comp + j];
            if (a == 0.0 && b == 0.0) continue;
            if (write) {
              matrix_1.col[at] = dofs.global(R.cell_of_block[k], 4 * comp + j);
              matrix_1.val[at] = a;
              matrix_2.val[at] = b;
            }
            ++at;
          }
        }
      if (!write) matrix_1.row_ptr[(size_t)row + 1] = at;
    }
  };
  matrix_1.n = matrix_2.n = dofs.n_dofs();
  matrix_1.row_ptr.assign((size_t)dofs.n_dofs() + 1, 0);
#pragma omp parallel
  {
    CellRows R;
#pragma omp for schedule(static)
    for (int c = 0; c < n; ++c) {
      terms.rows_of_cell(c, R);
      emit(c, R, false);
    }
  }
  for (int i = 0; i < dofs.n_dofs(); ++i) matrix_1.row_ptr[i + 1] += matrix_1.row_ptr[i];
  const size_t nnz = (size_t)matrix_1.row_ptr[dofs.n_dofs()];
  matrix_1.col.resize(nnz);
  matrix_1.val.resize(nnz);
  matrix_2.val.resize(nnz);
  const double t_count = omp_get_wtime();
#pragma omp parallel
  {
    CellRows R;
#pragma omp for schedule(static)
    for (int c = 0; c < n; ++c) {
      terms.rows_of_cell(c, R);
      emit(c, R, true);
    }
  }
  matrix_2.row_ptr = matrix_1.row_ptr;
  matrix_2.col = matrix_1.col;
  if (std::getenv("PECS_B200_SETUP_TIMING"))
    std::fprintf(stderr, "LDG::assemble_system_matrices: %d cells: count %.2f s, fill %.2f s\n", n, t_count - t_begin,
                 omp_get_wtime() - t_count);
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
