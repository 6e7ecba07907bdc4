// selftest.cpp -- libpecs_b200_selftest.so: CPU CHECKERS of the product's device formulas and setup tables.
// TEST INFRASTRUCTURE, a separate shared library: nothing here is in libpecs_b200.so, nothing in
// libpecs_b200.so can call it, and the per-step path of the product has no CPU implementation at all.
//   * pecs_solarcell_selftest_carrier_rhs / _poisson_rows / _field_patches: the inline arithmetic of the production
//     kernels (csrc/rhs_math.hpp, csrc/Assembly.hpp -- the very headers the CUDA kernels are compiled from) evaluated on
//     the host, so that `pytest -m "not gpu"` checks the formulas against the oracle;
//   * pecs_solarcell_selftest_direct_solve: the host reference of the two solve sweeps on the host-factorised tables,
//     which checks plan + factor tables against the matrix without a GPU.
// Declared in include/pecs_b200_selftest.h; links against libpecs_b200.so for the host classes.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../../include/pecs_b200_selftest.h"
#include "../Assembly.hpp"
#include "../host/SchurReduction.hpp"
#include "../host/SolverSetup.hpp"
#include "../host/SparseDirect.hpp"
#include "../host/capi_internal.hpp"
#include "../rhs_math.hpp"

namespace pecs {
namespace selftest {

// Host reference of the two solve sweeps (validates plan + factor tables).
void solve_host(const SolvePlan& plan, const std::vector<double>& fwd, const std::vector<double>& bwd, const double* b,
                double* x) {
  const int n = plan.n;
  std::vector<double> w(n), xp(n), cbuf((size_t)std::max<int64_t>(plan.upd_entries, 1), 0.0);
  for (int i = 0; i < n; ++i) w[plan.perm[i]] = b[i];
  for (int d = (int)plan.levels.size() - 1; d >= 0; --d)
    for (int f : plan.levels[d]) {
      const Front& F = plan.fronts[f];
      const int np = F.np, nb = F.nb;
      // finalise the pivot right-hand side with what the children eliminated into it
      for (int c = 0; c < 2; ++c)
        if (F.cbuf_off[c] >= 0)
          for (int l = 0; l < np; ++l) w[F.p0 + l] -= cbuf[(size_t)F.cbuf_off[c] + l];
      if (F.parent < 0) continue;
      const Front& P = plan.fronts[F.parent];
      double* out = cbuf.data() + P.cbuf_off[F.which_child];
      const int* omap = plan.out_map.data() + F.bd_off;
      for (int i = 0; i < nb; ++i) {
        double carry = 0.0;
        for (int c = 0; c < 2; ++c)
          if (F.cbuf_off[c] >= 0) carry += cbuf[(size_t)F.cbuf_off[c] + np + i];
        double s = 0;
        for (int j = 0; j < np; ++j) s += fwd[(size_t)F.fwd.index(i, j)] * w[F.p0 + j];
        out[omap[i]] = carry + s;
      }
    }
  for (size_t d = 0; d < plan.levels.size(); ++d)
    for (int f : plan.levels[d]) {
      const Front& F = plan.fronts[f];
      const int np = F.np, nb = F.nb;
      const int* bd = plan.bd(F);
      for (int i = 0; i < np; ++i) {
        double s = 0;
        for (int j = 0; j < np; ++j) s += bwd[(size_t)F.bwd.index(i, j)] * w[F.p0 + j];
        for (int j = 0; j < nb; ++j) s += bwd[(size_t)F.bwd.index(i, np + j)] * xp[bd[j]];
        xp[F.p0 + i] = s;
      }
    }
  for (int i = 0; i < n; ++i) x[i] = xp[plan.perm[i]];
}


namespace {
struct CarrierRef {
  const MeshTables* mesh;
  const CsrMatrix* A;
};
CarrierRef carrier_ref(SOLARCELL::SolarCellProblem& s, int which) {
  if (which < 0 || which > 3) throw StatusError(PECS_ERR_INVALID, "system selector must be 0..4");
  const bool semi = which <= 1;
  const ChargeCarrierSpace::CarrierPair& pair = semi ? s.electron_hole_pair : s.redox_pair;
  return {semi ? &s.semiconductor_triangulation.tables() : &s.electrolyte_triangulation.tables(),
          (which % 2 == 0) ? &pair.carrier_1.system_matrix : &pair.carrier_2.system_matrix};
}
} // namespace

// host reference of the complete solve of system `which` exactly as the device does it (Schur reduction for the
// carriers unless disabled, nested-dissection tables, two sweeps)
void solve_system_host(SOLARCELL::SolarCellProblem& s, int which, int leaf_nodes, const double* b, double* x) {
  std::vector<double> fwd, bwd;
  if (which == PECS_POISSON || !schur_reduction_enabled()) {
    const SolvePlan plan = plan_for_system(s, which, leaf_nodes);
    const CsrMatrix& A = which == PECS_POISSON ? s.Poisson_object.system_matrix : *carrier_ref(s, which).A;
    factorize_host(plan, A, fwd, bwd);
    solve_host(plan, fwd, bwd, b, x);
    return;
  }
  const CarrierRef ref = carrier_ref(s, which);
  const int n = ref.mesh->n_cells, nq = 8 * n, nu = 4 * n;
  SchurReduction R;
  if (!build_schur_reduction(*ref.A, n, R)) throw StatusError(PECS_ERR_INTERNAL, "carrier (q,q) block couples cells");
  pecs_domain_desc d{};
  d.n_cells = n;
  d.vertices = ref.mesh->vertices.data();
  const NodeLayout L = carrier_density_nodes(d);
  const SolvePlan plan = plan_from_layout(R.S, L, leaf_nodes > 0 ? leaf_nodes : default_leaf_nodes(false));
  factorize_host(plan, R.S, fwd, bwd);
  std::vector<double> t(nu), rt(nu), q1(nq), q2(nq);
  R.T1.vmult(t.data(), b);                       // T1 r_q
  for (int i = 0; i < nu; ++i) rt[i] = b[nq + i] - t[i];
  solve_host(plan, fwd, bwd, rt.data(), x + nq); // u
  R.Ainv.vmult(q1.data(), b);
  R.T2.vmult(q2.data(), x + nq);
  for (int i = 0; i < nq; ++i) x[i] = q1[i] - q2[i];
}


} // namespace selftest
} // namespace pecs

namespace {
using SOLARCELL::SolarCellProblem;
using pecs::capi::guarded;
using pecs::capi::tria;
} // namespace

extern "C" {

pecs_status pecs_solarcell_selftest_carrier_rhs(pecs_solarcell* p, int32_t which, const double* u1, const double* u2,
                                                const double* o1, const double* o2, const double* X, double* rhs1,
                                                double* rhs2) {
  return guarded([&] {
    if (which < 0 || which > 1 || !u1 || !u2 || !X || !rhs1 || !rhs2)
      throw pecs::StatusError(PECS_ERR_INVALID, "selftest_carrier_rhs: bad argument");
    SolarCellProblem& s = *p->problem;
    const pecs::MeshTables& mesh = tria(p, which).tables();
    const std::vector<int>& to_poisson = which == 0 ? s.s_2_p_map : s.e_2_p_map;
    const std::vector<int>& face_dof = s.Poisson_object.dofs.face_dof;
    double prm[32];
    s.fill_params(prm);
    const pecs::RhsParams rp = pecs::make_rhs_params(prm, PECS_KIND_PRODUCTION, which);
    const size_t n = (size_t)mesh.n_cells;
    const size_t n_other = (size_t)tria(p, 1 - which).tables().n_cells;
    // interface neighbour of a cell of this subdomain (one interface face per cell at most)
    std::vector<int> nb_cell(n, -1), nb_face(n, 0);
    const std::vector<int>& mine_c = which == 0 ? s.semi_interface_cells : s.elec_interface_cells;
    const std::vector<int>& other_c = which == 0 ? s.elec_interface_cells : s.semi_interface_cells;
    const std::vector<int>& other_f = which == 0 ? s.elec_interface_faces : s.semi_interface_faces;
    for (size_t k = 0; k < mine_c.size(); ++k) {
      nb_cell[mine_c[k]] = other_c[k];
      nb_face[mine_c[k]] = other_f[k];
    }
    // one AssemblyScratch / CopyData per cell, exactly what a device thread holds in registers (csrc/Assembly.hpp)
    for (size_t c = 0; c < n; ++c) {
      Assembly::AssemblyScratch scratch;
      const double* vt = mesh.vtx((int)c);
      for (int a = 0; a < 4; ++a) {
        scratch.vertices.x[a] = vt[2 * a];
        scratch.vertices.y[a] = vt[2 * a + 1];
        scratch.carrier_1_density[a] = u1[8 * n + 4 * c + a];
        scratch.carrier_2_density[a] = u2[8 * n + 4 * c + a];
        scratch.Poisson_flux[a] = X[face_dof[4 * (size_t)to_poisson[c] + a]];
        scratch.neighbor_carrier_1_density[a] = scratch.neighbor_carrier_2_density[a] = 0.0;
      }
      double m[4];
      pecs::rhsmath::static_cell_integrals(scratch.vertices, rp.gen_scale != 0.0, rp.gen_scale, rp.gen_alpha, rp.gen_location, m,
                                           scratch.generation_integrals);
      // face terms exactly as cuda/rhs_kernels.cu boundary_record adds them (skipped when o1 / o2 are not given)
      scratch.faces = pecs::rhsmath::BoundaryRecord{{-1, -1, -1, -1}, nb_cell[c], nb_face[c]};
      bool boundary = false;
      for (int f = 0; f < 4; ++f)
        if (mesh.face_kind[4 * c + f] == pecs::FACE_BOUNDARY) {
          scratch.faces.id[f] = mesh.boundary_id[4 * c + f];
          boundary = true;
        }
      scratch.at_boundary = boundary && o1 && o2;
      if (scratch.at_boundary) {
        pecs::rhsmath::boundary_geometry(scratch.vertices, rp.tau, &scratch.face_geometry[0][0]);
        if (scratch.faces.nb_cell >= 0)
          for (int a = 0; a < 4; ++a) {
            scratch.neighbor_carrier_1_density[a] = o1[8 * n_other + 4 * (size_t)scratch.faces.nb_cell + a];
            scratch.neighbor_carrier_2_density[a] = o2[8 * n_other + 4 * (size_t)scratch.faces.nb_cell + a];
          }
      }
      Assembly::DriftDiffusion::CopyData data;
      Assembly::assemble_local_carrier_rhs(scratch, rp, data);
      // the "copier": cell c owns rows 4c..4c+3 of every component block (reference SolarCell.cpp:999-1035)
      for (int k = 0; k < 3; ++k)
        for (int a = 0; a < 4; ++a) {
          rhs1[4 * k * n + 4 * c + a] = data.local_carrier_1_rhs[4 * k + a];
          rhs2[4 * k * n + 4 * c + a] = data.local_carrier_2_rhs[4 * k + a];
        }
    }
  });
}
pecs_status pecs_solarcell_selftest_poisson_rows(pecs_solarcell* p, const double* const densities[4], double* phi_rows) {
  return guarded([&] {
    if (!densities || !phi_rows) throw pecs::StatusError(PECS_ERR_INVALID, "selftest_poisson_rows: bad argument");
    SolarCellProblem& s = *p->problem;
    double prm[32];
    s.fill_params(prm);
    for (int w = 0; w < (s.full_system ? 2 : 1); ++w) {
      const pecs::MeshTables& mesh = tria(p, w).tables();
      const std::vector<int>& to_poisson = w == 0 ? s.s_2_p_map : s.e_2_p_map;
      const pecs::RhsParams rp = pecs::make_rhs_params(prm, PECS_KIND_PRODUCTION, w);
      const size_t n = (size_t)mesh.n_cells;
      const double *u1 = densities[2 * w], *u2 = densities[2 * w + 1];
      if (!u1 || !u2) throw pecs::StatusError(PECS_ERR_INVALID, "selftest_poisson_rows: missing carrier vector");
      for (size_t c = 0; c < n; ++c) {
        pecs::fe::CellVerts v;
        const double* vt = mesh.vtx((int)c);
        for (int a = 0; a < 4; ++a) {
          v.x[a] = vt[2 * a];
          v.y[a] = vt[2 * a + 1];
        }
        double m[4], g[4];
        pecs::rhsmath::static_cell_integrals(v, false, 0.0, 0.0, 0.0, m, g);
        phi_rows[to_poisson[c]] = pecs::rhsmath::poisson_charge_row(rp, m, u1 + 8 * n + 4 * c, u2 + 8 * n + 4 * c);
      }
    }
  });
}
pecs_status pecs_solarcell_selftest_field_patches(pecs_solarcell* p, const double* X, double scale, double* field) {
  return guarded([&] {
    if (!X || !field) throw pecs::StatusError(PECS_ERR_INVALID, "selftest_field_patches: bad argument");
    SolarCellProblem& s = *p->problem;
    const pecs::MeshTables& mesh = s.Poisson_triangulation.tables();
    const std::vector<int>& face_dof = s.Poisson_object.dofs.face_dof;
    for (size_t c = 0; c < (size_t)mesh.n_cells; ++c) {
      pecs::fe::CellVerts v;
      const double* vt = mesh.vtx((int)c);
      double Xf[4];
      for (int a = 0; a < 4; ++a) {
        v.x[a] = vt[2 * a];
        v.y[a] = vt[2 * a + 1];
        Xf[a] = X[face_dof[4 * c + a]];
      }
      for (int a = 0; a < 4; ++a)
        pecs::rhsmath::rt0_field_at_vertex(v, Xf, a, scale, field[2 * (4 * c + a)], field[2 * (4 * c + a) + 1]);
    }
  });
}

pecs_status pecs_solarcell_selftest_direct_solve(pecs_solarcell* p, int32_t which, int32_t leaf_nodes, const double* b,
                                                 double* x) {
  return guarded([&] {
    pecs::selftest::solve_system_host(*p->problem, which, leaf_nodes, b, x);
  });
}

pecs_status pecs_solarcell_selftest_prepared_hashes(pecs_solarcell* p, int32_t which, uint64_t hashes[8]) {
  return guarded([&] {
    if (!hashes || which < 0 || which > PECS_POISSON)
      throw pecs::StatusError(PECS_ERR_INVALID, "selftest_prepared_hashes: bad argument");
    SolarCellProblem& s = *p->problem;
    auto bytes = [](const void* q, size_t n) {
      uint64_t v = 1469598103934665603ull; // FNV-1a
      const unsigned char* b = static_cast<const unsigned char*>(q);
      for (size_t i = 0; i < n; ++i) v = (v ^ b[i]) * 1099511628211ull;
      return v;
    };
    auto matrix = [&](const pecs::CsrMatrix& M) {
      return bytes(M.row_ptr.data(), M.row_ptr.size() * sizeof(int)) ^ 3 * bytes(M.col.data(), M.col.size() * sizeof(int)) ^
             5 * bytes(M.val.data(), M.val.size() * sizeof(double));
    };
    auto as_csr = [](const pecs::CsrMatrix& M) {
      pecs_csr c;
      c.n = M.n;
      c.row_ptr = M.row_ptr.data();
      c.col = M.col.data();
      c.val = M.val.data();
      return c;
    };
    pecs::PreparedSystem ps;
    if (which == PECS_POISSON) {
      pecs_poisson_desc d{};
      const pecs::MeshTables& P = s.Poisson_triangulation.tables();
      d.n_cells = P.n_cells;
      d.vertices = P.vertices.data();
      d.n_rt = s.Poisson_object.dofs.n_rt;
      d.face_dof = s.Poisson_object.dofs.face_dof.data();
      d.system_matrix = as_csr(s.Poisson_object.system_matrix);
      ps = pecs::prepare_poisson(d, s.Poisson_object.system_matrix.n, true);
    } else {
      const bool semi = which <= 1;
      const pecs::MeshTables& M = semi ? s.semiconductor_triangulation.tables() : s.electrolyte_triangulation.tables();
      const ChargeCarrierSpace::CarrierPair& pair = semi ? s.electron_hole_pair : s.redox_pair;
      pecs_domain_desc d{};
      d.n_cells = M.n_cells;
      d.vertices = M.vertices.data();
      d.system_matrix[0] = as_csr(pair.carrier_1.system_matrix);
      d.system_matrix[1] = as_csr(pair.carrier_2.system_matrix);
      ps = pecs::prepare_carrier(d, which % 2, true);
    }
    auto ell = [&](const pecs::HostEll& E) {
      return bytes(E.col.data(), E.col.size() * sizeof(int)) ^ 7 * bytes(E.val.data(), E.val.size() * sizeof(double)) ^
             (uint64_t)(E.width * 8 + E.block);
    };
    hashes[0] = matrix(ps.A) ^ 11 * ell(ps.ell_A);
    hashes[1] = matrix(ps.R.T1) ^ 11 * ell(ps.ell_T1);
    hashes[2] = matrix(ps.R.T2) ^ 11 * ell(ps.ell_T2);
    hashes[3] = matrix(ps.R.Ainv) ^ 11 * ell(ps.ell_Ainv);
    hashes[4] = bytes(ps.plan.perm.data(), ps.plan.perm.size() * sizeof(int));
    hashes[5] = bytes(ps.plan.bd_index.data(), ps.plan.bd_index.size() * sizeof(int)) ^
                3 * bytes(ps.plan.out_map.data(), ps.plan.out_map.size() * sizeof(int));
    hashes[6] = matrix(ps.Ap);
    hashes[7] = matrix(ps.Apt);
  });
}

pecs_status pecs_solarcell_selftest_ell_matvec(pecs_solarcell* p, int32_t which, int32_t table, const double* x, double* y_ell,
                                               double* y_csr, int32_t* shape) {
  return guarded([&] {
    if (!x || !y_ell || !y_csr || !shape || which < 0 || which > 3 || table < 0 || table > 3)
      throw pecs::StatusError(PECS_ERR_INVALID, "selftest_ell_matvec: bad argument");
    SolarCellProblem& s = *p->problem;
    const bool semi = which <= 1;
    const ChargeCarrierSpace::CarrierPair& pair = semi ? s.electron_hole_pair : s.redox_pair;
    const pecs::MeshTables& mesh = semi ? s.semiconductor_triangulation.tables() : s.electrolyte_triangulation.tables();
    const pecs::CsrMatrix& A = which % 2 == 0 ? pair.carrier_1.system_matrix : pair.carrier_2.system_matrix;
    pecs::SchurReduction R;
    if (!pecs::build_schur_reduction(A, mesh.n_cells, R))
      throw pecs::StatusError(PECS_ERR_INTERNAL, "carrier (q,q) block couples cells");
    const pecs::CsrMatrix& M = table == 0 ? R.S : table == 1 ? R.T1 : table == 2 ? R.Ainv : R.T2;
    // rows in a scrambled order, as the device tables of S and T1 are stored in elimination order
    std::vector<int> order(M.n);
    for (int i = 0; i < M.n; ++i) order[i] = (int)(((long long)i * 7919) % M.n);
    const bool reorder = table <= 1 && M.n % 7919 != 0;
    const pecs::HostEll E = pecs::build_ell(M, reorder ? &order : nullptr, 3);
    shape[0] = E.n;
    shape[1] = E.width;
    shape[2] = E.block;
    // exactly the arithmetic of ell_row in cuda/schur_kernels.cu, slot by slot
    for (int i = 0; i < E.n; ++i) {
      double acc = 0.0;
      for (int k = 0; k < E.width; ++k) {
        const int j = E.col[(size_t)k * E.n + i];
        if (E.block == 4) {
          const double* v = E.val.data() + (size_t)4 * k * E.n + i;
          acc += (v[0] * x[j] + v[E.n] * x[j + 1]) + (v[2 * (size_t)E.n] * x[j + 2] + v[3 * (size_t)E.n] * x[j + 3]);
        } else {
          acc += E.val[(size_t)k * E.n + i] * x[j];
        }
      }
      y_ell[i] = acc;
    }
    std::vector<double> y(M.n);
    M.vmult(y.data(), x);
    for (int i = 0; i < M.n; ++i) y_csr[i] = y[reorder ? order[i] : i];
  });
}

} // extern "C"
