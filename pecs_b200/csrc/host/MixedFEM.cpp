// MixedFEM.cpp -- see MixedFEM.hpp.
#include "MixedFEM.hpp"

#include <stdexcept>

#include "../fe.hpp"

namespace MixedPoisson {

using pecs::CsrMatrix;
using pecs::MeshTables;
using pecs::PoissonDofs;
using pecs::TripletList;

CsrMatrix MixedFEM::assemble_Poisson_matrix(const MeshTables& mesh, const PoissonDofs& dofs, double semi_permittivity,
                                            double elec_permittivity, double scaled_debye_length) const {
  TripletList tl(dofs.n_dofs());
  tl.reserve(30 * (size_t)mesh.n_cells);
  for (int c = 0; c < mesh.n_cells; ++c) {
    const int mat = mesh.material_id[c];
    if (mat < 0 || mat > 3) throw std::runtime_error("MixedFEM: cell is neither semiconductor nor electrolyte");
    const double inv_eps = 1.0 / (mat <= 1 ? semi_permittivity : elec_permittivity);
    pecs::fe::CellVerts v;
    for (int a = 0; a < 4; ++a) {
      v.x[a] = mesh.vtx(c)[2 * a];
      v.y[a] = mesh.vtx(c)[2 * a + 1];
    }
    // local 5x5: flux-flux mass with the Piola map, flux-potential couplings are +-1 (div psi_f det J = +-1)
    double K[5][5] = {};
    for (int qy = 0; qy < 3; ++qy)
      for (int qx = 0; qx < 3; ++qx) {
        const double xi = pecs::fe::gauss_x(qx), eta = pecs::fe::gauss_x(qy);
        const double w = pecs::fe::gauss_w(qx) * pecs::fe::gauss_w(qy);
        const pecs::fe::Jac j = pecs::fe::jacobian(v, xi, eta);
        double px[4], py[4];
        pecs::fe::rt0_times_det(j, xi, eta, px, py);
        const double s = inv_eps * w / j.det; // (1/det)^2 * det * w
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) K[a][b] += (px[a] * px[b] + py[a] * py[b]) * s;
        for (int a = 0; a < 4; ++a) {
          K[a][4] += -pecs::fe::rt0_ref_div(a) * w;
          K[4][a] += -scaled_debye_length * pecs::fe::rt0_ref_div(a) * w;
        }
      }
    int g[5];
    for (int a = 0; a < 4; ++a) g[a] = dofs.face_dof[4 * c + a];
    g[4] = dofs.phi_dof(c);
    // condensed scatter: a constrained index is replaced by its master (times weight) or dropped
    for (int i = 0; i < 5; ++i) {
      const int ci = dofs.constraint_of[g[i]];
      int ri = g[i];
      double wi = 1.0;
      if (ci >= 0) {
        ri = dofs.constraints[ci].master;
        wi = dofs.constraints[ci].weight;
      }
      for (int j = 0; j < 5; ++j) {
        const int cj = dofs.constraint_of[g[j]];
        int rj = g[j];
        double wj = 1.0;
        if (cj >= 0) {
          rj = dofs.constraints[cj].master;
          wj = dofs.constraints[cj].weight;
        }
        if (ri >= 0 && rj >= 0 && K[i][j] != 0.0) tl.add(ri, rj, wi * wj * K[i][j]);
      }
      // keep the constrained row regular (its solution value is overwritten by distribute())
      if (ci >= 0) tl.add(g[i], g[i], K[i][i] != 0.0 ? K[i][i] : 1.0);
    }
  }
  return tl.compress();
}

void MixedFEM::output_rescaled_results(const pecs::VtuMesh& patches_mesh, const double* patches,
                                       const ParameterSpace::Parameters& sim_params, const unsigned int time_step_number,
                                       const std::string& directory) const {
  const PostProcessor postprocessor(sim_params, false, "noname");
  const std::vector<std::string> names = postprocessor.get_names();
  const size_t n = (size_t)patches_mesh.n_cells();
  std::string number = std::to_string(time_step_number);
  while (number.size() < 3) number = "0" + number;
  patches_mesh.write(directory + "/Poisson-" + number + ".vtu",
                     {{names[0], 3, patches}, {names[2], 1, patches + 12 * n}});
}

void distribute_local_to_global(const PoissonDofs& dofs, const double* local, const int* local_dofs, int n,
                                double* global) {
  for (int i = 0; i < n; ++i) {
    const int ci = dofs.constraint_of[local_dofs[i]];
    if (ci < 0)
      global[local_dofs[i]] += local[i];
    else if (dofs.constraints[ci].master >= 0)
      global[dofs.constraints[ci].master] += dofs.constraints[ci].weight * local[i];
  }
}

void distribute(const PoissonDofs& dofs, double* x) {
  for (const pecs::ConstraintLine& l : dofs.constraints) x[l.dof] = l.master >= 0 ? l.weight * x[l.master] : 0.0;
}

} // namespace MixedPoisson
