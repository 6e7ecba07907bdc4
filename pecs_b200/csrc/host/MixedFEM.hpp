// MixedFEM.hpp -- MixedPoisson::MixedFEM: one-time assembly of the constant mixed-FEM Poisson matrix.
//
// Host-side mirror of reference include/MixedFEM.hpp:118-257 / source/MixedFEM.cpp:29-164 for RT0 x DGQ0:
//   P_ij = int ( eps_r^-1 psi_i . psi_j  -  (div psi_i) phi_j  -  lambda^2 phi_i (div psi_j) )       
// with eps_r chosen by the cell's material id, assembled straight into the constraint-condensed matrix
// (ConstraintMatrix::distribute_local_to_global in reference source/SolarCell.cpp:408-417).
#pragma once
#include "Csr.hpp"
#include "DoFTables.hpp"
#include "PostProcessor.hpp"
#include "Triangulation.hpp"

namespace MixedPoisson {

class MixedFEM {
public:
  // materials 0,1 -> semiconductor permittivity, 2,3 -> electrolyte permittivity (reference MixedFEM.hpp:226-236)
  pecs::CsrMatrix assemble_Poisson_matrix(const pecs::MeshTables& mesh, const pecs::PoissonDofs& dofs,
                                          double semi_permittivity, double elec_permittivity,
                                          double scaled_debye_length) const;

  // reference MixedFEM.cpp:297-320: file "Poisson-<NNN>.vtu" with "Field" (vector) and "Potential";
  // `patches` = pecs_output_snapshot layout field | potential
  void output_rescaled_results(const pecs::VtuMesh& patches_mesh, const double* patches,
                               const ParameterSpace::Parameters& sim_params, const unsigned int time_step_number,
                               const std::string& directory = ".") const;
};

// scatter a local vector through the constraints (ConstraintMatrix::distribute_local_to_global, vector form)
void distribute_local_to_global(const pecs::PoissonDofs& dofs, const double* local, const int* local_dofs, int n,
                                double* global);
// x[constrained] = weight * x[master] (or 0): ConstraintMatrix::distribute
void distribute(const pecs::PoissonDofs& dofs, double* x);

} // namespace MixedPoisson
