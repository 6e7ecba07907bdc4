// LDG.hpp -- LDG_System::LDG: one-time assembly of the constant LDG matrices.
//
// Host-side mirror of reference include/LDG.hpp:160-441 / source/LDG.cpp:40-678 for degree 1.
// The reference builds dense 12x12 local matrices through FEValues loops and a sequential face loop; here the
// same bilinear forms are assembled in block form
//        [ mu^-1 A      G            ]   rows: current test functions (Jx, Jy blocks)
//        [ -G^T + ...   s/dt M + C   ]   rows: density test functions
// from three 4x4 cell tables (M, Dx, Dy) and 4x4 face-trace mass tables, which is also the form the device
// solver setup wants.  Bilinear forms:
//   cell      : (s/dt) v u + mu^-1 p.q - (div p) u - grad v . q                       LDG.cpp:147-171
//   Dirichlet : v ( n.q + (tau/h) u )                                                  LDG.cpp:220-232
//   other bdry: (p.n) u                                                                LDG.cpp:261-271
//   interior  : central flux + beta upwinding (beta = (1,1)/sqrt 2) + penalty tau/min(h,h')   LDG.cpp:429-622
// Hanging faces are integrated sub-face by sub-face from the coarse side with the coarse cell's basis evaluated
// on the sub-face (the intended integral; the reference reads a stale evaluator there).
#pragma once
#include "Carrier.hpp"
#include "Csr.hpp"
#include "DoFTables.hpp"
#include "PostProcessor.hpp"
#include "Triangulation.hpp"

namespace LDG_System {

class LDG {
public:
  // M_ij = (1/dt) int v u   on the density block (reference LDG.cpp:40-83)
  pecs::CsrMatrix assemble_mass_matrix(const pecs::MeshTables& mesh, double delta_t) const;

  // both carriers of a pair at once: they differ only in the mobility (reference LDG.cpp:85-678)
  void assemble_system_matrices(const pecs::MeshTables& mesh, int dirichlet_id, double scaled_mobility_1,
                                double scaled_mobility_2, double delta_t, double transient_or_steady, double penalty,
                                pecs::CsrMatrix& matrix_1, pecs::CsrMatrix& matrix_2) const;

  // reference LDG.cpp:1195-1232: file "<material_name><NNN>.vtu" holding "<carrier> Current" (vector) and
  // "<carrier> Density" of both carriers of the pair on one patch per cell.  `patches` are the rescaled patch values
  // the device produced (pecs_output_snapshot layout: current_1 | density_1 | current_2 | density_2).
  void output_rescaled_results(const pecs::VtuMesh& patches_mesh, const ChargeCarrierSpace::CarrierPair& carrier_pair,
                               const ParameterSpace::Parameters& sim_params, const double* patches,
                               const unsigned int time_step_number, const std::string& directory = ".") const;
};

// Utilities::int_to_string(n, 3)
std::string int_to_string_3(unsigned int n);

} // namespace LDG_System
