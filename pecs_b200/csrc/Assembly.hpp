// Assembly.hpp -- Assembly::AssemblyScratch, Assembly::DriftDiffusion::CopyData, Assembly::Poisson::CopyData.
//
// The reference's thread-private assembly buffers (reference include/Assembly.hpp:47-278, source/Assembly.cpp:9-208):
// one AssemblyScratch (FEValues evaluators + per-quadrature-point work arrays) and one CopyData (local right-hand sides,
// local matrices, dof indices) per TBB worker, handed from the "worker" to the "copier" of WorkStream::run.
//
// Here the per-step assembly runs one CUDA thread per cell and these objects are what that thread holds IN REGISTERS:
//   * AssemblyScratch = the INPUTS of one cell: its four vertices (the Jacobian at every quadrature point is recomputed
//     from them: there are no FEValues), the nodal densities of both carriers, the four RT0 fluxes of the matched
//     Poisson cell (reference SolarCell.cpp:1102-1110: s_2_p_map / e_2_p_map lookup, here a static index table), the
//     static generation integrals, and for a boundary cell its face record, face geometry and the interface neighbour's
//     densities.  The per-q-point arrays of the reference (old_carrier_*_density_values, electric_field_values,
//     generation_values, ...) never exist: the sum-factorised kernels contract them on the fly (csrc/rhs_math.hpp).
//   * DriftDiffusion::CopyData = the OUTPUTS: the two local right-hand sides, 12 entries each in the reference's local
//     order [Jx 0-3 | Jy 0-3 | rho 0-3].  The local matrices of the reference's struct belong to the one-time assembly
//     (host/LDG.cpp builds them in block form); local_dof_indices are implicit: cell c owns rows 4c..4c+3 of each of the
//     three component blocks (DG), which is why no copier / scatter exists.
//   * Poisson::CopyData = one scalar per carrier cell (only the DG0 potential test function is non-zero, reference
//     SolarCell.cpp:551-578) for row n_rt + poisson_cell.
// The structs are plain (host + device).  assemble_local_*_rhs below are the per-cell bodies with the reference's names
// (reference SolarCell.cpp:1073-1414, 1451-1726, 487-815); the device kernels (cuda/rhs_kernels.cu) run exactly this
// arithmetic from registers, the host build of the same header is the CPU check of it
// (tests/test_host_tables.py::test_production_rhs_arithmetic_matches_oracle_on_cpu).
#pragma once
#include "../../include/pecs_b200.h"
#include "fe.hpp"
#include "rhs_math.hpp"

namespace Assembly {

struct AssemblyScratch {
  pecs::fe::CellVerts vertices;               // FEValues::reinit(cell) shrinks to these 8 doubles
  double carrier_1_density[4], carrier_2_density[4]; // nodal values = old_carrier_*_density at the vertices
  double Poisson_flux[4];                     // RT0 dofs of the matched Poisson cell (electric_field_values come from them)
  double generation_integrals[4];             // int N_a G, static (Generation::value, reference Generation.cpp:29-44)
  // boundary cells only
  bool at_boundary = false;
  pecs::rhsmath::BoundaryRecord faces{{-1, -1, -1, -1}, -1, 0}; // boundary ids of the 4 faces, interface neighbour
  double face_geometry[4][4];                 // {n_x, n_y, ds, tau/h} per face (normals + JxW + penalty/h of the reference)
  double neighbor_carrier_1_density[4], neighbor_carrier_2_density[4]; // the other subdomain's traces (interface)
};

namespace DriftDiffusion {
struct CopyData {
  double local_carrier_1_rhs[12]; // [Jx | Jy | rho], reference Assembly.hpp:241-247
  double local_carrier_2_rhs[12];
};
} // namespace DriftDiffusion

namespace Poisson {
struct CopyData {
  double local_rhs; // the potential row; the four flux rows are static Dirichlet data (cuda: poisson_face_rhs_kernel)
};
} // namespace Poisson

// reference SolarCellProblem::assemble_local_semiconductor_rhs / assemble_local_electrolyte_rhs for ONE cell: cell terms
// (M u^{k-1} / dt + generation + drift) and, for a boundary cell, the Dirichlet / interface / Schottky face terms
PECS_HD void assemble_local_carrier_rhs(const AssemblyScratch& scratch, const pecs::RhsParams& p,
                                               DriftDiffusion::CopyData& data) {
  double* o1 = data.local_carrier_1_rhs;
  double* o2 = data.local_carrier_2_rhs;
  pecs::rhsmath::production_cell_terms(scratch.vertices.x, scratch.vertices.y, scratch.carrier_1_density,
                                       scratch.carrier_2_density, scratch.Poisson_flux, scratch.generation_integrals, p.inv_dt,
                                       p.charge1 * p.inv_eps, p.charge2 * p.inv_eps, o1, o1 + 4, o1 + 8, o2, o2 + 4, o2 + 8);
  if (p.srh) // SRH_Recombination, reference SolarCell.hpp:86-98 (0.0 there; the commented formula when switched on)
    pecs::rhsmath::srh_cell_terms(scratch.vertices.x, scratch.vertices.y, scratch.carrier_1_density, scratch.carrier_2_density,
                                  p.n_i, p.tau_n, p.tau_p, o1 + 8, o2 + 8);
  if (!scratch.at_boundary) return;
  double b[6][4] = {};
  pecs::rhsmath::boundary_terms_accumulate<PECS_KIND_PRODUCTION>(
      p, scratch.faces, scratch.face_geometry, scratch.vertices, scratch.carrier_1_density, scratch.carrier_2_density,
      scratch.neighbor_carrier_1_density, scratch.neighbor_carrier_2_density, b[0], b[1], b[2], b[3], b[4], b[5]);
  for (int a = 0; a < 4; ++a) { // same order of additions as the device: cell terms, then the face terms
    o1[a] += b[0][a];
    o1[4 + a] += b[1][a];
    o1[8 + a] += b[2][a];
    o2[a] += b[3][a];
    o2[4 + a] += b[4][a];
    o2[8 + a] += b[5][a];
  }
}

// reference assemble_local_Poisson_rhs_for_semiconductor / _for_electrolyte, cell part, for ONE carrier cell
PECS_HD void assemble_local_Poisson_rhs(const AssemblyScratch& scratch, const pecs::RhsParams& p,
                                               const double nodal_integrals[4], Poisson::CopyData& data) {
  data.local_rhs = pecs::rhsmath::poisson_charge_row(p, nodal_integrals, scratch.carrier_1_density, scratch.carrier_2_density);
}

} // namespace Assembly
