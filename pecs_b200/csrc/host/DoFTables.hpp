// DoFTables.hpp -- global degree-of-freedom numbering and constraints as flat tables.
//
// Replaces dealii::DoFHandler + DoFRenumbering::component_wise + ConstraintMatrix for the two finite
// element spaces the reference uses with degree 1 (reference source/SolarCell.cpp:23-31):
//   carrier : [DGQ1]^2 x DGQ1      -> 12 dofs per cell, blocks [Jx | Jy | rho], 4 nodal values per cell per
//             block, cell-major inside a block (reference source/CarrierPair.cpp:29-33); no constraints.
//   Poisson : RT0 x DGQ0           -> one flux dof per active edge + one potential dof per cell, blocks
//             [RT | Phi] (reference source/Poisson.cpp:25-28); constraints = hanging edges (child = 1/2 parent)
//             and zero normal flux on Neumann edges (reference source/Poisson.cpp:33-49).
#pragma once
#include <vector>

#include "Triangulation.hpp"

namespace pecs {

struct CarrierDofs {
  int n_cells = 0;
  int n_dofs() const { return 12 * n_cells; }
  // local index i in 0..11 (Jx0-3, Jy0-3, rho0-3) of cell c -> global index
  int global(int c, int i) const { return (i / 4) * 4 * n_cells + 4 * c + (i % 4); }
};

struct ConstraintLine {
  int dof;       // constrained dof
  int master;    // -1: dof = 0 (Neumann); else dof = weight * x[master]
  double weight;
};

struct PoissonDofs {
  int n_cells = 0;
  int n_rt = 0;                  // number of flux dofs; potential dof of cell c is n_rt + c
  std::vector<int> face_dof;     // [n_cells][4]
  std::vector<ConstraintLine> constraints;
  std::vector<int> constraint_of; // [n_dofs] index into constraints or -1
  int n_dofs() const { return n_rt + n_cells; }
  int phi_dof(int c) const { return n_rt + c; }
};

// Flux dofs are numbered in the order deal.II's distribute_dofs + component_wise would leave them:
// by first-visiting active cell, then face number.  neumann_id: boundary id whose flux dofs are pinned to 0.
PoissonDofs build_poisson_dofs(const MeshTables& mesh, int neumann_id);

} // namespace pecs
