// SolverSetup.hpp -- how the five constant systems are grouped into nodes for nested dissection.
//
//   carriers: one node per cell (its 12 LDG unknowns), located at the cell centre;
//   Poisson : one node per cell = its potential + every unconstrained edge flux the cell "owns"
//             (an edge is owned by the cell that sees it as face 1 or 3, else by its only cell).  Pairing each
//             potential with fluxes of its own cell keeps the pivot blocks of the saddle-point matrix
//             [A  -B^T; -lambda^2 B  0] invertible without pivoting across fronts.  Constrained fluxes (hanging
//             children, Neumann edges) are decoupled rows; they join the node of their cell as well.
#pragma once
#include <vector>

#include "../../../include/pecs_b200.h"
#include "SparseDirect.hpp"

namespace SOLARCELL {
class SolarCellProblem;
}

namespace pecs {

struct NodeLayout {
  std::vector<int> node_of_dof;
  std::vector<int> group_of_node; // the mesh cell a graph node is bisected with; empty: node == group
  std::vector<double> x, y;       // per group
};

// from the C-ABI tables (this is what pecs_ctx_create uses)
NodeLayout carrier_nodes(const pecs_domain_desc& d);
// the Schur-reduced carrier system (host/SchurReduction.hpp): every density unknown is its own graph node, bisected
// through its cell -- separators are then sets of unknowns (6 per cell row), not of whole cells (8 per cell row)
NodeLayout carrier_density_nodes(const pecs_domain_desc& d);
SolvePlan plan_from_layout(const CsrMatrix& A, const NodeLayout& L, int leaf_groups);
// PECS_B200_NO_SCHUR=1 factorises the full 12-unknowns-per-cell systems instead (debugging / comparison)
bool schur_reduction_enabled();
NodeLayout poisson_nodes(const pecs_poisson_desc& d);
// Plan of the saddle-point Poisson matrix [A  -B^T; -lambda^2 B  0] (RT0 fluxes on edges, one potential per cell).
// Default: EDGE separators.  The cells of a region are bisected geometrically and the separator is the set of edge
// fluxes shared by the two halves -- one unknown per cell row instead of the three of a whole cell.  A region whose
// boundary fluxes all belong to ancestors is a pure Neumann problem: its leading block is singular by the constant
// potential.  Therefore every region passes ONE potential per connected component up to its parent instead of
// eliminating it (a delayed pivot chosen symbolically): with it pinned the divergence block has full row rank and every
// pivot block is invertible without pivoting across fronts; the parent eliminates all but one of the potentials it
// receives per component of its own region, the root eliminates the rest.
// PECS_B200_POISSON_CELL_NODES=1 selects the older vertex-separator plan on poisson_nodes() (comparison).
SolvePlan poisson_plan(const CsrMatrix& A, const pecs_poisson_desc& d, int leaf_cells);
// recursion stops at this many nodes per leaf; PECS_B200_LEAF_NODES overrides (tuning)
int default_leaf_nodes(bool poisson);

// convenience for the host classes / CPU tests: which = 0..3 species, 4 Poisson; leaf_nodes <= 0 -> default
SolvePlan plan_for_system(SOLARCELL::SolarCellProblem& problem, int which, int leaf_nodes);

} // namespace pecs
