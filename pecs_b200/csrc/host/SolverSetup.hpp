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
  std::vector<double> x, y;
};

// from the C-ABI tables (this is what pecs_ctx_create uses)
NodeLayout carrier_nodes(const pecs_domain_desc& d);
// one node per cell with only its 4 density unknowns (the Schur-reduced carrier system, host/SchurReduction.hpp)
NodeLayout carrier_density_nodes(const pecs_domain_desc& d);
// PECS_B200_NO_SCHUR=1 factorises the full 12-unknowns-per-cell systems instead (debugging / comparison)
bool schur_reduction_enabled();
NodeLayout poisson_nodes(const pecs_poisson_desc& d);
// recursion stops at this many nodes per leaf; PECS_B200_LEAF_NODES overrides (tuning)
int default_leaf_nodes(bool poisson);

// convenience for the host classes / CPU tests: which = 0..3 species, 4 Poisson; leaf_nodes <= 0 -> default
SolvePlan plan_for_system(SOLARCELL::SolarCellProblem& problem, int which, int leaf_nodes);
// host reference of the complete solve of system `which` exactly as the device does it (Schur reduction for the
// carriers unless disabled, nested-dissection tables, two sweeps): CPU verification of the setup tables only.
void solve_system_host(SOLARCELL::SolarCellProblem& problem, int which, int leaf_nodes, const double* b, double* x);

} // namespace pecs
