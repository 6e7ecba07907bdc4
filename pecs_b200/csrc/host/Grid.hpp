// Grid.hpp -- Grid_Maker::Grid: builds the semiconductor, electrolyte and Poisson meshes.
//
// Mirror of reference include/Grid.hpp:27-180 / source/Grid.cpp on top of pecs::Triangulation.
// Geometry, refinement and the boundary tagging (including the reference's exact `==` coordinate
// tests and its quirks) follow source/Grid.cpp:44-461; unit-square test grids follow
// source/Grid.cpp:464-599.
#pragma once
#include "Parameters.hpp"
#include "Triangulation.hpp"

namespace Grid_Maker {

// reference include/Grid.hpp:141-156
enum BoundaryId : int { Interface = 0, Dirichlet = 1, Neumann = 2, Schottky = 3 };
enum MaterialId : int { semiconductor_id = 0, semi_boundary_layer_id = 1, electrolyte_id = 2, elec_boundary_layer_id = 3 };

class Grid {
public:
  explicit Grid(const ParameterSpace::Parameters& sim_params);

  void make_grids(pecs::Triangulation& semiconductor_triang, pecs::Triangulation& electrolyte_triang,
                  pecs::Triangulation& Poisson_triang, const bool& full_system);

  void make_semiconductor_grid(pecs::Triangulation& triangulation);
  void make_electrolyte_grid(pecs::Triangulation& triangulation);
  void make_merged_grid(const pecs::Triangulation& semiconductor_triang, const pecs::Triangulation& electrolyte_triang,
                        pecs::Triangulation& merged_triangulation);

  void make_Dirichlet_boundaries(pecs::Triangulation& triangulation);
  void make_Neumann_boundaries(pecs::Triangulation& triangulation);
  void make_Schottky_boundaries(pecs::Triangulation& triangulation);

  // unit-square grids of the manufactured-solution tests
  void make_test_grid(pecs::Triangulation& triangulation, const int& n_global_refine);
  void make_test_tran_grid(pecs::Triangulation& triangulation, const int& n_global_refine);
  void make_DD_Poisson_grid(pecs::Triangulation& triangulation, const int& n_global_refine);
  void refine_test_grid(pecs::Triangulation& triangulation, const unsigned int& local_refine);

private:
  double scaled_domain_height, scaled_domain_length, scaled_radius_one, scaled_radius_two, scaled_boundary_layer;
  unsigned int n_global_refine, n_local_refine;
  bool use_boundary_layer, insulated, schottky;
};

} // namespace Grid_Maker
