// Grid.cpp -- see Grid.hpp.
#include "Grid.hpp"

#include <future>

#include <iostream>

namespace Grid_Maker {

using pecs::Point2;
using pecs::Triangulation;

namespace {
// apply `tag(face_centre) -> new id or -1 (keep)` to every boundary face
template <class F>
void for_each_boundary_face(Triangulation& tria, F&& tag) {
  pecs::MeshTables& t = tria.tables();
  for (int c = 0; c < t.n_cells; ++c)
    for (int f = 0; f < 4; ++f)
      if (t.face_kind[4 * c + f] == pecs::FACE_BOUNDARY) {
        const int id = tag(t.face_center(c, f));
        if (id >= 0) t.boundary_id[4 * c + f] = id;
      }
}

Triangulation one_cell(Point2 bl, Point2 br, Point2 tl, Point2 tr, int material) {
  Triangulation t;
  t.create({{bl, br, tl, tr}}, {material});
  return t;
}
} // namespace

Grid::Grid(const ParameterSpace::Parameters& sim_params) {
  scaled_domain_height = sim_params.scaled_domain_height;
  scaled_domain_length = sim_params.scaled_domain_length;
  scaled_radius_one = sim_params.scaled_radius_one;
  scaled_radius_two = sim_params.scaled_radius_two;
  scaled_boundary_layer = sim_params.scaled_boundary_layer;
  n_global_refine = sim_params.n_global_refine;
  n_local_refine = sim_params.n_local_refine;
  insulated = sim_params.insulated;
  schottky = sim_params.schottky_status;
  use_boundary_layer = false;
  // reference Grid.cpp:24-40
  if (n_local_refine == 0) {
    scaled_boundary_layer = 0.0;
  } else if (scaled_boundary_layer > 0) {
    use_boundary_layer = true;
  } else {
    std::cerr << "Boundary layer & n_local_refine need to be >= 0\n";
  }
}

void Grid::make_semiconductor_grid(Triangulation& triangulation) {
  // trapezoids: bottom radius two, top radius one (reference Grid.cpp:143-149, 176-182)
  const double xb = scaled_radius_two - scaled_boundary_layer, xt = scaled_radius_one - scaled_boundary_layer;
  const double H = scaled_domain_height;
  Triangulation bulk = one_cell({0, 0}, {xb, 0}, {0, H}, {xt, H}, semiconductor_id);
  if (!use_boundary_layer) {
    triangulation = bulk;
    return;
  }
  Triangulation layer =
      one_cell({xb, 0}, {scaled_radius_two, 0}, {xt, H}, {scaled_radius_one, H}, semi_boundary_layer_id);
  triangulation = Triangulation::merge(bulk, layer);
}

void Grid::make_electrolyte_grid(Triangulation& triangulation) {
  // reference Grid.cpp:239-246, 269-276
  const double xb = scaled_radius_two + scaled_boundary_layer, xt = scaled_radius_one + scaled_boundary_layer;
  const double H = scaled_domain_height, L = scaled_domain_length;
  Triangulation bulk = one_cell({xb, 0}, {L, 0}, {xt, H}, {L, H}, electrolyte_id);
  if (!use_boundary_layer) {
    triangulation = bulk;
    return;
  }
  Triangulation layer =
      one_cell({scaled_radius_two, 0}, {xb, 0}, {scaled_radius_one, H}, {xt, H}, elec_boundary_layer_id);
  triangulation = Triangulation::merge(layer, bulk);
}

void Grid::make_merged_grid(const Triangulation& semiconductor_triang, const Triangulation& electrolyte_triang,
                            Triangulation& merged_triangulation) {
  merged_triangulation = Triangulation::merge(semiconductor_triang, electrolyte_triang);
}

void Grid::make_grids(Triangulation& semiconductor_triang, Triangulation& electrolyte_triang,
                      Triangulation& Poisson_triang, const bool& full_system) {
  make_semiconductor_grid(semiconductor_triang);
  make_electrolyte_grid(electrolyte_triang);
  if (full_system)
    make_merged_grid(semiconductor_triang, electrolyte_triang, Poisson_triang);
  else
    make_semiconductor_grid(Poisson_triang);

  // from here on the three triangulations share nothing: refined and marked side by side
  // (boundary-layer cells are refined n_local_refine more times, reference Grid.cpp:70-106)
  auto finish = [&](Triangulation& t, std::vector<int> layer_ids, bool schottky_edge) {
    t.refine_global((int)n_global_refine);
    for (unsigned int r = 0; r < n_local_refine; ++r) t.refine_material(layer_ids);
    make_Dirichlet_boundaries(t);
    if (insulated) make_Neumann_boundaries(t);
    if (schottky && schottky_edge) make_Schottky_boundaries(t);
  };
  std::future<void> electrolyte = std::async(std::launch::async, [&] { finish(electrolyte_triang, {elec_boundary_layer_id}, false); });
  std::future<void> poisson =
      std::async(std::launch::async, [&] { finish(Poisson_triang, {semi_boundary_layer_id, elec_boundary_layer_id}, true); });
  finish(semiconductor_triang, {semi_boundary_layer_id}, true);
  electrolyte.get();
  poisson.get();
}

void Grid::make_Dirichlet_boundaries(Triangulation& triangulation) {
  // every outer (non-interface) boundary face; exact == on coordinates as in reference Grid.cpp:361-364
  const double L = scaled_domain_length, H = scaled_domain_height;
  for_each_boundary_face(triangulation, [&](Point2 c) {
    return (c.x == 0.0 || c.x == L || c.y == 0.0 || c.y == H) ? (int)Dirichlet : -1;
  });
}

void Grid::make_Neumann_boundaries(Triangulation& triangulation) {
  // reference Grid.cpp:374-428.  Note the bottom test uses radius ONE although the bottom radius is
  // radius two; kept as is.
  const double H = scaled_domain_height, r1 = scaled_radius_one;
  for_each_boundary_face(triangulation, [&](Point2 c) {
    int id = -1;
    if (c.y == H) id = (c.x > r1) ? Neumann : Dirichlet;
    if (c.y == 0.0) id = (c.x > r1) ? Neumann : Dirichlet;
    if (c.x == 0.0) id = Neumann;
    return id;
  });
}

void Grid::make_Schottky_boundaries(Triangulation& triangulation) {
  // the whole top edge, reference Grid.cpp:430-461
  const double H = scaled_domain_height;
  for_each_boundary_face(triangulation, [&](Point2 c) { return (c.y == H) ? (int)Schottky : -1; });
}

namespace {
Triangulation unit_square(int n_global_refine) {
  Triangulation t = one_cell({0, 0}, {1, 0}, {0, 1}, {1, 1}, semiconductor_id);
  t.refine_global(n_global_refine);
  return t;
}
} // namespace

void Grid::make_test_grid(Triangulation& triangulation, const int& n_refine) {
  // top/bottom Neumann, left/right Dirichlet (reference Grid.cpp:464-504)
  triangulation = unit_square(n_refine);
  for_each_boundary_face(triangulation, [](Point2 c) { return (c.y == 0 || c.y == 1.0) ? (int)Neumann : (int)Dirichlet; });
}

void Grid::make_test_tran_grid(Triangulation& triangulation, const int& n_refine) {
  // x==1 keeps id 0 (Interface = Robin), top/bottom Neumann, left Dirichlet (reference Grid.cpp:530-566)
  triangulation = unit_square(n_refine);
  for_each_boundary_face(triangulation, [](Point2 c) {
    int id = -1;
    if (c.x != 1.0) id = Dirichlet;
    if (c.y == 0 || c.y == 1.0) id = Neumann;
    return id;
  });
}

void Grid::make_DD_Poisson_grid(Triangulation& triangulation, const int& n_refine) {
  // all Dirichlet (reference Grid.cpp:568-599)
  triangulation = unit_square(n_refine);
  for_each_boundary_face(triangulation, [](Point2) { return (int)Dirichlet; });
}

void Grid::refine_test_grid(Triangulation& triangulation, const unsigned int& local_refine) {
  // reference Grid.cpp:506-528; boundary ids must be re-tagged by the caller afterwards
  for (unsigned int i = 0; i < local_refine; ++i) triangulation.refine_near({0.5, 0.5}, 0.2);
}

} // namespace Grid_Maker
