// DoFTables.cpp -- see DoFTables.hpp.
#include "DoFTables.hpp"

namespace pecs {

PoissonDofs build_poisson_dofs(const MeshTables& mesh, int neumann_id) {
  PoissonDofs d;
  d.n_cells = mesh.n_cells;
  d.face_dof.assign(4 * (size_t)mesh.n_cells, -1);
  int next = 0;
  for (int c = 0; c < mesh.n_cells; ++c)
    for (int f = 0; f < 4; ++f) {
      int& slot = d.face_dof[4 * c + f];
      if (slot >= 0) continue;
      slot = next++;
      if (mesh.face_kind[4 * c + f] == FACE_SAME_LEVEL) {
        // the neighbour sees the same edge through its opposite face (all cells share one orientation)
        const int nb = mesh.neighbor[4 * c + f];
        d.face_dof[4 * nb + (f ^ 1)] = slot;
      }
    }
  d.n_rt = next;
  d.constraint_of.assign((size_t)d.n_dofs(), -1);
  for (int c = 0; c < mesh.n_cells; ++c)
    for (int f = 0; f < 4; ++f) {
      const int kind = mesh.face_kind[4 * c + f];
      const int dof = d.face_dof[4 * c + f];
      if (kind == FACE_COARSER) {
        // child edge carries half of the parent edge's flux
        const int nb = mesh.neighbor[4 * c + f];
        d.constraint_of[dof] = (int)d.constraints.size();
        d.constraints.push_back({dof, d.face_dof[4 * nb + (f ^ 1)], 0.5});
      } else if (kind == FACE_BOUNDARY && mesh.boundary_id[4 * c + f] == neumann_id) {
        d.constraint_of[dof] = (int)d.constraints.size();
        d.constraints.push_back({dof, -1, 0.0});
      }
    }
  return d;
}

} // namespace pecs
