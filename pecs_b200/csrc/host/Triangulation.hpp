// Triangulation.hpp -- quadtree-forest mesh substrate for the host side of pecs_b200.
//
// The reference keeps its meshes in dealii::Triangulation<2> (reference include/SolarCell.hpp:351-367)
// and walks them with active_cell_iterators.  deal.II does not exist in this build, and the GPU path
// wants flat structure-of-arrays tables anyway, so this is an own, minimal replacement that produces
// exactly the tables the device kernels (and the CPU oracle) consume:
//   * active cells in deal.II traversal order (level-major, creation order inside a level),
//   * 4 vertices per cell in deal.II lexicographic order (v0=(0,0) v1=(1,0) v2=(0,1) v3=(1,1)),
//   * per face (0: xi=0, 1: xi=1, 2: eta=0, 3: eta=1) the neighbour relation incl. hanging faces,
//   * material ids and boundary ids.
// Isotropic bisection with straight (bilinear) new vertices and 2:1 level balance, which is what
// Triangulation::refine_global / execute_coarsening_and_refinement do for these meshes
// (reference source/Grid.cpp:66-106).
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <utility>
#include <vector>

namespace pecs {

struct Point2 {
  double x = 0, y = 0;
};

enum FaceKind : int {
  FACE_SAME_LEVEL = 0,   // interior face, neighbour is active and on the same level
  FACE_BOUNDARY = 1,     // face lies on the boundary of this triangulation
  FACE_HAS_CHILDREN = 2, // interior face, neighbour is refined once more (this is the coarse side)
  FACE_COARSER = 3       // interior face, neighbour is one level coarser (this is the fine side)
};

// Flat tables of the ACTIVE cells; everything downstream works on this.
struct MeshTables {
  int n_cells = 0;
  std::vector<double> vertices;      // [n_cells][4][2]
  std::vector<int> material_id;      // [n_cells]
  std::vector<int> level;            // [n_cells]
  std::vector<int> face_kind;        // [n_cells][4]  FaceKind
  std::vector<int> neighbor;         // [n_cells][4]  same level / coarser: that cell; has children: child on subface 0
  std::vector<int> neighbor2;        // [n_cells][4]  has children: child on subface 1; coarser: our subface number; else -1
  std::vector<int> boundary_id;      // [n_cells][4]  valid on boundary faces, else -1
  std::vector<double> nb_parent_diameter; // [n_cells][4] has children: diameter of the refined (inactive) neighbour, else 0

  const double* vtx(int c) const { return &vertices[8 * (size_t)c]; }
  Point2 center(int c) const;
  Point2 face_center(int c, int f) const;
  double diameter(int c) const;
};

class Triangulation {
public:
  // coarse mesh: cells given by 4 vertex coordinates each (lexicographic order) + material id.
  void create(const std::vector<std::array<Point2, 4>>& coarse_cells, const std::vector<int>& material_ids);
  // append the coarse cells of another (unrefined) triangulation (GridGenerator::merge_triangulations).
  static Triangulation merge(const Triangulation& a, const Triangulation& b);
  void refine_global(int times);
  // refine every active cell whose material id is in `materials` once (with 2:1 smoothing).
  void refine_material(const std::vector<int>& materials);
  // refine every active cell whose centre is closer than r to p (reference Grid.cpp:522-544).
  void refine_near(Point2 p, double r);
  // (re)build the active-cell tables; boundary ids are all 0 afterwards.
  void build_tables();
  MeshTables& tables() { return tab_; }
  const MeshTables& tables() const { return tab_; }
  int n_active_cells() const { return tab_.n_cells; }

private:
  struct Cell {
    int v[4];
    int parent = -1;
    int child0 = -1; // children are child0..child0+3 (contiguous), -1 if active
    int level = 0;
    int material = 0;
  };
  int add_vertex(double x, double y);
  int midpoint(int a, int b);
  void refine_flagged(std::vector<char>& flag);
  void refine_cell(int c);
  std::vector<int> active_order() const;

  std::vector<Point2> verts_;
  std::map<std::pair<double, double>, int> vert_index_;
  std::vector<Cell> cells_;
  std::vector<std::vector<int>> by_level_; // cell ids per level in creation order
  MeshTables tab_;
};

} // namespace pecs
