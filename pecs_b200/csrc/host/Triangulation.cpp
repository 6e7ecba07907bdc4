// Triangulation.cpp -- see Triangulation.hpp.  Own quadtree forest; no deal.II.
#include "Triangulation.hpp"

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <unordered_map>

namespace pecs {

namespace {
// local vertex pairs of the four faces, in the direction of the face's reference coordinate
// (faces 0,1 run in eta, faces 2,3 run in xi) -- deal.II GeometryInfo<2> convention.
const int kFaceVerts[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};

inline std::uint64_t edge_key(int a, int b) {
  const std::uint64_t lo = (std::uint64_t)std::min(a, b), hi = (std::uint64_t)std::max(a, b);
  return (hi << 32) | lo;
}
inline double dist(const double* p, const double* q) { return std::hypot(p[0] - q[0], p[1] - q[1]); }
} // namespace

Point2 MeshTables::center(int c) const {
  const double* v = vtx(c);
  return {0.25 * (v[0] + v[2] + v[4] + v[6]), 0.25 * (v[1] + v[3] + v[5] + v[7])};
}
Point2 MeshTables::face_center(int c, int f) const {
  const double* v = vtx(c);
  const int a = kFaceVerts[f][0], b = kFaceVerts[f][1];
  return {0.5 * (v[2 * a] + v[2 * b]), 0.5 * (v[2 * a + 1] + v[2 * b + 1])};
}
double MeshTables::diameter(int c) const {
  const double* v = vtx(c);
  return std::max(dist(v + 0, v + 6), dist(v + 2, v + 4));
}

int Triangulation::add_vertex(double x, double y) {
  auto key = std::make_pair(x, y);
  auto it = vert_index_.find(key);
  if (it != vert_index_.end()) return it->second;
  const int id = (int)verts_.size();
  verts_.push_back({x, y});
  vert_index_.emplace(key, id);
  return id;
}

int Triangulation::midpoint(int a, int b) {
  return add_vertex(0.5 * (verts_[a].x + verts_[b].x), 0.5 * (verts_[a].y + verts_[b].y));
}

void Triangulation::create(const std::vector<std::array<Point2, 4>>& coarse_cells,
                           const std::vector<int>& material_ids) {
  verts_.clear();
  vert_index_.clear();
  cells_.clear();
  by_level_.assign(1, {});
  for (size_t i = 0; i < coarse_cells.size(); ++i) {
    Cell c;
    for (int k = 0; k < 4; ++k) c.v[k] = add_vertex(coarse_cells[i][k].x, coarse_cells[i][k].y);
    c.material = material_ids[i];
    by_level_[0].push_back((int)cells_.size());
    cells_.push_back(c);
  }
  build_tables();
}

Triangulation Triangulation::merge(const Triangulation& a, const Triangulation& b) {
  std::vector<std::array<Point2, 4>> cc;
  std::vector<int> mats;
  for (const Triangulation* t : {&a, &b}) {
    if (t->by_level_.size() != 1) throw std::runtime_error("merge: triangulations must be unrefined");
    for (int id : t->by_level_[0]) {
      std::array<Point2, 4> p;
      for (int k = 0; k < 4; ++k) p[k] = t->verts_[t->cells_[id].v[k]];
      cc.push_back(p);
      mats.push_back(t->cells_[id].material);
    }
  }
  Triangulation out;
  out.create(cc, mats);
  return out;
}

std::vector<int> Triangulation::active_order() const {
  std::vector<int> act;
  for (const auto& lvl : by_level_)
    for (int id : lvl)
      if (cells_[id].child0 < 0) act.push_back(id);
  return act;
}

void Triangulation::refine_cell(int c) {
  // NOTE: take a copy, cells_ may reallocate below
  const Cell par = cells_[c];
  const int v0 = par.v[0], v1 = par.v[1], v2 = par.v[2], v3 = par.v[3];
  const int m01 = midpoint(v0, v1), m23 = midpoint(v2, v3), m02 = midpoint(v0, v2), m13 = midpoint(v1, v3);
  const int ctr = add_vertex(0.25 * (verts_[v0].x + verts_[v1].x + verts_[v2].x + verts_[v3].x),
                             0.25 * (verts_[v0].y + verts_[v1].y + verts_[v2].y + verts_[v3].y));
  const int cv[4][4] = {{v0, m01, m02, ctr}, {m01, v1, ctr, m13}, {m02, ctr, v2, m23}, {ctr, m13, m23, v3}};
  const int first = (int)cells_.size();
  if ((int)by_level_.size() <= par.level + 1) by_level_.resize(par.level + 2);
  for (int k = 0; k < 4; ++k) {
    Cell ch;
    for (int j = 0; j < 4; ++j) ch.v[j] = cv[k][j];
    ch.parent = c;
    ch.level = par.level + 1;
    ch.material = par.material;
    by_level_[ch.level].push_back((int)cells_.size());
    cells_.push_back(ch);
  }
  cells_[c].child0 = first;
}

void Triangulation::refine_flagged(std::vector<char>& flag) {
  // flag is indexed by ACTIVE cell number of the current tables.  Enforce 2:1 balance across faces.
  bool changed = true;
  while (changed) {
    changed = false;
    for (int c = 0; c < tab_.n_cells; ++c) {
      if (!flag[c]) continue;
      for (int f = 0; f < 4; ++f)
        if (tab_.face_kind[4 * c + f] == FACE_COARSER) {
          const int nb = tab_.neighbor[4 * c + f];
          if (!flag[nb]) {
            flag[nb] = 1;
            changed = true;
          }
        }
    }
  }
  const std::vector<int> act = active_order();
  for (int c = 0; c < (int)act.size(); ++c)
    if (flag[c]) refine_cell(act[c]);
  build_tables();
}

void Triangulation::refine_global(int times) {
  for (int t = 0; t < times; ++t) {
    std::vector<char> flag(tab_.n_cells, 1);
    refine_flagged(flag);
  }
}

void Triangulation::refine_material(const std::vector<int>& materials) {
  std::vector<char> flag(tab_.n_cells, 0);
  for (int c = 0; c < tab_.n_cells; ++c)
    flag[c] = std::find(materials.begin(), materials.end(), tab_.material_id[c]) != materials.end();
  refine_flagged(flag);
}

void Triangulation::refine_near(Point2 p, double r) {
  std::vector<char> flag(tab_.n_cells, 0);
  for (int c = 0; c < tab_.n_cells; ++c) {
    const Point2 ctr = tab_.center(c);
    flag[c] = std::hypot(ctr.x - p.x, ctr.y - p.y) < r;
  }
  refine_flagged(flag);
}

void Triangulation::build_tables() {
  const std::vector<int> act = active_order();
  const int n = (int)act.size();
  MeshTables t;
  t.n_cells = n;
  t.vertices.resize(8 * (size_t)n);
  t.material_id.resize(n);
  t.level.resize(n);
  t.face_kind.assign(4 * (size_t)n, -1);
  t.neighbor.assign(4 * (size_t)n, -1);
  t.neighbor2.assign(4 * (size_t)n, -1);
  t.boundary_id.assign(4 * (size_t)n, -1);
  t.nb_parent_diameter.assign(4 * (size_t)n, 0.0);

  struct Slot {
    int cell[2] = {-1, -1};
    int face[2] = {-1, -1};
    int n = 0;
  };
  std::unordered_map<std::uint64_t, Slot> edges;
  edges.reserve(4 * (size_t)n);
  for (int c = 0; c < n; ++c) {
    const Cell& cell = cells_[act[c]];
    for (int k = 0; k < 4; ++k) {
      t.vertices[8 * (size_t)c + 2 * k] = verts_[cell.v[k]].x;
      t.vertices[8 * (size_t)c + 2 * k + 1] = verts_[cell.v[k]].y;
    }
    t.material_id[c] = cell.material;
    t.level[c] = cell.level;
    for (int f = 0; f < 4; ++f) {
      Slot& s = edges[edge_key(cell.v[kFaceVerts[f][0]], cell.v[kFaceVerts[f][1]])];
      if (s.n >= 2) throw std::runtime_error("build_tables: edge shared by more than two active cells");
      s.cell[s.n] = c;
      s.face[s.n] = f;
      ++s.n;
    }
  }
  // pass 1: same-level faces and hanging faces seen from the coarse side
  for (int c = 0; c < n; ++c) {
    const Cell& cell = cells_[act[c]];
    for (int f = 0; f < 4; ++f) {
      const int a = cell.v[kFaceVerts[f][0]], b = cell.v[kFaceVerts[f][1]];
      const Slot& s = edges[edge_key(a, b)];
      if (s.n == 2) {
        const int other = (s.cell[0] == c && s.face[0] == f) ? 1 : 0;
        t.face_kind[4 * c + f] = FACE_SAME_LEVEL;
        t.neighbor[4 * c + f] = s.cell[other];
        continue;
      }
      auto vm = vert_index_.find(std::make_pair(0.5 * (verts_[a].x + verts_[b].x), 0.5 * (verts_[a].y + verts_[b].y)));
      if (vm == vert_index_.end()) continue;
      auto e0 = edges.find(edge_key(a, vm->second)), e1 = edges.find(edge_key(vm->second, b));
      if (e0 == edges.end() || e1 == edges.end()) continue;
      if (e0->second.n != 1 || e1->second.n != 1) throw std::runtime_error("build_tables: inconsistent hanging face");
      const int n0 = e0->second.cell[0], n1 = e1->second.cell[0];
      t.face_kind[4 * c + f] = FACE_HAS_CHILDREN;
      t.neighbor[4 * c + f] = n0;
      t.neighbor2[4 * c + f] = n1;
      // diameter of the refined neighbour (the parent of the two fine cells)
      const Cell& par = cells_[cells_[act[n0]].parent];
      const Point2 p0 = verts_[par.v[0]], p1 = verts_[par.v[1]], p2 = verts_[par.v[2]], p3 = verts_[par.v[3]];
      t.nb_parent_diameter[4 * c + f] =
          std::max(std::hypot(p3.x - p0.x, p3.y - p0.y), std::hypot(p2.x - p1.x, p2.y - p1.y));
      const int nf0 = e0->second.face[0], nf1 = e1->second.face[0];
      t.face_kind[4 * n0 + nf0] = FACE_COARSER;
      t.neighbor[4 * n0 + nf0] = c;
      t.neighbor2[4 * n0 + nf0] = 0;
      t.face_kind[4 * n1 + nf1] = FACE_COARSER;
      t.neighbor[4 * n1 + nf1] = c;
      t.neighbor2[4 * n1 + nf1] = 1;
    }
  }
  // pass 2: everything still unset is boundary (default boundary id 0, like deal.II)
  for (size_t i = 0; i < t.face_kind.size(); ++i)
    if (t.face_kind[i] < 0) {
      t.face_kind[i] = FACE_BOUNDARY;
      t.boundary_id[i] = 0;
    }
  tab_ = std::move(t);
}

} // namespace pecs
