// Geometry of the relocation seed grid (see SeedGrid in walk_core.cuh).
#pragma once
#include <algorithm>
#include <cmath>

#include "tet_mesh.hpp"
#include "walk_core.cuh"

namespace ptb {

// Roughly one grid cell per `tets_per_cell` tets, cubic cells, at most 2^24 cells.
// cell_tet is left null; the caller fills the table by localising the seed points.
inline SeedGrid choose_seed_grid(const HostMesh &m, double tets_per_cell = 4.0) {
  SeedGrid g{};
  const double lx = m.bbox_hi[0] - m.bbox_lo[0], ly = m.bbox_hi[1] - m.bbox_lo[1],
               lz = m.bbox_hi[2] - m.bbox_lo[2];
  const double vol = std::max(lx * ly * lz, 1e-300);
  double cells = std::min(std::max(double(m.ntets) / tets_per_cell, 1.0), double(1 << 24));
  double h = std::cbrt(vol / cells);
  auto dim = [&](double l) { return int32_t(std::min(std::max(std::ceil(l / h), 1.0), 1024.0)); };
  g.nx = dim(lx); g.ny = dim(ly); g.nz = dim(lz);
  // grow h if the 1024 clamp cut an axis short, so the grid still covers the box
  h = std::max({h, lx / g.nx, ly / g.ny, lz / g.nz});
  g.h = h;
  g.inv_h = 1.0 / h;
  g.x0 = m.bbox_lo[0]; g.y0 = m.bbox_lo[1]; g.z0 = m.bbox_lo[2];
  g.far2 = 4.0 * h * h;  // seed when the target is more than two cells away
  g.cell_tet = nullptr;
  return g;
}

}  // namespace ptb
