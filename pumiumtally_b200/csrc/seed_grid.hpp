// Geometry of the relocation seed grid (see SeedGrid in walk_core.cuh).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

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

// Rank of every grid cell along the Morton (Z-order) curve through the cells that exist: the
// binning pass sorts particles by this rank, so a run of consecutive particles covers a compact
// 3-D block of the mesh instead of a one-cell-thick slab (smaller L2 working set per window).
inline std::vector<int32_t> morton_cell_ranks(const SeedGrid &g) {
  auto spread = [](uint64_t v) {  // 21 bits -> every third bit
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
  };
  const int64_t n = int64_t(g.nx) * g.ny * g.nz;
  std::vector<std::pair<uint64_t, int32_t>> key(static_cast<size_t>(n));
  for (int cz = 0; cz < g.nz; ++cz)
    for (int cy = 0; cy < g.ny; ++cy)
      for (int cx = 0; cx < g.nx; ++cx) {
        const int32_t c = (cz * g.ny + cy) * g.nx + cx;
        key[c] = {spread(cx) | spread(cy) << 1 | spread(cz) << 2, c};
      }
  std::sort(key.begin(), key.end());
  std::vector<int32_t> rank(static_cast<size_t>(n));
  for (int64_t r = 0; r < n; ++r) rank[key[r].second] = int32_t(r);
  return rank;
}

}  // namespace ptb
