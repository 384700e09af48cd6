// TEST-ONLY host build of the device walk logic.
//
// Compiles pumiumtally_b200/csrc/walk_core.cuh (the per-ray state machine the
// CUDA kernels run) and tet_mesh.cpp (the record packer) with g++ so that the
// arithmetic can be checked against the oracle on a machine without a GPU.
// It is never linked into libpumitally.so and is not a fallback: the product
// refuses to construct an engine without a CUDA device.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "seed_grid.hpp"
#include "tet_mesh.hpp"
#include "walk_compact.cuh"
#include "walk_core.cuh"

using namespace ptb;

namespace {
struct Emul {
  HostMesh mesh;
  std::vector<TetRecord> recs;
  std::vector<double> flux;
  std::vector<ParticleState> state;
  DeviceStats stats{};
  int n = 0;
  SeedGrid grid{};
  std::vector<int32_t> cell_tet;
  unsigned long long degenerate_rays = 0;
  bool edge = false;  // compact layout + edge-function exit test (walk_compact.cuh)
  // same construction as Engine::build_seed_grid(): localise every seed point from the
  // centroid of element 0, keep the tet where the point was reached
  void build_grid() {
    grid = choose_seed_grid(mesh);
    const int nc = grid.nx * grid.ny * grid.nz;
    std::vector<double> xyz(3 * size_t(nc));
    std::vector<ParticleState> ts(nc, ParticleState{mesh.centroid0[0], mesh.centroid0[1], mesh.centroid0[2], mesh.start_elem, 0});
    for (int i = 0; i < nc; ++i)
      seed_point(grid, i % grid.nx, (i / grid.nx) % grid.ny, i / (grid.nx * grid.ny), xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    DeviceStats scratch{};
    walk(ts.data(), nc, xyz.data(), nullptr, nullptr, nullptr, &scratch, nullptr);
    cell_tet.resize(nc);
    for (int i = 0; i < nc; ++i)
      cell_tet[i] = (ts[i].x == xyz[3 * i] && ts[i].y == xyz[3 * i + 1] && ts[i].z == xyz[3 * i + 2]) ? ts[i].elem : -1;
    grid.cell_tet = cell_tet.data();
  }
  void run(const double *origin, const double *dest, const int8_t *flying, const double *weights) {
    walk(state.data(), n, origin, dest, flying, weights, &stats,
         grid.cell_tet ? &grid : nullptr);
  }
  void walk(ParticleState *qs, int count, const double *origin,
            const double *dest, const int8_t *flying, const double *weights, DeviceStats *st,
            const SeedGrid *g) {
    WalkParams P{};
    P.tets = recs.data();
    P.flux = flux.data();
    P.state = qs;
    P.origin = origin; P.dest = dest; P.flying = flying; P.weights = weights;
    P.begin = 0; P.end = count;
    P.max_iters = int32_t(mesh.ntets + 16);
    P.stats = st;
    P.cx = mesh.center[0]; P.cy = mesh.center[1]; P.cz = mesh.center[2];
    if (g) P.grid = *g;
    DeviceStats &stats = *st;
    const int n = count;
    const bool trace = std::getenv("PTB_EMUL_TRACE") != nullptr;
    for (int i = 0; i < n; ++i) {
      Counters c;
      Ray r;
      begin_particle(P, i, r, c, true);
      EdgeRay g{};
      bool planes = !edge;  // per ray: a degenerate ray finishes on the plane records
      while (r.stage != kStageDone) {
        if (!planes) {
          // mirror of the compact kernel: first tet of a ray reads the TetStart line, every
          // later one the TetLinks sector + the one new vertex
          double texit;
          int32_t next, roles;
          const TetStart &S = mesh.starts[r.e];
          bool ok;
          if (r.entry < 0) {
            ok = edge_first(r, g, S.links, S.v, texit, next, roles);
          } else {
            const VertexRec &D = mesh.cverts[g.dv];
            ok = edge_step(r, g, S.links, D.x, D.y, D.z, texit, next, roles);
          }
          if (!ok) {  // coplanar edge: redo this tet, and the rest of the ray, with the planes
            planes = true;
            r.entry = -1;
            ++degenerate_rays;
          } else {
            advance(P, i, r, texit, next, roles, c, true);
          }
          continue;
        }
        // mirror of the persistent kernel's fetch: the entry face's sector is never read
        const double *rec = recs[r.e].d;
        const int en = r.entry;
        ExitScan sc;
        for (int k = 0; k < (en < 0 ? 4 : 3); ++k) {
          const int fk = k + ((en >= 0 && k >= en) ? 1 : 0);
          int32_t nb, bk;
          face_payload(rec[4 * fk], rec[4 * fk + 1], rec[4 * fk + 2], rec[4 * fk + 3], r.e, fk, nb, bk);
          scan_face(sc, rec[4 * fk], rec[4 * fk + 1], rec[4 * fk + 2], rec[4 * fk + 3], nb, bk, r.ox, r.oy,
                    r.oz, r.ux, r.uy, r.uz);
        }
        if (trace) fprintf(stderr, "  p%d tet %d entry %d stage %d tcur %.17g texit %.17g (num %.17g den %.17g) next %d\n", i, r.e, en,
                           r.stage, r.tcur, exit_parameter(sc), sc.bnum, sc.bden, sc.nbr);
        advance(P, i, r, exit_parameter(sc), sc.nbr, sc.back, c, true);
        if (r.iters == 0) planes = !edge;  // a new ray (phase 2 after phase 1) starts on the fast path again
      }
      stats.segments += c.segs; stats.tracks += c.tracks;
      stats.relocations += c.relocs; stats.lost += c.lost;
    }
  }
};
}  // namespace

extern "C" {
void *ptb_emul_create(const double *coords, int64_t nverts, const int32_t *t2v, int64_t ntets, int n) {
  auto *e = new Emul;
  std::string err;
  if (!e->mesh.from_arrays(coords, nverts, t2v, ntets, &err)) { delete e; return nullptr; }
  e->recs = e->mesh.records;
  e->n = n;
  e->flux.assign(size_t(ntets), 0.0);
  e->state.assign(size_t(n), ParticleState{e->mesh.centroid0[0], e->mesh.centroid0[1], e->mesh.centroid0[2], e->mesh.start_elem, 0});
  return e;
}
void *ptb_emul_create_spec(const char *spec, int n) {
  auto *e = new Emul;
  std::string err;
  if (!e->mesh.load(spec, &err)) { delete e; return nullptr; }
  e->recs = e->mesh.records;
  e->n = n;
  e->flux.assign(size_t(e->mesh.ntets), 0.0);
  e->state.assign(size_t(n), ParticleState{e->mesh.centroid0[0], e->mesh.centroid0[1], e->mesh.centroid0[2], e->mesh.start_elem, 0});
  return e;
}
int ptb_emul_build_grid(void *h) {
  auto *e = static_cast<Emul *>(h);
  e->build_grid();
  int valid = 0;
  for (int32_t t : e->cell_tet) valid += t >= 0;
  return valid;
}
int ptb_emul_set_layout(void *h, int edge) {
  auto *e = static_cast<Emul *>(h);
  std::string err;
  if (edge && e->mesh.starts.empty() && !e->mesh.build_compact(&err)) return 1;
  e->edge = edge != 0;
  return 0;
}
unsigned long long ptb_emul_degenerate_rays(void *h) { return static_cast<Emul *>(h)->degenerate_rays; }
void ptb_emul_grid_dims(void *h, int32_t *out) {
  auto *e = static_cast<Emul *>(h);
  out[0] = e->grid.nx; out[1] = e->grid.ny; out[2] = e->grid.nz;
}
int ptb_emul_hull_convex(void *h) { return static_cast<Emul *>(h)->mesh.hull_convex ? 1 : 0; }

void ptb_emul_sizes(void *h, int64_t *out) {
  auto *e = static_cast<Emul *>(h);
  out[0] = e->mesh.nverts; out[1] = e->mesh.ntets;
}
void ptb_emul_mesh(void *h, double *coords, int32_t *t2v, double *vol) {
  auto *e = static_cast<Emul *>(h);
  std::memcpy(coords, e->mesh.coords.data(), e->mesh.coords.size() * 8);
  const auto t2v_o = e->mesh.to_original(e->mesh.t2v.data(), 4);
  const auto vol_o = e->mesh.to_original(e->mesh.volume.data());
  std::memcpy(t2v, t2v_o.data(), t2v_o.size() * 4);
  std::memcpy(vol, vol_o.data(), vol_o.size() * 8);
}
void ptb_emul_destroy(void *h) { delete static_cast<Emul *>(h); }
void ptb_emul_localize(void *h, const double *xyz) { static_cast<Emul *>(h)->run(xyz, nullptr, nullptr, nullptr); }
void ptb_emul_move(void *h, const double *origin, const double *dest, int8_t *flying, const double *w) {
  auto *e = static_cast<Emul *>(h);
  e->run(origin, dest, flying, w);
  std::memset(flying, 0, size_t(e->n));
}
void ptb_emul_get(void *h, double *flux, int32_t *elem, double *pos, unsigned long long *stats, int32_t *adj) {
  auto *e = static_cast<Emul *>(h);
  if (flux) {
    const auto f = e->mesh.to_original(e->flux.data());
    std::memcpy(flux, f.data(), f.size() * 8);
  }
  if (elem)
    for (int i = 0; i < e->n; ++i) elem[i] = e->mesh.orig_of_internal[e->state[i].elem];
  if (pos)
    for (int i = 0; i < e->n; ++i) { pos[3 * i] = e->state[i].x; pos[3 * i + 1] = e->state[i].y; pos[3 * i + 2] = e->state[i].z; }
  if (stats) { stats[0] = e->stats.segments; stats[1] = e->stats.tracks; stats[2] = e->stats.relocations; stats[3] = e->stats.lost; }
  if (adj) {
    const auto a = e->mesh.adjacency_original();
    std::memcpy(adj, a.data(), a.size() * 4);
  }
}
}
