"""Shared parity-test plumbing: run the same seeded batches through an engine
under test and through the oracle, then compare with the tolerances of
SURVEY.md section 8c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from oracle.oracle import OraclePumiTally
from pumiumtally_b200.mesh import kuhn_box
from pumiumtally_b200.workload import SyntheticWorkload

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# flux parity (north_star: 1e-6 relative, element for element; the absolute
# floor only guards empty / nearly empty tets)
FLUX_RTOL = 1e-6
FLUX_ATOL_FRACTION = 1e-12
# positions: a reached destination is stored exactly; a point clipped at the hull is computed from
# the 44-bit-mantissa face plane (tet_mesh.hpp), i.e. ~1e-13 relative, amplified at grazing angles
POS_RTOL = 1e-10


def assert_flux_close(got, want, label=""):
    got, want = np.asarray(got), np.asarray(want)
    tol = FLUX_RTOL * np.abs(want) + FLUX_ATOL_FRACTION * np.abs(want).sum()
    bad = np.abs(got - want) > tol
    assert not bad.any(), (
        f"{label}: {int(bad.sum())} of {bad.size} elements differ; worst rel "
        f"{np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)):.3e}")


def assert_positions_close(got, want, label=""):
    scale = np.maximum(np.abs(want).max(), 1.0)
    err = np.abs(np.asarray(got) - np.asarray(want)).max() if len(want) else 0.0
    assert err <= POS_RTOL * scale, f"{label}: position error {err:.3e}"


def run_workload(engine, oracle, workload, steps, check_each_step=True, label=""):
    """Drive engine and oracle with identical batches; assert parity after each move."""
    init = workload.initial_positions()
    engine.CopyInitialPosition(init.reshape(-1).copy())
    oracle.CopyInitialPosition(init.reshape(-1).copy())
    np.testing.assert_array_equal(engine.elem_ids, oracle.elem_ids, err_msg=f"{label} parent elems after localisation")
    assert not engine.flux.any(), "flux must stay zero during localisation"
    for s in range(steps):
        o, d, f, w = workload.next_step()
        f1, f2 = f.copy(), f.copy()
        engine.MoveToNextLocation(o.reshape(-1).copy(), d.reshape(-1).copy(), f1, w.copy())
        oracle.MoveToNextLocation(o.reshape(-1).copy(), d.reshape(-1).copy(), f2, w.copy())
        assert not f1.any() and not f2.any(), "flying[] must be zeroed on return"
        if check_each_step or s == steps - 1:
            assert_flux_close(engine.flux, oracle.flux, f"{label} step {s}")
            np.testing.assert_array_equal(engine.elem_ids, oracle.elem_ids, err_msg=f"{label} elems step {s}")
            assert_positions_close(engine.positions, oracle.positions, f"{label} step {s}")


def oracle_binned_move(oracle, flux_bins, o, d, f, w, bins):
    """A binned move on the (unfiltered) oracle: particles are independent and non-flying particles are not
    touched, so one oracle move per bin with the flying flags masked to that bin is the filtered tally; the
    flux each of them adds goes to flux_bins[bin].  Bins outside [0, nbins) fly unscored."""
    nb = flux_bins.shape[0]
    inside = (bins >= 0) & (bins < nb)
    for b in list(range(nb)) + [nb]:
        mask = ((f == 1) & ((bins == b) if b < nb else ~inside)).astype(np.int8)
        if not mask.any():
            continue
        before = oracle.flux.copy()
        oracle.MoveToNextLocation(o.reshape(-1).copy(), d.reshape(-1).copy(), mask, w.copy())
        if b < nb:
            flux_bins[b] += oracle.flux - before


# ---------------------------------------------------------------------------
# test-only host build of the device walk logic (tests/host_emul/emul_walk.cpp)
# ---------------------------------------------------------------------------
_EMUL = {}


def emul_lib(fma=None):
    """fma=True: let g++ contract a*b+c into fused multiply-adds, as nvcc does for the device code, so
    that rounding-sensitive (degenerate-input) tests see both flavours of the arithmetic.  Default:
    the environment variable PTB_EMUL_FMA=1, else off."""
    if fma is None:
        fma = os.environ.get("PTB_EMUL_FMA") == "1"
    if fma not in _EMUL:
        so = os.path.join(HERE, "host_emul", "libptb_emul_fma.so" if fma else "libptb_emul.so")
        srcs = [os.path.join(HERE, "host_emul", "emul_walk.cpp"),
                os.path.join(ROOT, "pumiumtally_b200", "csrc", "tet_mesh.cpp"),
                os.path.join(ROOT, "pumiumtally_b200", "csrc", "osh_reader.cpp"),
                os.path.join(ROOT, "pumiumtally_b200", "csrc", "gmsh_reader.cpp")]
        deps = srcs + [os.path.join(ROOT, "pumiumtally_b200", "csrc", h) for h in ("walk_core.cuh", "walk_compact.cuh", "tet_mesh.hpp", "seed_grid.hpp")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp",
                                   *(["-mfma", "-ffp-contract=fast"] if fma else ["-ffp-contract=off"]),
                                   "-I", os.path.join(ROOT, "pumiumtally_b200", "csrc"),
                                   "-I", os.path.join(ROOT, "include"), *srcs, "-o", so, "-lz"])
        L = C.CDLL(so)
        L.ptb_emul_create.restype = C.c_void_p
        L.ptb_emul_create.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]
        L.ptb_emul_create_spec.restype = C.c_void_p
        L.ptb_emul_create_spec.argtypes = [C.c_char_p, C.c_int]
        L.ptb_emul_destroy.argtypes = [C.c_void_p]
        L.ptb_emul_localize.argtypes = [C.c_void_p, C.c_void_p]
        L.ptb_emul_move.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ptb_emul_get.argtypes = [C.c_void_p] * 6
        L.ptb_emul_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.ptb_emul_build_grid.argtypes = [C.c_void_p]
        L.ptb_emul_grid_dims.argtypes = [C.c_void_p, C.c_void_p]
        L.ptb_emul_set_layout.argtypes = [C.c_void_p, C.c_int]
        L.ptb_emul_degenerate_rays.restype = C.c_ulonglong
        L.ptb_emul_degenerate_rays.argtypes = [C.c_void_p]
        L.ptb_emul_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ptb_emul_hull_convex.argtypes = [C.c_void_p]
        _EMUL[fma] = L
    return _EMUL[fma]


class HostEmulTally:
    """The CUDA kernels' per-ray state machine compiled for the host (tests only)."""

    def __init__(self, coords=None, tet2vert=None, num_particles=0, spec=None, seed_grid=False, layout="planes",
                 fma=None):
        self._L = emul_lib(fma)
        self._want_grid = seed_grid
        self.num_particles = int(num_particles)
        if spec is not None:
            self._h = self._L.ptb_emul_create_spec(spec.encode(), self.num_particles)
        else:
            coords = np.ascontiguousarray(coords, dtype=np.float64)
            t2v = np.ascontiguousarray(tet2vert, dtype=np.int32)
            self._h = self._L.ptb_emul_create(coords.ctypes.data, coords.shape[0], t2v.ctypes.data, t2v.shape[0],
                                              self.num_particles)
        if not self._h:
            raise RuntimeError("mesh rejected")
        sz = np.zeros(2, dtype=np.int64)
        self._L.ptb_emul_sizes(self._h, sz.ctypes.data)
        self.num_verts, self.num_elements = int(sz[0]), int(sz[1])
        if layout == "edge" and self._L.ptb_emul_set_layout(self._h, 1) != 0:
            raise RuntimeError("compact layout rejected")
        self.grid_valid_cells = self._L.ptb_emul_build_grid(self._h) if seed_grid else 0

    @property
    def hull_convex(self):
        return bool(self._L.ptb_emul_hull_convex(self._h))

    def grid_dims(self):
        d = np.zeros(3, dtype=np.int32)
        self._L.ptb_emul_grid_dims(self._h, d.ctypes.data)
        return tuple(int(x) for x in d)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ptb_emul_destroy(self._h)

    def CopyInitialPosition(self, xyz, size=None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self._L.ptb_emul_localize(self._h, xyz.ctypes.data)

    def MoveToNextLocation(self, o, d, flying, w, size=None):
        o, d, w = (np.ascontiguousarray(a, dtype=np.float64) for a in (o, d, w))
        self._L.ptb_emul_move(self._h, o.ctypes.data, d.ctypes.data, flying.ctypes.data, w.ctypes.data)

    def _get(self):
        flux = np.empty(self.num_elements)
        elem = np.empty(self.num_particles, dtype=np.int32)
        pos = np.empty((self.num_particles, 3))
        stats = np.zeros(4, dtype=np.uint64)
        adj = np.empty((self.num_elements, 4), dtype=np.int32)
        self._L.ptb_emul_get(self._h, flux.ctypes.data, elem.ctypes.data, pos.ctypes.data, stats.ctypes.data,
                             adj.ctypes.data)
        return flux, elem, pos, stats, adj

    flux = property(lambda s: s._get()[0])
    elem_ids = property(lambda s: s._get()[1])
    positions = property(lambda s: s._get()[2])
    adjacency = property(lambda s: s._get()[4])

    @property
    def degenerate_rays(self):
        """Rays the edge-function walk handed to the plane records (layout="edge")."""
        return int(self._L.ptb_emul_degenerate_rays(self._h))

    def stats(self):
        st = self._get()[3]
        return dict(segments=int(st[0]), tracks=int(st[1]), relocations=int(st[2]), lost=int(st[3]))

    def mesh_arrays(self):
        c = np.empty((self.num_verts, 3))
        t = np.empty((self.num_elements, 4), dtype=np.int32)
        v = np.empty(self.num_elements)
        self._L.ptb_emul_mesh(self._h, c.ctypes.data, t.ctypes.data, v.ctypes.data)
        return c, t, v


def carve(coords, t2v, keep):
    """Sub-mesh of the tets selected by the boolean mask `keep`, vertices renumbered."""
    t = t2v[keep]
    verts, inv = np.unique(t.ravel(), return_inverse=True)
    return np.ascontiguousarray(coords[verts]), inv.reshape(-1, 4).astype(np.int32)


def l_shaped_mesh(nx=4, ny=4, nz=2):
    """Kuhn box with one quadrant (x > nx/2 and y > ny/2) removed: an L-shaped prism, hull not convex."""
    coords, t2v = kuhn_box(nx, ny, nz)
    c = coords[t2v].mean(1)
    return carve(coords, t2v, ~((c[:, 0] > nx / 2) & (c[:, 1] > ny / 2)))


def non_convex_relocation_scenario(make_engine, expect_reference=True):
    """Particles re-sourced across the notch of an L-shaped mesh.  The reference walks straight from the
    old position to the new one and stops at the hull when that segment leaves the mesh, the flight then
    starts from where the relocation stopped (PumiTallyImpl.cpp:71-145) -- so must this engine, unless the
    seed-grid shortcut is forced on, in which case the new position is reached."""
    coords, t2v = l_shaped_mesh(16, 16, 4)  # fine enough for the far-relocation threshold of the seed grid
    n = 400
    rng = np.random.default_rng(2)
    a = np.column_stack([rng.uniform(9.2, 15.8, n), rng.uniform(0.2, 6.8, n), rng.uniform(0.2, 3.8, n)])  # one arm
    b = np.column_stack([rng.uniform(0.2, 6.8, n), rng.uniform(9.2, 15.8, n), rng.uniform(0.2, 3.8, n)])  # the other arm
    d = b + rng.normal(0, 0.2, (n, 3))
    d[:, 2] = np.clip(d[:, 2], 0.05, 3.95)
    w = rng.uniform(0.5, 1.0, n)
    eng, orc = make_engine(coords, t2v, n), OraclePumiTally(coords, t2v, n)
    for e in (eng, orc):
        e.CopyInitialPosition(a.reshape(-1).copy())
        e.MoveToNextLocation(b.reshape(-1).copy(), d.reshape(-1).copy(), np.ones(n, dtype=np.int8), w.copy())
    crosses_notch = (orc.positions != d).any(1)  # the oracle (= the reference's walk) did not get there
    assert crosses_notch.sum() > n // 4
    if expect_reference:
        assert_flux_close(eng.flux, orc.flux, "non-convex relocation")
        np.testing.assert_array_equal(eng.elem_ids, orc.elem_ids)
        np.testing.assert_allclose(eng.positions, orc.positions, rtol=0, atol=1e-12)
    else:
        reached = (eng.positions == d).all(1)
        assert reached[crosses_notch].mean() > 0.9  # the shortcut gets (nearly) everybody to the new position
    return eng, orc


def box_case(cells, n, **kw):
    coords, t2v = kuhn_box(*cells)
    wl = SyntheticWorkload(box=tuple(float(c) for c in cells), num_particles=n, **kw)
    return coords, t2v, wl


def edge_case_scenario(make_engine):
    """Clipping at the hull, zero-length and non-flying particles, relocation, zero weight."""
    coords, t2v = kuhn_box(2, 2, 2)
    n = 6
    eng, orc = make_engine(coords, t2v, n), OraclePumiTally(coords, t2v, n)
    init = np.array([[0.3, 0.2, 0.1], [1.7, 1.2, 0.4], [0.5, 1.5, 1.9], [1.1, 0.9, 0.3], [0.2, 0.3, 1.7], [1.9, 1.8, 1.7]])
    for e in (eng, orc):
        e.CopyInitialPosition(init.reshape(-1).copy())
    np.testing.assert_array_equal(eng.elem_ids, orc.elem_ids)
    origin = init.copy()
    dest = init.copy()
    dest[0] = [5.0, 0.2, 0.1]       # leaves through +x: clipped at x = 2
    dest[1] = init[1]               # zero-length flight
    dest[2] = [0.5, 1.5, -3.0]      # leaves through z = 0
    origin[3] = [0.4, 1.6, 1.2]     # re-sampled: relocation without tally, then flight
    dest[3] = [0.6, 1.1, 0.2]
    dest[4] = [1.3, 1.4, 0.2]       # not flying: must not move or tally
    dest[5] = [1.2, 1.1, 1.3]
    fly = np.array([1, 1, 1, 1, 0, 1], dtype=np.int8)
    w = np.array([1.0, 2.0, 0.5, 0.25, 9.0, 0.0])  # zero weight still moves
    for e in (eng, orc):
        e.MoveToNextLocation(origin.reshape(-1).copy(), dest.reshape(-1).copy(), fly.copy(), w.copy())
    assert_flux_close(eng.flux, orc.flux, "edge")
    np.testing.assert_array_equal(eng.elem_ids, orc.elem_ids)
    np.testing.assert_allclose(eng.positions, orc.positions, atol=1e-13)
    p = eng.positions
    np.testing.assert_allclose(p[0], [2.0, 0.2, 0.1], atol=1e-13)
    np.testing.assert_allclose(p[2], [0.5, 1.5, 0.0], atol=1e-13)
    np.testing.assert_array_equal(p[4], init[4])
    # total tally = sum of weighted in-mesh track lengths
    want = 1.0 * 1.7 + 0.5 * 1.9 + 0.25 * np.linalg.norm(dest[3] - origin[3])
    np.testing.assert_allclose(eng.flux.sum(), want, rtol=1e-13)
    # a second move starting on the hull, heading further out: nothing to tally
    o2 = eng.positions.copy()
    d2 = o2.copy()
    d2[0] = [9.0, 0.2, 0.1]
    fly2 = np.array([1, 0, 0, 0, 0, 0], dtype=np.int8)
    before = eng.flux.sum()
    for e in (eng, orc):
        e.MoveToNextLocation(o2.reshape(-1).copy(), d2.reshape(-1).copy(), fly2.copy(), np.ones(n))
    np.testing.assert_allclose(eng.flux.sum(), before, atol=1e-13)
    np.testing.assert_allclose(eng.positions[0], [2.0, 0.2, 0.1], atol=1e-13)
    np.testing.assert_array_equal(eng.elem_ids, orc.elem_ids)
    return eng


def non_finite_input_scenario(make_engine):
    """NaN / infinity in the caller's arrays must not poison the tally or the stored particle state:
    such particles sit the move out (counted as lost), everything else is unaffected."""
    coords, t2v, wl = box_case((6, 6, 5), 4000)
    n = wl.n
    eng, orc = make_engine(coords, t2v, n), OraclePumiTally(coords, t2v, n)
    init = wl.initial_positions()
    for e in (eng, orc):
        e.CopyInitialPosition(init.reshape(-1).copy())
    o, d, f, w = wl.next_step()
    f[:] = 1
    bad_dest, bad_w, bad_origin = np.arange(0, 40), np.arange(40, 60), np.arange(60, 90)
    d_bad, w_bad, o_bad = d.copy(), w.copy(), o.copy()
    d_bad[bad_dest[:20], 1] = np.nan
    d_bad[bad_dest[20:], 2] = np.inf
    w_bad[bad_w[:10]] = np.nan
    w_bad[bad_w[10:]] = -np.inf
    o_bad[bad_origin[:15], 0] = np.nan
    o_bad[bad_origin[15:], 2] = -np.inf
    skip = np.concatenate([bad_dest, bad_w, bad_origin])
    f_ref = f.copy()
    f_ref[skip] = 0
    eng.MoveToNextLocation(o_bad.reshape(-1).copy(), d_bad.reshape(-1).copy(), f.copy(), w_bad.copy())
    orc.MoveToNextLocation(o.reshape(-1).copy(), d.reshape(-1).copy(), f_ref, w.copy())
    flux = eng.flux
    assert np.isfinite(flux).all() and np.isfinite(eng.positions).all()
    assert_flux_close(flux, orc.flux, "non-finite inputs")
    good = np.ones(n, dtype=bool)
    good[skip] = False
    np.testing.assert_array_equal(eng.elem_ids[good], orc.elem_ids[good])
    np.testing.assert_array_equal(eng.positions[bad_origin], orc.positions[bad_origin])  # untouched
    assert eng.stats()["lost"] == len(skip)
    # the next, clean move works for every particle
    o2, d2, f2, w2 = wl.next_step()
    f2[:] = 1
    o2[skip] = d[skip]  # re-source the affected particles at valid points
    for e in (eng, orc):
        e.MoveToNextLocation(o2.reshape(-1).copy(), d2.reshape(-1).copy(), f2.copy(), w2.copy())
    assert_flux_close(eng.flux, orc.flux, "after non-finite inputs")
    np.testing.assert_array_equal(eng.elem_ids, orc.elem_ids)
    return eng


def lattice_track_scenario(make_engine, seeds, exact_destinations=True):
    """Tracks between points of the quarter-cell lattice of Kuhn boxes: along the hull surface, along
    edges, inside face planes, through vertices.  Attribution to a particular tet is a tie-break there;
    what must hold: nothing lost or stopped early, destinations reached exactly, total tally equal to the
    total track length, final tet containing the final position."""
    for seed in seeds:
        rng = np.random.default_rng(seed)
        dims = tuple(int(x) for x in rng.integers(1, 6, 3))
        lengths = tuple(float(x) for x in rng.choice([0.5, 1.0, 2.0, 3.0], 3)) if seed % 2 else tuple(float(d) for d in dims)
        coords, t2v = kuhn_box(*dims, *lengths)
        h = np.array(lengths) / np.array(dims)
        a = rng.integers(0, 4 * np.array(dims) + 1, size=(400, 3)) * h / 4.0
        b = rng.integers(0, 4 * np.array(dims) + 1, size=(400, 3)) * h / 4.0
        keep = np.abs(a - b).sum(1) > 0
        a, b = a[keep], b[keep]
        n = len(a)
        w = rng.uniform(0.5, 1.0, n)
        eng = make_engine(coords, t2v, n)
        eng.CopyInitialPosition(a.reshape(-1).copy())
        if exact_destinations:
            np.testing.assert_array_equal(eng.positions, a, err_msg=f"seed {seed}: localisation")
        else:
            np.testing.assert_allclose(eng.positions, a, rtol=0, atol=4e-16 * max(lengths), err_msg=f"seed {seed}: localisation")
        eng.MoveToNextLocation(a.reshape(-1).copy(), b.reshape(-1).copy(), np.ones(n, dtype=np.int8), w.copy())
        assert eng.stats()["lost"] == 0
        if exact_destinations:
            np.testing.assert_array_equal(eng.positions, b, err_msg=f"seed {seed}: destinations")
        else:  # the experimental edge-function walk has no outward-erring hull: a destination exactly on
            # the hull may count as a clip point there, which is the same point to within an ulp
            np.testing.assert_allclose(eng.positions, b, rtol=0, atol=4e-16 * max(lengths), err_msg=f"seed {seed}: destinations")
        np.testing.assert_allclose(eng.flux.sum(), (np.linalg.norm(b - a, axis=1) * w).sum(), rtol=1e-12)
        v = coords[t2v[eng.elem_ids]]
        T = np.transpose(v[:, 1:] - v[:, :1], (0, 2, 1))
        lam = np.linalg.solve(T, (eng.positions - v[:, 0])[..., None])[..., 0]
        assert min(lam.min(), (1.0 - lam.sum(1)).min()) > -1e-11, f"seed {seed}"


def unstructured_special_point_scenario(make_engine, seeds):
    """Tracks between vertices, edge points, face centroids and tet centroids of Delaunay / jittered
    meshes: rays through vertices, along edges, inside faces.  Nothing may be lost or stopped early,
    destinations are reached, the total tally equals the total track length."""
    from pumiumtally_b200.mesh import delaunay_box, jitter_interior

    for seed in seeds:
        rng = np.random.default_rng(1000 + seed)
        if seed % 2:
            coords, t2v = delaunay_box(int(rng.integers(20, 200)), seed=seed)
        else:
            dims = tuple(int(x) for x in rng.integers(1, 5, 3))
            coords, t2v = jitter_interior(*kuhn_box(*dims), amplitude=float(rng.uniform(0.05, 0.3)), seed=seed)
        tets = t2v[rng.choice(len(t2v), min(len(t2v), 30), replace=False)]
        pts = []
        for t in tets:
            v = coords[t]
            pts += [v[0], v[3], 0.5 * (v[0] + v[1]), 0.25 * v[1] + 0.75 * v[2], (v[0] + v[1] + v[2]) / 3.0,
                    (v[1] + v[2] + v[3]) / 3.0, v.mean(0)]
        pts = np.array(pts)
        a, b = rng.integers(0, len(pts), 500), rng.integers(0, len(pts), 500)
        keep = a != b
        start, dest = pts[a[keep]], pts[b[keep]]
        n = len(start)
        w = rng.uniform(0.5, 1.0, n)
        eng = make_engine(coords, t2v, n)
        eng.CopyInitialPosition(start.reshape(-1).copy())
        eng.MoveToNextLocation(start.reshape(-1).copy(), dest.reshape(-1).copy(), np.ones(n, dtype=np.int8), w.copy())
        assert eng.stats()["lost"] == 0, f"seed {seed}"
        np.testing.assert_allclose(eng.positions, dest, rtol=0, atol=1e-11, err_msg=f"seed {seed}")
        np.testing.assert_allclose(eng.flux.sum(), (np.linalg.norm(dest - start, axis=1) * w).sum(), rtol=1e-11)
        v = coords[t2v[eng.elem_ids]]
        T = np.transpose(v[:, 1:] - v[:, :1], (0, 2, 1))
        lam = np.linalg.solve(T, (eng.positions - v[:, 0])[..., None])[..., 0]
        assert min(lam.min(), (1.0 - lam.sum(1)).min()) > -1e-10, f"seed {seed}"
