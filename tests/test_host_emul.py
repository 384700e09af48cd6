"""The CUDA kernels' arithmetic (record packing, plane decode, exit-face rule,
per-ray state machine, seed-grid relocation), compiled for the host by g++ as a
TEST-ONLY build, checked against the oracle.  This is what can be verified
without a GPU; the `-m gpu` tests repeat the same comparisons through
libpumitally.so."""
import os

import numpy as np
import pytest

from helpers import (HostEmulTally, assert_flux_close, box_case, carve, edge_case_scenario, l_shaped_mesh,
                     lattice_track_scenario, non_convex_relocation_scenario, non_finite_input_scenario, run_workload,
                     unstructured_special_point_scenario)
from oracle.oracle import OraclePumiTally
from pumiumtally_b200.mesh import delaunay_box, jitter_interior, kuhn_box, tet_volumes
from pumiumtally_b200.workload import SyntheticWorkload
from test_oracle_golden import check_c1_fixture, golden_scenario

SEED = [pytest.param(dict(seed_grid=False), id="planes"), pytest.param(dict(seed_grid=True), id="planes-seeded"),
        pytest.param(dict(layout="edge"), id="edge"), pytest.param(dict(seed_grid=True, layout="edge"), id="edge-seeded")]


@pytest.mark.parametrize("seed", SEED)
def test_reference_known_answers_device_logic(seed):
    eng = golden_scenario(lambda c, t, n: HostEmulTally(c, t, n, **seed))
    st = eng.stats()
    assert st["lost"] == 0 and st["tracks"] == 7 and st["segments"] == 18


@pytest.mark.parametrize("seed", SEED)
def test_device_logic_matches_committed_c1_fixture(seed):
    eng, segs, tracks = check_c1_fixture(lambda c, t, n: HostEmulTally(c, t, n, **seed))
    assert eng.stats()["segments"] == segs and eng.stats()["tracks"] == tracks


def test_box_spec_matches_numpy_generator():
    e = HostEmulTally(spec="box:3,2,4,1.5,1.0,2.0", num_particles=1)
    c, t, v = e.mesh_arrays()
    cn, tn = kuhn_box(3, 2, 4, 1.5, 1.0, 2.0)
    np.testing.assert_array_equal(t, tn)
    np.testing.assert_allclose(c, cn, rtol=0, atol=1e-15)
    np.testing.assert_allclose(v, tet_volumes(cn, tn), rtol=1e-13)


@pytest.mark.parametrize("binary", [False, True], ids=["ascii", "binary"])
@pytest.mark.parametrize("version", ["2.2", "4.1"])
def test_gmsh_ingest(tmp_path, version, binary):
    """Gmsh .msh (2.2 / 4.1, ASCII or binary) -> same mesh as the arrays it was written from; non-tet
    elements and 1-based node ids are handled."""
    from pumiumtally_b200.mesh import save_gmsh

    c, t = jitter_interior(*kuhn_box(3, 2, 2), amplitude=0.1)
    path = str(tmp_path / "mesh.msh")
    save_gmsh(path, c, t, version=version, binary=binary)
    e = HostEmulTally(spec=path, num_particles=1)
    cm, tm, vm = e.mesh_arrays()
    np.testing.assert_array_equal(tm, t)
    np.testing.assert_array_equal(cm, c)
    np.testing.assert_allclose(vm, tet_volumes(c, t), rtol=1e-13)


OSH_CASES = [dict(), dict(compressed=False), dict(version=9, tag_layout="direct"),
             dict(version=10, tag_layout="class_ids"), dict(version=10, tag_layout="class_ids", compressed=False),
             dict(version=4, tag_layout="flags"), dict(extra_tags=False), dict(bare_stream=True, version=10),
             dict(version_in_stream=True), dict(version=8, family_byte=False), dict(version=5, family_byte=True),
             dict(version_in_stream=True, family_byte=False, compressed=False)]


@pytest.mark.parametrize("kw", OSH_CASES, ids=lambda k: ",".join(f"{a}={b}" for a, b in k.items()) or "default")
def test_osh_ingest(tmp_path, kw):
    """Omega_h .osh directory (restated stream layout, csrc/osh_reader.cpp): element order and
    coordinates survive, tet vertex SETS are recovered from tet->tri->edge->vert."""
    from pumiumtally_b200.mesh import save_osh

    c, t = delaunay_box(60)
    path = str(tmp_path / "mesh.osh")
    save_osh(path, c, t, **kw)
    e = HostEmulTally(spec=path, num_particles=1)
    cm, tm, vm = e.mesh_arrays()
    np.testing.assert_array_equal(np.sort(tm, axis=1), np.sort(t, axis=1))
    np.testing.assert_array_equal(cm, c)
    np.testing.assert_allclose(vm, tet_volumes(c, t), rtol=1e-12)
    o = OraclePumiTally(c, t, 1)  # local face order follows local vertex order: compare as sets
    np.testing.assert_array_equal(np.sort(e.adjacency, axis=1), np.sort(o.adjacency, axis=1))


def test_osh_t1_cube_walks_like_the_reference_fixture(tmp_path):
    """The reference test writes its 6-tet cube to mesh.osh and reads it back
    (test_pumi_tally_impl_methods.cpp:45-46); same round trip here, then the T1 move."""
    from pumiumtally_b200.mesh import save_osh

    c, t = kuhn_box(1, 1, 1)
    path = str(tmp_path / "mesh.osh")
    save_osh(path, c, t)
    e = HostEmulTally(spec=path, num_particles=5)
    e.CopyInitialPosition(np.tile([0.1, 0.4, 0.5], 5))
    assert (e.elem_ids == 2).all()
    fly = np.ones(5, dtype=np.int8)
    e.MoveToNextLocation(np.tile([0.1, 0.4, 0.5], 5), np.tile([1.2, 0.4, 0.5], 5), fly, np.ones(5))
    np.testing.assert_allclose(e.flux, [0, 0, 1.5, 0.5, 2.5, 0], atol=1e-12)


@pytest.mark.parametrize("damage", ["magic", "truncate", "nparts", "dim", "zlib"])
def test_osh_ingest_rejects_damaged_input(tmp_path, damage):
    from pumiumtally_b200.mesh import save_osh

    c, t = kuhn_box(2, 2, 1)
    path = str(tmp_path / "mesh.osh")
    save_osh(path, c, t)
    stream = os.path.join(path, "0.osh")
    raw = bytearray(open(stream, "rb").read())
    if damage == "magic":
        raw[0] = 0
    elif damage == "truncate":
        raw = raw[: len(raw) // 3]
    elif damage == "nparts":
        open(os.path.join(path, "nparts"), "w").write("4\n")
    elif damage == "dim":
        raw[4] = 2                      # magic(2) compressed(1) family(1) dim(1)
    elif damage == "zlib":
        raw[40:48] = b"\xff" * 8
    open(stream, "wb").write(bytes(raw))
    with pytest.raises(RuntimeError):
        HostEmulTally(spec=path, num_particles=1)


@pytest.mark.parametrize("mesh", ["kuhn", "jitter", "delaunay"])
def test_adjacency_matches_oracle(mesh):
    if mesh == "kuhn":
        c, t = kuhn_box(4, 3, 2)
    elif mesh == "jitter":
        c, t = jitter_interior(*kuhn_box(4, 3, 3), amplitude=0.15)
    else:
        c, t = delaunay_box(200)
    e = HostEmulTally(c, t, 1)
    o = OraclePumiTally(c, t, 1)
    np.testing.assert_array_equal(e.adjacency, o.adjacency)
    adj = e.adjacency
    for a in range(len(t)):  # symmetry
        for b in adj[a]:
            if b >= 0:
                assert a in adj[b]


def test_seed_grid_covers_a_box_mesh():
    c, t = kuhn_box(6, 5, 4)
    e = HostEmulTally(c, t, 1, seed_grid=True)
    nx, ny, nz = e.grid_dims()
    assert nx * ny * nz >= len(t) // 8
    assert e.grid_valid_cells == nx * ny * nz  # every seed point of a box mesh is inside it


@pytest.mark.parametrize("seed", SEED)
def test_config_c1_parity(seed):
    """BASELINE.json configs[0]: ~1k-tet cube, 10k particles (plumbing config)."""
    coords, t2v, wl = box_case((6, 6, 5), 10_000)
    eng = HostEmulTally(coords, t2v, wl.n, **seed)
    orc = OraclePumiTally(coords, t2v, wl.n)
    run_workload(eng, orc, wl, steps=4, label="c1")
    st = eng.stats()
    assert st["segments"] == orc.n_segments and st["tracks"] == orc.n_tracks and st["lost"] == 0
    if seed.get("seed_grid"):  # the seeded walks are the point: far fewer tally-off crossings than the reference walk
        plain = HostEmulTally(coords, t2v, wl.n, layout=seed.get("layout", "planes"))
        wl2 = SyntheticWorkload(box=(6.0, 6.0, 5.0), num_particles=wl.n)
        run_workload(plain, OraclePumiTally(coords, t2v, wl.n), wl2, steps=4, check_each_step=False)
        assert st["relocations"] < 0.6 * plain.stats()["relocations"]


@pytest.mark.parametrize("seed", SEED)
@pytest.mark.parametrize("mesh", ["jitter", "delaunay"])
def test_unstructured_parity(mesh, seed):
    if mesh == "jitter":
        c, t = jitter_interior(*kuhn_box(6, 5, 4), amplitude=0.18)
        box = (6.0, 5.0, 4.0)
    else:
        c, t = delaunay_box(600)
        box = (1.0, 1.0, 1.0)
    n = 4000
    wl = SyntheticWorkload(box=box, num_particles=n, mean_length=0.5 * min(box), seed=3)
    eng, orc = HostEmulTally(c, t, n, **seed), OraclePumiTally(c, t, n)
    run_workload(eng, orc, wl, steps=3, label=mesh)
    assert eng.stats()["segments"] == orc.n_segments


@pytest.mark.parametrize("seed", SEED)
def test_long_axial_tracks(seed):
    """Config c4 in miniature: forward-peaked tracks crossing many tets."""
    coords, t2v, wl = box_case((4, 4, 40), 1500, mean_length=60.0, mu_min=0.9)
    eng, orc = HostEmulTally(coords, t2v, wl.n, **seed), OraclePumiTally(coords, t2v, wl.n)
    run_workload(eng, orc, wl, steps=2, label="c4-mini")
    assert eng.stats()["segments"] / max(eng.stats()["tracks"], 1) > 15


@pytest.mark.parametrize("seed", SEED)
def test_edge_cases(seed):
    edge_case_scenario(lambda c, t, n: HostEmulTally(c, t, n, **seed))


@pytest.mark.parametrize("seed", SEED)
def test_out_of_mesh_origin_falls_back_to_reference_walk(seed):
    """A relocation target outside the hull must be clipped along the line from the OLD
    position (reference semantics), which the seeded walk cannot know: it has to fall back."""
    coords, t2v = kuhn_box(4, 4, 4)
    n = 64
    rng = np.random.default_rng(5)
    init = rng.uniform(0.2, 3.8, size=(n, 3))
    eng, orc = HostEmulTally(coords, t2v, n, **seed), OraclePumiTally(coords, t2v, n)
    for e in (eng, orc):
        e.CopyInitialPosition(init.reshape(-1).copy())
    origin = rng.uniform(-3.0, 7.0, size=(n, 3))  # many outside the box, far from the old position
    dest = rng.uniform(0.2, 3.8, size=(n, 3))
    fly = np.ones(n, dtype=np.int8)
    w = rng.uniform(0.5, 1.0, size=n)
    for e in (eng, orc):
        e.MoveToNextLocation(origin.reshape(-1).copy(), dest.reshape(-1).copy(), fly.copy(), w.copy())
    assert_flux_close(eng.flux, orc.flux, "outside-origin")
    np.testing.assert_array_equal(eng.elem_ids, orc.elem_ids)
    np.testing.assert_allclose(eng.positions, orc.positions, atol=1e-12)
    assert eng.stats()["lost"] == 0


def test_edge_walk_uses_plane_records_only_for_coplanar_rays():
    """walk_compact.cuh: an edge function is exactly zero only when the ray is coplanar with a mesh
    edge.  Random tracks never are; the axis-parallel tracks of the reference fixture and of the
    edge-case scenario are, and finish on the plane records with oracle-identical results."""
    coords, t2v, wl = box_case((6, 6, 5), 10_000)
    eng = HostEmulTally(coords, t2v, wl.n, layout="edge", seed_grid=True)
    run_workload(eng, OraclePumiTally(coords, t2v, wl.n), wl, steps=3, label="generic")
    assert eng.degenerate_rays == 0
    c, t = delaunay_box(300)
    n = 5000
    wl = SyntheticWorkload(box=(1.0, 1.0, 1.0), num_particles=n, mean_length=0.4, seed=5)
    eng = HostEmulTally(c, t, n, layout="edge")
    run_workload(eng, OraclePumiTally(c, t, n), wl, steps=3, label="delaunay")
    assert eng.degenerate_rays == 0
    assert edge_case_scenario(lambda c, t, n: HostEmulTally(c, t, n, layout="edge")).degenerate_rays > 0
    assert golden_scenario(lambda c, t, n: HostEmulTally(c, t, n, layout="edge")).degenerate_rays > 0


@pytest.mark.parametrize("seed", SEED)
def test_degenerate_starts_and_tracks_conserve_length(seed):
    """Particles on mesh vertices, edges and faces, tracks along edges, inside face planes and through
    vertices: which of the touching tets gets a piece is a tie-break, but no particle may be lost, the
    total tally must equal the total in-mesh track length and every particle must end in a tet that
    contains its final position."""
    coords, t2v = kuhn_box(3, 3, 3)
    special = np.array([[0, 0, 0], [1, 1, 1], [2, 1, 0], [1.5, 1.5, 1.5], [1.5, 1.0, 1.0], [0.5, 0.5, 1.0],
                        [1.0, 2.0, 1.0], [2.5, 2.5, 2.5], [1.0, 1.0, 0.5], [3, 3, 3], [0.25, 0.25, 0.25],
                        [2.0, 0.5, 1.5], [1.25, 1.25, 2.0]], dtype=float)
    rng = np.random.default_rng(5)
    pairs = [(a, b) for a in range(len(special)) for b in range(len(special)) if a != b]
    n = len(pairs)
    start = np.array([special[a] for a, _ in pairs])
    dest = np.array([special[b] for _, b in pairs])
    w = rng.uniform(0.5, 1.0, n)
    eng = HostEmulTally(coords, t2v, n, **seed)
    eng.CopyInitialPosition(start.reshape(-1).copy())
    assert not eng.flux.any()
    eng.MoveToNextLocation(start.reshape(-1).copy(), dest.reshape(-1).copy(), np.ones(n, dtype=np.int8), w.copy())
    assert eng.stats()["lost"] == 0
    np.testing.assert_allclose(eng.flux.sum(), (np.linalg.norm(dest - start, axis=1) * w).sum(), rtol=1e-12)
    np.testing.assert_array_equal(eng.positions, dest)  # hull corner (3,3,3) included: hull planes err outwards
    v = coords[t2v[eng.elem_ids]]
    T = np.transpose(v[:, 1:] - v[:, :1], (0, 2, 1))
    lam = np.linalg.solve(T, (eng.positions - v[:, 0])[..., None])[..., 0]
    assert min(lam.min(), (1.0 - lam.sum(1)).min()) > -1e-12
    # a second move back to generic points: the degenerate positions are valid starting states
    back = rng.uniform(0.1, 2.9, size=(n, 3))
    eng.MoveToNextLocation(dest.reshape(-1).copy(), back.reshape(-1).copy(), np.ones(n, dtype=np.int8), w.copy())
    assert eng.stats()["lost"] == 0
    np.testing.assert_array_equal(eng.positions, back)
    orc = OraclePumiTally(coords, t2v, n)
    orc.CopyInitialPosition(back.reshape(-1).copy())
    np.testing.assert_array_equal(eng.elem_ids, orc.elem_ids)


@pytest.mark.parametrize("seed", SEED)
def test_non_finite_inputs_do_not_poison_tally_or_state(seed):
    non_finite_input_scenario(lambda c, t, n: HostEmulTally(c, t, n, **seed))


@pytest.mark.parametrize("offset", [0.0, 1e4, 1e6, 1e9])
def test_meshes_far_from_the_origin_keep_their_digits(offset, capfd):
    """Face planes are stored relative to the centre of the mesh and ray origins are translated once
    (tet_mesh.hpp): the same mesh and tracks placed 1e4 .. 1e9 tet edges from the origin give the same
    tally to ~1e-12.  The yardstick is the oracle run on the nearby problem the far inputs represent
    exactly ((x + offset) - offset is exact), because the oracle itself -- plane offsets from absolute
    coordinates -- loses digits in proportion to the distance (2e-9 at 1e6).  Beyond ~5e6 tet edges a
    warning says what remains: the granularity of the caller's own double-precision coordinates."""
    c0, t = jitter_interior(*kuhn_box(5, 5, 4), amplitude=0.15)
    n = 3000

    def run(make, far):
        tr = (lambda a: a + offset) if far else (lambda a: (a + offset) - offset)
        wl = SyntheticWorkload(box=(5.0, 5.0, 4.0), num_particles=n, mean_length=2.0, seed=9)
        e = make(tr(c0), t, n)
        e.CopyInitialPosition(tr(wl.initial_positions()).reshape(-1).copy())
        for _ in range(3):
            o, d, f, w = wl.next_step()
            e.MoveToNextLocation(tr(o).reshape(-1).copy(), tr(d).reshape(-1).copy(), f.copy(), w.copy())
        return e

    capfd.readouterr()
    eng = run(lambda c, tt, m: HostEmulTally(c, tt, m, seed_grid=True), True)
    assert ("WARNING: mesh coordinates" in capfd.readouterr().err) == (offset >= 1e9)
    truth = run(lambda c, tt, m: OraclePumiTally(c, tt, m), False)
    np.testing.assert_array_equal(eng.elem_ids, truth.elem_ids)
    ref = truth.flux
    err = np.abs(eng.flux - ref) / (np.abs(ref) + 1e-12 * ref.sum())
    assert err.max() < 1e-10, f"offset {offset:g}: {err.max():.2e}"
    # positions: exact where a destination was reached, within the ulp of the far coordinates at the hull
    np.testing.assert_allclose(eng.positions - offset, truth.positions, rtol=0, atol=1e-10 + 4 * np.spacing(offset))
    assert eng.stats()["lost"] == 0


@pytest.mark.parametrize("block", range(4))
def test_randomised_meshes_and_tracks_parity(block):
    """Seeded sweep: random Delaunay / jittered / anisotropic Kuhn meshes, random particle counts, track
    lengths and collimation, both layouts, with and without the seed grid -- five moves each against the oracle."""
    for seed in range(12 * block, 12 * block + 12):
        rng = np.random.default_rng(seed)
        kind = seed % 3
        if kind == 0:
            c, t = delaunay_box(int(rng.integers(30, 400)), seed=seed)
            box = (1.0, 1.0, 1.0)
        elif kind == 1:
            dims = tuple(int(x) for x in rng.integers(1, 7, 3))
            c, t = jitter_interior(*kuhn_box(*dims), amplitude=float(rng.uniform(0, 0.3)), seed=seed)
            box = tuple(float(d) for d in dims)
        else:
            dims = tuple(int(x) for x in rng.integers(1, 9, 3))
            box = tuple(float(x) for x in rng.uniform(0.3, 5, 3))
            c, t = kuhn_box(*dims, *box)
        n = int(rng.integers(1, 3000))
        mean_length = float(rng.uniform(0.05, 3.0)) * min(box)
        mu_min = float(rng.choice([-1.0, 0.5, 0.99]))
        for layout in ("planes", "edge"):
            for seed_grid in (False, True):
                wl = SyntheticWorkload(box=box, num_particles=n, mean_length=mean_length, seed=seed, mu_min=mu_min)
                eng = HostEmulTally(c, t, n, layout=layout, seed_grid=seed_grid)
                run_workload(eng, OraclePumiTally(c, t, n), wl, steps=5, label=f"seed {seed} {layout} grid={seed_grid}")
                assert eng.stats()["lost"] == 0


@pytest.mark.parametrize("fma", [False, True], ids=["plain", "fma"])
@pytest.mark.parametrize("seed", SEED)
def test_lattice_tracks_on_hull_faces_edges_and_vertices(seed, fma):
    """Both flavours of the arithmetic: plain, and with fused multiply-adds as the device code has them."""
    lattice_track_scenario(lambda c, t, n: HostEmulTally(c, t, n, fma=fma, **seed), range(30),
                           exact_destinations=seed.get("layout", "planes") != "edge")


@pytest.mark.parametrize("seed", SEED)
def test_reference_scenarios_with_fused_multiply_adds(seed):
    """The golden, edge-case and degenerate scenarios once more with the device flavour of rounding."""
    mk = lambda c, t, n: HostEmulTally(c, t, n, fma=True, **seed)
    golden_scenario(mk)
    check_c1_fixture(mk)
    edge_case_scenario(mk)
    non_finite_input_scenario(mk)


@pytest.mark.parametrize("fmt", ["osh", "osh_raw", "msh4", "msh2", "msh4_binary", "msh2_binary"])
def test_damaged_mesh_files_are_rejected_or_loaded_never_crash(tmp_path, fmt):
    """Random byte mutations / truncations of valid mesh files: the loader must either load a valid mesh
    or report an error (the same campaign ran clean under ASan/UBSan with 14,000 mutations)."""
    from pumiumtally_b200.mesh import save_gmsh, save_osh

    c, t = kuhn_box(2, 2, 1)
    if fmt.startswith("osh"):
        path = str(tmp_path / "m.osh")
        save_osh(path, c, t, compressed=(fmt == "osh"), tag_layout="direct" if fmt == "osh" else "class_ids")
        target = os.path.join(path, "0.osh")
    else:
        path = target = str(tmp_path / "m.msh")
        save_gmsh(path, c, t, version="4.1" if fmt.startswith("msh4") else "2.2", binary=fmt.endswith("binary"))
    original = open(target, "rb").read()
    rng = np.random.default_rng(17)
    loaded = rejected = 0
    for _ in range(250):
        buf = bytearray(original)
        for _ in range(int(rng.integers(1, 6))):
            if not buf:
                break
            pos = int(rng.integers(0, len(buf)))
            mode = int(rng.integers(0, 4))
            if mode == 0:
                buf[pos] = int(rng.integers(0, 256))
            elif mode == 1:
                buf[pos] ^= 1 << int(rng.integers(0, 8))
            elif mode == 2:
                del buf[int(rng.integers(0, len(buf))):]
            else:
                buf[pos:pos + 8] = b"\\xff" * min(8, len(buf) - pos)
        open(target, "wb").write(bytes(buf))
        try:
            e = HostEmulTally(spec=path, num_particles=1)
            assert e.num_elements >= 1
            loaded += 1
        except RuntimeError:
            rejected += 1
    assert loaded + rejected == 250 and rejected > 50


@pytest.mark.parametrize("fma", [False, True], ids=["plain", "fma"])
@pytest.mark.parametrize("seed", SEED)
def test_tracks_through_vertices_and_along_edges_of_unstructured_meshes(seed, fma):
    unstructured_special_point_scenario(lambda c, t, n: HostEmulTally(c, t, n, fma=fma, **seed), range(24))


@pytest.mark.parametrize("fma", [False, True], ids=["plain", "fma"])
def test_extreme_meshes_and_batches_parity(fma):
    """Needle- and plate-shaped cells (aspect ratios 1e-3 .. 1e6), tracks of 1e-9 cell sizes, batches that mostly
    leave the mesh, batches where every particle is re-sourced: four moves each against the oracle, both layouts."""
    for seed in range(24):
        rng = np.random.default_rng(5000 + seed)
        kind = seed % 4
        if kind == 0:
            c, t = delaunay_box(int(rng.integers(50, 300)), seed=seed)
            hi = np.ones(3)
        elif kind == 1:
            dims = (1, 1, int(rng.integers(5, 40)))
            hi = np.array([0.01, 0.01, float(dims[2])])
            c, t = kuhn_box(*dims, *hi)
        elif kind == 2:
            dims = tuple(int(x) for x in rng.integers(2, 6, 3))
            c, t = jitter_interior(*kuhn_box(*dims), amplitude=0.25, seed=seed)
            hi = np.array(dims, dtype=float)
        else:
            dims = (int(rng.integers(10, 30)), 1, 1)
            hi = np.array([float(dims[0]), 1e-3, 1e3])
            c, t = kuhn_box(*dims, *hi)
        n = int(rng.integers(50, 1500))
        pos = rng.uniform(0.01, 0.99, (n, 3)) * hi
        for layout in ("planes", "edge"):
            eng, orc = HostEmulTally(c, t, n, layout=layout, seed_grid=bool(seed & 1), fma=fma), OraclePumiTally(c, t, n)
            for e in (eng, orc):
                e.CopyInitialPosition(pos.reshape(-1).copy())
            np.testing.assert_array_equal(eng.elem_ids, orc.elem_ids)
            cur, r2 = pos.copy(), np.random.default_rng(seed)
            for step in range(4):
                mode = int(r2.integers(0, 4))
                origin = cur.copy()
                res = r2.random(n) < (1.0 if mode == 3 else 0.1)
                origin[res] = r2.uniform(0.01, 0.99, (int(res.sum()), 3)) * hi
                if mode == 0:
                    dest = origin + r2.normal(0, 1e-9, (n, 3)) * hi
                elif mode == 1:
                    dest = r2.uniform(-2, 3, (n, 3)) * hi
                else:
                    dest = r2.uniform(0.0, 1.0, (n, 3)) * hi
                fly, w = (r2.random(n) < 0.9).astype(np.int8), r2.uniform(0, 2, n)
                for e in (eng, orc):
                    e.MoveToNextLocation(origin.reshape(-1).copy(), dest.reshape(-1).copy(), fly.copy(), w.copy())
                cur = orc.positions.copy()
                assert_flux_close(eng.flux, orc.flux, f"seed {seed} {layout} step {step}")
                np.testing.assert_array_equal(eng.elem_ids, orc.elem_ids)
                assert np.abs(eng.positions - orc.positions).max() <= 1e-9 * max(1.0, hi.max())
                assert eng.stats()["lost"] == 0


def test_hull_convexity_is_recognised():
    """Convex: Kuhn boxes (also jittered inside), Delaunay meshes of a point cloud.  Not convex: an L-shaped
    prism, a box with a void, two boxes touching along an edge."""
    for c, t in (kuhn_box(3, 2, 2), jitter_interior(*kuhn_box(4, 3, 3), amplitude=0.2), delaunay_box(200, seed=4)):
        assert HostEmulTally(c, t, 1).hull_convex
    assert not HostEmulTally(*l_shaped_mesh(), 1).hull_convex
    c, t = kuhn_box(5, 5, 5)
    cen = c[t].mean(1)
    void = (np.abs(cen - 2.5) < 0.5).all(1)
    assert not HostEmulTally(*carve(c, t, ~void), 1).hull_convex
    two = ((cen[:, 0] < 2) & (cen[:, 1] < 2)) | ((cen[:, 0] > 2) & (cen[:, 1] > 2))
    assert not HostEmulTally(*carve(c, t, two), 1).hull_convex


def test_relocation_across_a_concavity_follows_the_reference_walk():
    """No seed grid in the host build by default: the plain walk is the reference's."""
    non_convex_relocation_scenario(lambda c, t, n: HostEmulTally(c, t, n))
