"""Spatial partition (pumiumtally_b200/partition.py): bisection, picparts with ghost layers, and the
routing / hand-off / ghost-exchange logic of PartitionedTally on two gloo ranks with the oracle as
the walker -- against one oracle holding the whole mesh and all particles."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.oracle import OraclePumiTally
from pumiumtally_b200.mesh import delaunay_box, jitter_interior, kuhn_box
from pumiumtally_b200.partition import (OracleWalker, PartitionedTally, Picpart, face_adjacency, rcb_locate,
                                        rcb_partition, tet_centroids)
from pumiumtally_b200.workload import SyntheticWorkload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CELLS, N_TOTAL, STEPS = (8, 6, 5), 4000, 4


@pytest.mark.parametrize("nparts", [2, 3, 8])
def test_bisection_is_balanced_and_locates_centroids(nparts):
    coords, t2v = kuhn_box(7, 5, 4)
    c = tet_centroids(coords, t2v)
    part, tree = rcb_partition(c, nparts)
    sizes = np.bincount(part, minlength=nparts)
    assert sizes.sum() == len(t2v) and sizes.max() - sizes.min() <= nparts
    # a centroid is located in its own part unless it sits exactly on a cut (Kuhn centroids come in
    # planes, so whole planes of them can sit on a cut)
    assert (rcb_locate(tree, c) == part).mean() > 0.85


@pytest.mark.parametrize("mesh", ["kuhn", "delaunay"])
def test_face_adjacency_matches_the_oracles(mesh):
    coords, t2v = kuhn_box(4, 3, 3) if mesh == "kuhn" else delaunay_box(150, seed=3)
    t2t = face_adjacency(t2v, len(coords))
    np.testing.assert_array_equal(t2t, OraclePumiTally(coords, t2v, 1).adjacency)


@pytest.mark.parametrize("layers", [1, 3])
def test_picparts_cover_the_mesh_with_ghost_layers(layers):
    coords, t2v = jitter_interior(*kuhn_box(6, 5, 4), amplitude=0.1)
    t2t = face_adjacency(t2v, len(coords))
    part, _ = rcb_partition(tet_centroids(coords, t2v), 4)
    owned_all = []
    for r in range(4):
        p = Picpart(coords, t2v, t2t, part, r, layers)
        owned = p.global_of_local[: p.n_owned]
        owned_all.append(owned)
        assert (part[owned] == r).all() and (part[p.global_of_local[p.n_owned:]] != r).all()
        # local mesh is the same geometry
        np.testing.assert_array_equal(p.coords[p.t2v], coords[t2v[p.global_of_local]])
        # faces: interior (-2) iff the neighbour is local; hull (-1) iff the global face is hull
        nb = t2t[p.global_of_local]
        assert ((p.face_next_global == -1) == (nb < 0)).all()
        remote = p.face_next_global >= 0
        assert (p.local_of_global[p.face_next_global[remote]] < 0).all()
        # every owned tet is at least `layers` face-steps away from the picpart's outer boundary
        depth = np.full(p.n_local, 10**6)
        depth[remote.any(1)] = 0
        lt2t = np.where(nb >= 0, p.local_of_global[np.maximum(nb, 0)], -1)
        for _ in range(layers + 1):
            for f in range(4):
                ok = lt2t[:, f] >= 0
                depth[ok] = np.minimum(depth[ok], depth[lt2t[ok, f]] + 1)
        assert depth[: p.n_owned].min() >= layers
        planes = p.face_planes()
        cen = p.coords[p.t2v].mean(1)
        assert ((planes[:, :, :3] * cen[:, None, :]).sum(2) < planes[:, :, 3]).all()  # outward
    assert np.array_equal(np.sort(np.concatenate(owned_all)), np.arange(len(t2v)))


def _reference():
    coords, t2v = kuhn_box(*CELLS)
    wl = SyntheticWorkload(box=tuple(float(c) for c in CELLS), num_particles=N_TOTAL, mean_length=2.5)
    o = OraclePumiTally(coords, t2v, N_TOTAL)
    o.CopyInitialPosition(wl.initial_positions().reshape(-1))
    for _ in range(STEPS):
        a, b, f, w = wl.next_step()
        o.MoveToNextLocation(a.reshape(-1), b.reshape(-1), f, w)
    return o.flux, o.elem_ids, o.positions, o.n_segments


def _worker(rank, world, port, layers, q):
    sys.path.insert(0, ROOT)
    from pumiumtally_b200.distributed import particle_stripe

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    coords, t2v = kuhn_box(*CELLS)
    b, e = particle_stripe(N_TOTAL, rank, world)
    wl = SyntheticWorkload(box=tuple(float(c) for c in CELLS), num_particles=e - b, mean_length=2.5, id_offset=b)
    pt = PartitionedTally(coords, t2v, e - b, dist, torch.device("cpu"), layers=layers, walker=OracleWalker,
                          capacity_factor=2.5)
    pt.CopyInitialPosition(torch.from_numpy(wl.initial_positions()))
    for _ in range(STEPS):
        o, d, f, w = (torch.from_numpy(np.ascontiguousarray(x)) for x in wl.next_step())
        pt.MoveToNextLocation(o, d, f, w)
        assert not bool(f.any())
    flux = pt.global_flux()
    q.put((rank, b, e, flux.numpy(), pt.elem_ids.numpy(), pt.positions.numpy(), pt.stats_handoffs, pt.stats_rounds,
           pt.pic.n_local, pt.ghost_values_sent))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("layers", [1, 2])
def test_two_rank_partitioned_tally_equals_single_mesh_tally(layers):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 1500) + layers
    procs = [ctx.Process(target=_worker, args=(r, 2, port, layers, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_flux, ref_elem, ref_pos, _ = _reference()
    handoffs = 0
    for rank, b, e, flux, elem, pos, nh, rounds, n_local, n_ghost in got:
        # a handed-over track is re-started at the crossing point: same pieces, re-parametrised
        tol = 1e-9 * np.abs(ref_flux) + 1e-12 * ref_flux.sum()
        assert (np.abs(flux - ref_flux) <= tol).all()
        np.testing.assert_array_equal(elem, ref_elem[b:e])
        np.testing.assert_allclose(pos, ref_pos[b:e], rtol=0, atol=1e-12)
        assert n_local < len(ref_flux) and n_ghost < n_local  # a piece of the mesh, a thin ghost exchange
        assert rounds >= STEPS
        handoffs += nh
    assert handoffs > 0  # tracks did cross the partition boundary
