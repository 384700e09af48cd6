"""world_size-2 gloo test of the N>1 host logic: particle stripes with counter-based
batches, per-rank tallies on a replicated mesh, batch-end sum == single-rank tally."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pumiumtally_b200.distributed import allreduce_sum_host, broadcast_unique_id, particle_stripe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, STEPS, CELLS = 3001, 3, (5, 4, 3)


def _tally(begin, end):
    from oracle.oracle import OraclePumiTally
    from pumiumtally_b200.mesh import kuhn_box
    from pumiumtally_b200.workload import SyntheticWorkload

    coords, t2v = kuhn_box(*CELLS)
    n = end - begin
    wl = SyntheticWorkload(box=tuple(float(c) for c in CELLS), num_particles=n, mean_length=2.0, id_offset=begin)
    o = OraclePumiTally(coords, t2v, n)
    o.CopyInitialPosition(wl.initial_positions().reshape(-1))
    for _ in range(STEPS):
        a, b, f, w = wl.next_step()
        o.MoveToNextLocation(a.reshape(-1), b.reshape(-1), f, w)
    return o.flux, o.n_segments


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = broadcast_unique_id(dist, lambda: bytes(range(128)))
    assert uid == bytes(range(128))
    b, e = particle_stripe(N_TOTAL, rank, world)
    flux, segs = _tally(b, e)
    total = allreduce_sum_host(dist, flux)
    s = torch.tensor([segs], dtype=torch.int64)
    dist.all_reduce(s)
    if rank == 0:
        q.put((total, int(s[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_stripes_cover_everything():
    for n, w in [(10, 3), (3001, 2), (5, 8), (0, 2)]:
        edges = [particle_stripe(n, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        sizes = [e - b for b, e in edges]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_tally_sums_to_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, segs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_flux, ref_segs = _tally(0, N_TOTAL)
    np.testing.assert_allclose(total, ref_flux, rtol=1e-12, atol=1e-12)
    assert segs == ref_segs
