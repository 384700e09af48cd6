"""Host-side helpers for the multi-GPU layout (one process per GPU).

Every rank holds the full mesh (full-buffer picpart) and a contiguous stripe of
the global particle index space; the only exchange is the batch-end sum of the
per-rank tallies.  The reference has no counterpart: it
runs full-mesh replicas without any flux reduction (PumiTallyImpl.cpp:530-539).
"""
from __future__ import annotations


def particle_stripe(n_total: int, rank: int, world: int):
    """Contiguous [begin, end) stripe of rank; stripes differ in size by at most one."""
    base, rem = divmod(int(n_total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_unique_id(dist, make_id, device="cpu") -> bytes:
    """Rank 0 creates the 128-byte NCCL unique id (``make_id()``), everyone receives it."""
    import torch

    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        t.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def allreduce_sum_host(dist, array):
    """Sum a numpy array over ranks through torch.distributed (gloo in the CPU tests;
    on GPUs the engine does this itself with ncclAllReduce on the flux in HBM)."""
    import torch

    t = torch.from_numpy(array.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
