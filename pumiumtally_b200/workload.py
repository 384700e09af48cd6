"""Deterministic synthetic particle batches.

Stands in for the batches OpenMC hands to ``MoveToNextLocation``
(reference: src/pumitally/PumiTally.h:70-89): straight tracks with isotropic
(or forward-peaked) directions and exponential lengths inside an axis-aligned
box mesh.  A track whose destination lies outside the box is clipped by the
engine's vacuum boundary; the generator then re-samples that particle at a
fresh uniform position on its next flight, which exercises the "relocate to
origin without tallying" phase (reference: PumiTallyImpl.cpp:71-112).

Random numbers are counter based -- SplitMix64 of (seed, step, particle id,
stream) -- so numpy (uint64) and torch (int64, any device) produce the same
bits and any particle sub-range can be generated independently.
"""
from __future__ import annotations

import math

import numpy as np

_GOLD = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_K_STEP = 0xD1B54A32D192ED03
_K_STREAM = 0x8CB92BA72F3D8DD7


def _s64(x: int) -> int:
    """Python int -> signed 64-bit value with the same bit pattern."""
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


class _NumpyOps:
    name = "numpy"

    def __init__(self, device=None):
        pass

    def ids(self, begin, end):
        return np.arange(begin, end, dtype=np.uint64)

    def uniform(self, ids, seed, step, stream):
        with np.errstate(over="ignore"):
            base = np.uint64((seed + step * _K_STEP + stream * _K_STREAM) & ((1 << 64) - 1))
            z = ids * np.uint64(_GOLD) + base
            z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
            z = z ^ (z >> np.uint64(31))
        return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)

    where = staticmethod(np.where)
    sqrt = staticmethod(np.sqrt)
    cos = staticmethod(np.cos)
    sin = staticmethod(np.sin)
    log = staticmethod(np.log)

    def stack3(self, x, y, z):
        return np.stack([x, y, z], axis=1)

    def full_like_bool(self, a, v):
        return np.full(a.shape[0], v, dtype=bool)

    def to_int8(self, m):
        return m.astype(np.int8)

    def clone(self, a):
        return a.copy()


class _TorchOps:
    name = "torch"

    def __init__(self, device=None):
        import torch

        self.t = torch
        self.device = torch.device(device or "cpu")

    def ids(self, begin, end):
        return self.t.arange(begin, end, dtype=self.t.int64, device=self.device)

    def _lsr(self, z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)

    def uniform(self, ids, seed, step, stream):
        base = _s64(seed + step * _K_STEP + stream * _K_STREAM)
        z = ids * _s64(_GOLD) + base
        z = (z ^ self._lsr(z, 30)) * _s64(_M1)
        z = (z ^ self._lsr(z, 27)) * _s64(_M2)
        z = z ^ self._lsr(z, 31)
        return self._lsr(z, 11).to(self.t.float64) * (2.0 ** -53)

    def where(self, c, a, b):
        return self.t.where(c, a, b)

    def sqrt(self, a):
        return self.t.sqrt(a)

    def cos(self, a):
        return self.t.cos(a)

    def sin(self, a):
        return self.t.sin(a)

    def log(self, a):
        return self.t.log(a)

    def stack3(self, x, y, z):
        return self.t.stack([x, y, z], dim=1)

    def full_like_bool(self, a, v):
        return self.t.full((a.shape[0],), v, dtype=self.t.bool, device=self.device)

    def to_int8(self, m):
        return m.to(self.t.int8)

    def clone(self, a):
        return a.clone()


class SyntheticWorkload:
    """Particle batches for an axis-aligned box mesh ``[0,lx]x[0,ly]x[0,lz]``.

    Parameters
    ----------
    box : (lx, ly, lz)
    num_particles : N
    mean_length : mean of the exponential track length (same units as the box)
    mu_min : directions have cos(theta) = mu in [mu_min, 1] w.r.t. +z
             (-1 = isotropic; 0.9 gives the long axial tracks of config c4)
    fly_prob : probability that a particle flies in a given step
    id_offset : global id of local particle 0 (multi-GPU striping)
    """

    def __init__(self, box, num_particles, seed=0x5EED, mean_length=3.0, mu_min=-1.0,
                 fly_prob=0.95, backend="numpy", device=None, id_offset=0):
        self.box = tuple(float(b) for b in box)
        self.n = int(num_particles)
        self.seed = int(seed)
        self.mean_length = float(mean_length)
        self.mu_min = float(mu_min)
        self.fly_prob = float(fly_prob)
        self.ops = _TorchOps(device) if backend == "torch" else _NumpyOps()
        self.ids = self.ops.ids(id_offset, id_offset + self.n)
        self.cur = None        # position the engine holds for each particle
        self.resample = None   # particle left the box: next flight starts somewhere new
        self.step_index = 0

    def _uniform_positions(self, step, stream0):
        o, eps = self.ops, 1e-6
        cols = []
        for k, L in enumerate(self.box):
            u = o.uniform(self.ids, self.seed, step, stream0 + k)
            cols.append((eps + (1.0 - 2.0 * eps) * u) * L)
        return o.stack3(*cols)

    def initial_positions(self):
        """Positions for ``CopyInitialPosition`` ([N,3], strictly inside the box)."""
        self.cur = self._uniform_positions(0, 0)
        self.resample = self.ops.full_like_bool(self.cur, False)
        self.step_index = 0
        return self.ops.clone(self.cur)

    def next_step(self):
        """Returns (origin[N,3], dest[N,3], flying int8[N], weights[N]) for the next move."""
        assert self.cur is not None, "call initial_positions() first"
        o = self.ops
        self.step_index += 1
        s = self.step_index
        fly = o.uniform(self.ids, self.seed, s, 0) < self.fly_prob
        fresh = self._uniform_positions(s, 1)
        origin = o.where((fly & self.resample)[:, None], fresh, self.cur)
        mu = self.mu_min + (1.0 - self.mu_min) * o.uniform(self.ids, self.seed, s, 4)
        phi = (2.0 * math.pi) * o.uniform(self.ids, self.seed, s, 5)
        # u in [0,1) -> -log(1-u) is finite and >= 0
        length = -self.mean_length * o.log(1.0 - o.uniform(self.ids, self.seed, s, 6))
        st = o.sqrt(1.0 - mu * mu)
        direction = o.stack3(st * o.cos(phi), st * o.sin(phi), mu)
        dest = o.where(fly[:, None], origin + length[:, None] * direction, origin)
        weights = 0.5 + 0.5 * o.uniform(self.ids, self.seed, s, 7)
        inside = ((dest[:, 0] >= 0.0) & (dest[:, 0] <= self.box[0]) &
                  (dest[:, 1] >= 0.0) & (dest[:, 1] <= self.box[1]) &
                  (dest[:, 2] >= 0.0) & (dest[:, 2] <= self.box[2]))
        # engine state after the move: flying particles sit at dest (or on the hull,
        # in which case they are re-sampled before their next flight)
        self.cur = o.where(fly[:, None], dest, self.cur)
        self.resample = o.where(fly, ~inside, self.resample)
        return origin, dest, o.to_int8(fly), weights


# BASELINE.json configs -> box cells (nx,ny,nz), total particles, mean track length, mu_min, and the
# GPU count the config is quoted on (bench.py gives every GPU particles // gpus of them).
CONFIGS = {
    "c1": dict(cells=(6, 6, 5), particles=10_000, mean_length=3.0, mu_min=-1.0, gpus=1),
    "c2": dict(cells=(55, 55, 55), particles=10_000_000, mean_length=3.0, mu_min=-1.0, gpus=1),
    "c3": dict(cells=(20, 20, 20), particles=100_000_000, mean_length=3.0, mu_min=-1.0, gpus=1),
    # long axial tracks: a 5.7-degree cone and tracks that run to the hull (~3 tets per cell layer)
    "c4": dict(cells=(32, 32, 163), particles=1_000_000, mean_length=1000.0, mu_min=0.995, gpus=4),
    "c5": dict(cells=(118, 118, 118), particles=100_000_000, mean_length=3.0, mu_min=-1.0, gpus=8),
}
