"""Spatially partitioned multi-GPU tally (prototype of SURVEY.md section 8e rows 2-3).

The reference prepares for a partitioned mesh but never partitions it
(``PartitionMesh`` assigns every element to rank 0 and every rank holds the whole
mesh: reference src/pumitally/PumiTallyImpl.cpp:530-539, 238-241).  This module
builds what that hook stands for and measures it against the replica scheme the
engine uses by default:

* ``rcb_partition``      recursive coordinate bisection of the tet centroids into G parts;
* ``build_picpart``      rank g's picpart = the tets it owns + ``layers`` ghost layers (face
                         neighbours), renumbered locally, with the global id of the tet behind
                         every face on the picpart's outer boundary;
* ``PartitionedTally``   one process per GPU.  Every rank runs an ordinary engine on its picpart
                         (160 MB of tet records instead of 1.26 GB for config c5).  Per move the
                         caller's particles are routed (all-to-all over NCCL) to the rank that
                         owns their parent tet, walked there, handed on to the next owner when a
                         track leaves the picpart (the picpart boundary acts as the engine's
                         vacuum boundary, the rest of the track travels as a new record), and their
                         final state is routed back to the rank whose caller owns them.  At batch
                         end only the ghost-layer tallies are exchanged (added into the owners'
                         copies); ``global_flux`` assembles the owned values.

Everything here is host-side orchestration over the C ABI (engine per picpart, device tensors,
``torch.distributed``); results equal the single-GPU tally to rounding (a track that is handed
over is re-started at the crossing point, so its pieces are parametrised from there).
The walker behind the driver is pluggable so that the routing / hand-off / ghost-exchange logic
runs on CPU under ``gloo`` with the oracle in the tests.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- mesh side (numpy)


def tet_centroids(coords, t2v):
    return coords[t2v].mean(axis=1)


def rcb_partition(centroids, nparts):
    """Recursive coordinate bisection: part id per tet, parts of (nearly) equal size.  Returns
    (part int32[E], tree) where tree is a list of (axis, cut, left, right) nodes / leaf part ids
    usable by ``rcb_locate`` to map arbitrary points to parts."""
    n = len(centroids)
    part = np.zeros(n, dtype=np.int32)
    nodes = []

    def rec(idx, p0, k):
        if k == 1:
            part[idx] = p0
            return -(p0 + 1)  # leaf: encoded as negative
        kl = k // 2
        c = centroids[idx]
        ext = c.max(0) - c.min(0)
        axis = int(np.argmax(ext))
        nl = int(round(len(idx) * kl / k))
        order = np.argpartition(c[:, axis], nl) if 0 < nl < len(idx) else np.arange(len(idx))
        left, right = idx[order[:nl]], idx[order[nl:]]
        cut = 0.5 * (centroids[left, axis].max() + centroids[right, axis].min()) if len(left) and len(right) else 0.0
        me = len(nodes)
        nodes.append(None)
        a = rec(left, p0, kl)
        b = rec(right, p0 + kl, k - kl)
        nodes[me] = (axis, float(cut), a, b)
        return me

    root = rec(np.arange(n), 0, int(nparts))
    return part, (nodes, root)


def rcb_locate(tree, points):
    """Part of each point under the bisection tree (points on a cut go right)."""
    nodes, root = tree
    out = np.empty(len(points), dtype=np.int32)

    def rec(node, idx):
        if node < 0:
            out[idx] = -node - 1
            return
        axis, cut, a, b = nodes[node]
        m = points[idx, axis] < cut
        rec(a, idx[m])
        rec(b, idx[~m])

    rec(root, np.arange(len(points)))
    return out


def face_adjacency(t2v, nverts):
    """t2t[e, f] = tet across the face opposite local vertex f, -1 on the hull."""
    t2v = np.asarray(t2v, dtype=np.int64)
    E = len(t2v)
    bits = max(int(nverts - 1).bit_length(), 1)
    tris = [np.sort(np.delete(t2v, f, axis=1), axis=1) for f in range(4)]
    if 3 * bits <= 63:  # the sorted vertex triple packs into one 64-bit key
        keys = np.empty((E, 4), dtype=np.int64)
        for f in range(4):
            keys[:, f] = (tris[f][:, 0] << (2 * bits)) | (tris[f][:, 1] << bits) | tris[f][:, 2]
        flat = keys.ravel()
        order = np.argsort(flat, kind="stable")
        s = flat[order]
    else:  # huge meshes: sort on the three columns
        tri = np.stack(tris, axis=1).reshape(-1, 3)
        order = np.lexsort((tri[:, 2], tri[:, 1], tri[:, 0]))
        t = tri[order]
        s = np.concatenate([[0], np.cumsum((t[1:] != t[:-1]).any(1))])  # equal triples get equal ranks
    same = s[1:] == s[:-1]
    t2t = np.full(4 * E, -1, dtype=np.int64)
    a, b = order[:-1][same], order[1:][same]
    t2t[a] = b // 4
    t2t[b] = a // 4
    return t2t.reshape(E, 4)


class Picpart:
    """Rank-local piece of the mesh (see module docstring).  Local ids: owned tets first (ascending
    global id), then ghost layer 1, 2, ..."""

    def __init__(self, coords, t2v, t2t, part, rank, layers):
        E = len(t2v)
        owned = np.flatnonzero(part == rank)
        in_set = np.zeros(E, dtype=bool)
        in_set[owned] = True
        pieces, frontier = [owned], owned
        for _ in range(int(layers)):
            nb = t2t[frontier].ravel()
            nb = nb[nb >= 0]
            new = np.unique(nb[~in_set[nb]])
            in_set[new] = True
            pieces.append(new)
            frontier = new
        self.global_of_local = np.concatenate(pieces).astype(np.int64)
        self.n_owned = len(owned)
        self.n_local = len(self.global_of_local)
        self.local_of_global = np.full(E, -1, dtype=np.int64)
        self.local_of_global[self.global_of_local] = np.arange(self.n_local)
        lt2v = t2v[self.global_of_local]
        verts, inv = np.unique(lt2v.ravel(), return_inverse=True)
        self.coords = np.ascontiguousarray(coords[verts])
        self.t2v = inv.reshape(-1, 4).astype(np.int32)
        nb = t2t[self.global_of_local]  # [n_local, 4] global neighbours
        nb_local = np.where(nb >= 0, self.local_of_global[np.maximum(nb, 0)], -1)
        # behind each face: -2 = another tet of this picpart, -1 = the true hull, >= 0 = global id of
        # the tet beyond the picpart's boundary
        self.face_next_global = np.where(nb < 0, -1, np.where(nb_local >= 0, -2, nb)).astype(np.int64)
        self.rank, self.layers = rank, int(layers)

    def face_planes(self):
        """Outward unit normals and offsets of the four faces of every local tet: [n_local, 4, 4]."""
        v = self.coords[self.t2v]  # [n, 4, 3]
        planes = np.empty((self.n_local, 4, 4))
        for f in range(4):
            idx = [i for i in range(4) if i != f]
            a, b, c = v[:, idx[0]], v[:, idx[1]], v[:, idx[2]]
            n = np.cross(b - a, c - a)
            n /= np.linalg.norm(n, axis=1, keepdims=True)
            off = (n * a).sum(1)
            flip = (n * v[:, f]).sum(1) - off > 0
            n[flip] *= -1
            off[flip] *= -1
            planes[:, f, :3], planes[:, f, 3] = n, off
        return planes


# --------------------------------------------------------------------------- walkers


class GpuWalker:
    """The CUDA engine on one picpart (device tensors in, device tensors out)."""

    def __init__(self, pic, capacity, device):
        import torch

        from .tally import PumiTally

        self.torch, self.dev, self.cap = torch, device, int(capacity)
        self.eng = PumiTally.from_arrays(pic.coords, pic.t2v, self.cap, device=device.index)
        self.n_local = pic.n_local
        self._fly = torch.zeros(self.cap, dtype=torch.int8, device=device)
        self._elem = torch.empty(self.cap, dtype=torch.int32, device=device)
        self._pad3 = torch.zeros((self.cap, 3), dtype=torch.float64, device=device)
        self._pad1 = torch.zeros(self.cap, dtype=torch.float64, device=device)
        self._flux = torch.zeros(self.n_local, dtype=torch.float64, device=device)
        self._park = torch.as_tensor(pic.coords[pic.t2v[0]].mean(0), device=device)

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def _padded(self, a, buf):
        buf[: len(a)] = a
        return buf

    def localise(self, xyz):
        m = len(xyz)
        assert m <= self.cap, f"picpart engine capacity {self.cap} < {m} particles"
        buf = self._pad3
        buf[:] = self._park
        buf[:m] = xyz
        self.eng.copy_initial_position_device(buf.data_ptr(), self._stream())
        return self.state(m)

    def state(self, m):
        pos = self.torch.empty((m, 3), dtype=self.torch.float64, device=self.dev)
        self.eng.get_state_device(pos.data_ptr(), self._elem.data_ptr(), 0, m, self._stream())
        return pos, self._elem[:m].to(self.torch.int64)

    def walk(self, pos, elem, origin, dest, w):
        m = len(pos)
        assert m <= self.cap, f"picpart engine capacity {self.cap} < {m} particles"
        if m == 0:
            return pos, elem
        t = self.torch
        self._elem[:m] = elem.to(t.int32)
        self.eng.set_state_device(pos.contiguous().data_ptr(), self._elem.data_ptr(), 0, m, self._stream())
        self._fly.zero_()
        self._fly[:m] = 1
        o = t.empty((self.cap, 3), dtype=t.float64, device=self.dev)
        d = t.empty((self.cap, 3), dtype=t.float64, device=self.dev)
        ww = t.zeros(self.cap, dtype=t.float64, device=self.dev)
        o[:m], d[:m], ww[:m] = origin, dest, w
        self.eng.move_device(o.data_ptr(), d.data_ptr(), self._fly.data_ptr(), ww.data_ptr(), self._stream())
        return self.state(m)

    def flux(self):
        self.eng.get_flux_device(self._flux.data_ptr(), self._stream())
        return self._flux.clone()

    def stats(self):
        return self.eng.stats()


class OracleWalker:
    """The CPU oracle on one picpart (tests of the driver logic under gloo; test infrastructure)."""

    def __init__(self, pic, capacity, device):
        import torch

        from oracle.oracle import OraclePumiTally

        self.torch, self.cap = torch, int(capacity)
        self.orc = OraclePumiTally(pic.coords, pic.t2v, self.cap)
        self.n_local = pic.n_local
        self._park = pic.coords[pic.t2v[0]].mean(0)

    def localise(self, xyz):
        m = len(xyz)
        assert m <= self.cap
        buf = np.tile(self._park, (self.cap, 1))
        buf[:m] = xyz.numpy()
        self.orc.CopyInitialPosition(buf.reshape(-1))
        return self.state(m)

    def state(self, m):
        t = self.torch
        return t.from_numpy(self.orc.positions[:m].copy()), t.from_numpy(self.orc.elem_ids[:m].astype(np.int64))

    def walk(self, pos, elem, origin, dest, w):
        m = len(pos)
        assert m <= self.cap
        if m == 0:
            return pos, elem
        self.orc.set_state(pos.numpy(), elem.numpy().astype(np.int32))
        o = np.zeros((self.cap, 3)); d = np.zeros((self.cap, 3)); ww = np.zeros(self.cap)
        f = np.zeros(self.cap, dtype=np.int8)
        o[:m], d[:m], ww[:m], f[:m] = origin.numpy(), dest.numpy(), w.numpy(), 1
        self.orc.MoveToNextLocation(o.reshape(-1), d.reshape(-1), f, ww)
        return self.state(m)

    def flux(self):
        return self.torch.from_numpy(self.orc.flux.copy())

    def stats(self):
        return {"segments": self.orc.n_segments, "lost": self.orc.n_lost}


# --------------------------------------------------------------------------- the driver


class PartitionedTally:
    """Mirror of the PumiTally interface (CopyInitialPosition / MoveToNextLocation on this rank's
    particles) over a spatially partitioned mesh.  Arrays are torch tensors on ``device``."""

    def __init__(self, coords, t2v, num_particles, dist, device, layers=2, capacity_factor=1.6,
                 walker=GpuWalker, min_capacity=1024):
        import torch

        self.torch, self.dist, self.dev = torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        t2v = np.ascontiguousarray(t2v, dtype=np.int64)
        self.n = int(num_particles)
        self.num_elements = len(t2v)
        part, self.tree = rcb_partition(tet_centroids(coords, t2v), self.world)
        t2t = face_adjacency(t2v, len(coords))
        self.pic = Picpart(coords, t2v, t2t, part, self.rank, layers)
        del t2t
        to = lambda a, dt: torch.as_tensor(a, dtype=dt, device=device)
        self.owner_of_tet = to(part, torch.int64)
        self.local_of_global = to(self.pic.local_of_global, torch.int64)
        self.global_of_local = to(self.pic.global_of_local, torch.int64)
        self.face_next_global = to(self.pic.face_next_global, torch.int64)
        self.planes = to(self.pic.face_planes(), torch.float64)
        # every rank may have to host the particles of all ranks that sit in its part
        n_all = torch.tensor([self.n], dtype=torch.int64, device=device)
        dist.all_reduce(n_all)
        cap = max(int(capacity_factor * int(n_all) / self.world), min_capacity)
        self.walker = walker(self.pic, cap, device)
        self.pos = torch.zeros((self.n, 3), dtype=torch.float64, device=device)
        self.gelem = torch.zeros(self.n, dtype=torch.int64, device=device)
        self.stats_rounds, self.stats_handoffs, self.stats_routed = 0, 0, 0
        self.profile = False  # True: synchronise around every phase and accumulate wall time in self.timers
        self.timers = {}

    def _tick(self, name, t0):
        if not self.profile:
            return 0.0
        import time

        if self.dev.type == "cuda":
            self.torch.cuda.synchronize()
        now = time.perf_counter()
        if name:
            self.timers[name] = self.timers.get(name, 0.0) + (now - t0)
        return now

    # ---- exchange: rows of several tensors to the rank given per row --------------------------
    def _exchange(self, dest_rank, fields):
        t, dist = self.torch, self.dist
        order = t.argsort(dest_rank, stable=True)
        counts = t.bincount(dest_rank, minlength=self.world)
        recv_counts = t.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts)
        in_splits, out_splits = counts.tolist(), recv_counts.tolist()
        total = int(sum(out_splits))
        out = []
        for f in fields:
            src = f[order].contiguous()
            dst = t.empty((total,) + tuple(f.shape[1:]), dtype=f.dtype, device=f.device)
            dist.all_to_all_single(dst, src, out_splits, in_splits)
            out.append(dst)
        return out

    def _global_sum(self, v):
        x = self.torch.tensor([int(v)], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(x)
        return int(x)

    # ---- reference interface -------------------------------------------------------------------
    def CopyInitialPosition(self, xyz):
        """xyz: [n, 3] positions of this rank's particles.  Each is localised by the rank whose part
        of the bisection contains the point (its picpart holds the tet, as owned or ghost)."""
        t = self.torch
        xyz = xyz.reshape(-1, 3)
        dest = t.as_tensor(rcb_locate(self.tree, xyz.cpu().numpy()), dtype=t.int64, device=self.dev)
        idx = t.arange(self.n, device=self.dev)
        home = t.full_like(idx, self.rank)
        r_xyz, r_idx, r_home = self._exchange(dest, [xyz, idx, home])
        pos, lelem = self.walker.localise(r_xyz)
        b_pos, b_idx, b_gelem = self._exchange(r_home, [pos, r_idx, self.global_of_local[lelem]])
        self.pos[b_idx], self.gelem[b_idx] = b_pos, b_gelem

    def MoveToNextLocation(self, origin, dest, flying, weights):
        """One transport step for this rank's particles (tensors [n,3], [n,3], int8 [n], [n]).
        flying is zeroed on return, as in the reference (PumiTallyImpl.cpp:169-172)."""
        t = self.torch
        fly = t.nonzero(flying == 1).squeeze(1)
        home = t.full_like(fly, self.rank)
        org_f, pos_f = origin.reshape(-1, 3)[fly], self.pos[fly]
        # Records of the flying particles go to the owner of their parent tet.  Re-sourced particles
        # (origin differs from where the particle is: phase 1 of the reference, PumiTallyImpl.cpp:71-112,
        # would walk them there with tallying off) go to the rank whose part holds the origin instead,
        # which localises them inside its picpart; marked by parent tet -1.
        moved = (org_f != pos_f).any(1)
        to_rank = self.owner_of_tet[self.gelem[fly]]
        gel_f = self.gelem[fly].clone()
        if bool(moved.any()):
            where = t.as_tensor(rcb_locate(self.tree, org_f[moved].cpu().numpy()), dtype=t.int64, device=self.dev)
            to_rank[moved] = where
            gel_f[moved] = -1
        t0 = self._tick(None, 0.0)
        rec = self._exchange(to_rank, [pos_f, org_f, dest.reshape(-1, 3)[fly], weights[fly], gel_f, fly, home])
        t0 = self._tick("route_out", t0)
        self.stats_routed += len(fly)
        done = []
        rounds = 0
        while True:
            pos, org, dst, w, gel, idx, hom = rec
            fresh = gel < 0
            lelem = self.local_of_global[gel.clamp(min=0)]
            if bool(fresh.any()):  # start the relocation walk from this picpart's parking position
                pos = pos.clone()
                pos[fresh] = self.walker._park if t.is_tensor(self.walker._park) else t.as_tensor(self.walker._park)
                lelem = t.where(fresh, t.zeros_like(lelem), lelem)
            assert bool((lelem >= 0).all()), "a routed particle's tet is not in this picpart"
            t0 = self._tick("prepare", t0)
            new_pos, new_lelem = self.walker.walk(pos, lelem, org, dst, w)
            t0 = self._tick("walk", t0)
            # where did each walk end?  reached its destination, on the true hull, or on the picpart boundary
            stopped = (new_pos != dst).any(1)
            cand = t.nonzero(stopped).squeeze(1)
            go_on = t.zeros(len(new_pos), dtype=t.bool, device=self.dev)
            nxt = t.zeros(len(new_pos), dtype=t.int64, device=self.dev)
            if len(cand):
                pl = self.planes[new_lelem[cand]]                                   # [m, 4, 4]
                dist_f = ((pl[:, :, :3] * new_pos[cand, None, :]).sum(2) - pl[:, :, 3]).abs()
                fng = self.face_next_global[new_lelem[cand]]                        # [m, 4]
                dist_f = t.where(fng == -2, t.full_like(dist_f, float("inf")), dist_f)  # interior faces are no exits
                f_exit = dist_f.argmin(1)
                beyond = fng.gather(1, f_exit[:, None]).squeeze(1)
                on_boundary = dist_f.gather(1, f_exit[:, None]).squeeze(1) < 1e-9 * (1.0 + new_pos[cand].abs().amax(1))
                handoff = (beyond >= 0) & on_boundary
                go_on[cand[handoff]] = True
                nxt[cand[handoff]] = beyond[handoff]
            fin = ~go_on
            done.append((hom[fin], idx[fin], new_pos[fin], self.global_of_local[new_lelem[fin]]))
            n_on = int(go_on.sum())
            rounds += 1
            t0 = self._tick("classify", t0)
            if self._global_sum(n_on) == 0:
                break
            self.stats_handoffs += n_on
            # the rest of the track starts at the crossing point, in the tet beyond the boundary
            rec = self._exchange(self.owner_of_tet[nxt[go_on]],
                                 [new_pos[go_on], new_pos[go_on], dst[go_on], w[go_on], nxt[go_on], idx[go_on], hom[go_on]])
            t0 = self._tick("handoff", t0)
        self.stats_rounds += rounds
        hom = t.cat([d[0] for d in done]); idx = t.cat([d[1] for d in done])
        pos = t.cat([d[2] for d in done]); gel = t.cat([d[3] for d in done])
        b_idx, b_pos, b_gel = self._exchange(hom, [idx, pos, gel])
        self.pos[b_idx], self.gelem[b_idx] = b_pos, b_gel
        flying.zero_()
        self._tick("route_back", t0)

    # ---- batch end -----------------------------------------------------------------------------
    def exchange_ghost_tallies(self):
        """Adds what this rank tallied in its ghost tets into the owners' copies; returns this rank's
        owned flux (global ids ``owned_global``).  Only ghost-layer values travel."""
        t = self.torch
        flux = self.walker.flux()
        n_owned = self.pic.n_owned
        ghost_global = self.global_of_local[n_owned:]
        vals, gids = self._exchange(self.owner_of_tet[ghost_global], [flux[n_owned:], ghost_global])
        owned = flux[:n_owned].clone()
        owned.index_add_(0, self.local_of_global[gids], vals)
        self.ghost_values_sent = int(len(ghost_global))
        return owned

    def global_flux(self):
        """The whole mesh's flux on every rank (owned pieces all-gathered), for output and parity checks."""
        t, dist = self.torch, self.dist
        owned = self.exchange_ghost_tallies()
        out = t.zeros(self.num_elements, dtype=t.float64, device=self.dev)
        out[self.global_of_local[: self.pic.n_owned]] = owned
        dist.all_reduce(out)
        return out

    @property
    def elem_ids(self):
        return self.gelem

    @property
    def positions(self):
        return self.pos
