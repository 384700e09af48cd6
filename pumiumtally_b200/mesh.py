"""Synthetic tetrahedral meshes for the tally engine (host side, numpy).

These generators stand in for the Omega_h ``.osh`` meshes the reference's
constructor reads (reference: src/pumitally/PumiTallyImpl.cpp:553-568).  They
produce plain ``coords float64[V,3]`` / ``tet2vert int32[E,4]`` arrays that are
handed both to the CUDA library (``pumitally_create_from_arrays``) and to the
CPU oracle, so both sides see bit-identical input.

The 1x1x1 Kuhn box reproduces the fixture of the reference's known-answer test
(reference: test/test_pumi_tally_impl_methods.cpp:34-35 --
``Omega_h::build_box(world, OMEGA_H_SIMPLEX, 1,1,1, 1,1,1)``): six tets around
the (0,0,0)-(1,1,1) diagonal with element 0 = {y>=x>=z} (centroid
(0.5,0.75,0.25), test line 83), 2 = {z>=y>=x}, 3 = {z>=x>=y}, 4 = {x>=z>=y}.
"""
from __future__ import annotations

import os

import numpy as np

# Hex-local corner numbering (x fastest): 0:(0,0,0) 1:(1,0,0) 2:(1,1,0) 3:(0,1,0)
#                                         4:(0,0,1) 5:(1,0,1) 6:(1,1,1) 7:(0,1,1)
_HEX_CORNER = np.array(
    [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]],
    dtype=np.int64,
)
# Six tets sharing the 0-6 diagonal, in cyclic order around it; element-local id
# k uses entry (k+1) % 6 so that id 0 is {y>=x>=z} as the reference test pins.
_KUHN_CYCLE = np.array(
    [[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]],
    dtype=np.int64,
)
KUHN_TETS = _KUHN_CYCLE[(np.arange(6) + 1) % 6]


def kuhn_box(nx: int, ny: int, nz: int, lx: float = None, ly: float = None, lz: float = None):
    """Kuhn (Freudenthal) split of the box [0,lx]x[0,ly]x[0,lz] into nx*ny*nz*6 tets.

    Returns (coords float64[V,3], tet2vert int32[E,4]).  Element id =
    6*cell + k with cell = (kz*ny + jy)*nx + ix.  Default cell size is 1.
    """
    lx = float(nx) if lx is None else float(lx)
    ly = float(ny) if ly is None else float(ly)
    lz = float(nz) if lz is None else float(lz)
    xs = np.arange(nx + 1, dtype=np.float64) * (lx / nx)
    ys = np.arange(ny + 1, dtype=np.float64) * (ly / ny)
    zs = np.arange(nz + 1, dtype=np.float64) * (lz / nz)
    # make the far faces land exactly on lx/ly/lz
    xs[-1], ys[-1], zs[-1] = lx, ly, lz
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    kz, jy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ix, jy, kz = ix.ravel(), jy.ravel(), kz.ravel()

    def vid(i, j, k):
        return (k * (ny + 1) + j) * (nx + 1) + i

    corner = np.stack(
        [vid(ix + c[0], jy + c[1], kz + c[2]) for c in _HEX_CORNER], axis=1
    )  # [cells, 8]
    tets = corner[:, KUHN_TETS]  # [cells, 6, 4]
    tet2vert = tets.reshape(-1, 4).astype(np.int32)
    return np.ascontiguousarray(coords), np.ascontiguousarray(tet2vert)


def jitter_interior(coords, tet2vert, amplitude: float, seed: int = 7):
    """Move interior vertices of a box mesh by uniform noise (keeps the hull
    planar and convex).  ``amplitude`` is a fraction of the smallest cell edge;
    keep it below ~0.2 so no tet inverts."""
    lo, hi = coords.min(0), coords.max(0)
    on_hull = np.any((coords <= lo + 1e-12) | (coords >= hi - 1e-12), axis=1)
    e = coords[tet2vert[:, 1:]] - coords[tet2vert[:, :1]]
    h = np.abs(e).max(axis=(1, 2)).min()
    rng = np.random.default_rng(seed)
    out = coords.copy()
    out[~on_hull] += rng.uniform(-amplitude * h, amplitude * h, size=(int((~on_hull).sum()), 3))
    return out, tet2vert


def delaunay_box(n_points: int, seed: int = 7, renumber: bool = True):
    """Delaunay tetrahedralisation of random points in the unit cube plus its 8
    corners (convex hull = the cube).  Slivers with volume < 1e-12 are dropped
    only if that keeps the mesh a convex, face-connected complex; by
    construction (general-position random points) there are none in practice.
    """
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    corners = np.array([[i, j, k] for k in (0, 1) for j in (0, 1) for i in (0, 1)], dtype=np.float64)
    pts = np.vstack([corners, rng.uniform(0.02, 0.98, size=(n_points, 3))])
    tri = Delaunay(pts)
    t2v = tri.simplices.astype(np.int32)
    if renumber:
        t2v = t2v[rng.permutation(len(t2v))]
    return np.ascontiguousarray(pts), np.ascontiguousarray(t2v)


def tet_volumes(coords, tet2vert):
    """Unsigned tet volumes (reference: PumiTallyImpl.cpp:393-402 uses
    Omega_h::simplex_size_from_basis, i.e. det(basis)/6)."""
    v = coords[tet2vert]
    b = v[:, 1:] - v[:, :1]
    return np.abs(np.linalg.det(b)) / 6.0


def save_raw_mesh(path: str, coords, tet2vert) -> None:
    """Write the library's raw mesh file: magic 'PUMITB2\\0', int64 nverts,
    int64 ntets, float64 coords[V*3], int32 tet2vert[E*4] (little endian)."""
    coords = np.ascontiguousarray(coords, dtype="<f8")
    tet2vert = np.ascontiguousarray(tet2vert, dtype="<i4")
    with open(path, "wb") as f:
        f.write(b"PUMITB2\0")
        f.write(np.array([coords.shape[0], tet2vert.shape[0]], dtype="<i8").tobytes())
        f.write(coords.tobytes())
        f.write(tet2vert.tobytes())


def save_gmsh(path: str, coords, tet2vert, version: str = "2.2", node_id_offset: int = 1, binary: bool = False) -> None:
    """Write a Gmsh file (format 2.2 or 4.1, ASCII or binary) holding the nodes and the tets as element
    type 4, plus one boundary triangle and one point element that readers must skip."""
    coords = np.asarray(coords, dtype=np.float64)
    t2v = np.asarray(tet2vert, dtype=np.int64) + node_id_offset
    nv, nt = len(coords), len(t2v)
    if binary:
        import struct

        with open(path, "wb") as f:
            if version.startswith("2"):
                f.write(b"$MeshFormat\n2.2 1 8\n" + struct.pack("<i", 1) + b"\n$EndMeshFormat\n$Nodes\n%d\n" % nv)
                for i, xyz in enumerate(coords):
                    f.write(struct.pack("<i3d", i + node_id_offset, *xyz))
                f.write(b"\n$EndNodes\n$Elements\n%d\n" % (nt + 2))
                f.write(struct.pack("<3i", 15, 1, 2) + struct.pack("<4i", 1, 0, 1, int(t2v[0, 0])))
                f.write(struct.pack("<3i", 2, 1, 2) + struct.pack("<6i", 2, 0, 1, *(int(v) for v in t2v[0, :3])))
                f.write(struct.pack("<3i", 4, nt, 2))
                for e, t in enumerate(t2v):
                    f.write(struct.pack("<7i", e + 3, 0, 1, *(int(v) for v in t)))
                f.write(b"\n$EndElements\n")
            else:
                f.write(b"$MeshFormat\n4.1 1 8\n" + struct.pack("<i", 1) + b"\n$EndMeshFormat\n$Nodes\n")
                f.write(struct.pack("<4Q", 1, nv, node_id_offset, nv + node_id_offset - 1))
                f.write(struct.pack("<3iQ", 3, 1, 0, nv))
                f.write(np.arange(node_id_offset, nv + node_id_offset, dtype=np.uint64).tobytes())
                f.write(coords.tobytes())
                f.write(b"\n$EndNodes\n$Elements\n")
                f.write(struct.pack("<4Q", 2, nt + 1, 1, nt + 1))
                f.write(struct.pack("<3iQ", 2, 1, 2, 1) + struct.pack("<4Q", 1, *(int(v) for v in t2v[0, :3])))
                f.write(struct.pack("<3iQ", 3, 1, 4, nt))
                rec = np.empty((nt, 5), dtype=np.uint64)
                rec[:, 0] = np.arange(2, nt + 2)
                rec[:, 1:] = t2v
                f.write(rec.tobytes())
                f.write(b"\n$EndElements\n")
        return
    with open(path, "w") as f:
        if version.startswith("2"):
            f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % nv)
            for i, (x, y, z) in enumerate(coords):
                f.write("%d %.17g %.17g %.17g\n" % (i + node_id_offset, x, y, z))
            f.write("$EndNodes\n$Elements\n%d\n" % (nt + 2))
            f.write("1 15 2 0 1 %d\n" % t2v[0, 0])
            f.write("2 2 2 0 1 %d %d %d\n" % tuple(t2v[0, :3]))
            for e, t in enumerate(t2v):
                f.write("%d 4 2 0 1 %d %d %d %d\n" % (e + 3, *t))
            f.write("$EndElements\n")
        else:
            f.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
            f.write("$Nodes\n1 %d %d %d\n3 1 0 %d\n" % (nv, node_id_offset, nv + node_id_offset - 1, nv))
            for i in range(nv):
                f.write("%d\n" % (i + node_id_offset))
            for x, y, z in coords:
                f.write("%.17g %.17g %.17g\n" % (x, y, z))
            f.write("$EndNodes\n$Elements\n2 %d 1 %d\n" % (nt + 1, nt + 1))
            f.write("2 1 2 1\n1 %d %d %d\n" % tuple(t2v[0, :3]))
            f.write("3 1 4 %d\n" % nt)
            for e, t in enumerate(t2v):
                f.write("%d %d %d %d %d\n" % (e + 2, *t))
            f.write("$EndElements\n")


def simplex_downward(tet2vert):
    """Edges, triangles and the downward adjacencies tet->tri, tri->edge, edge->vert of a tet
    mesh (entities numbered in order of first appearance of their sorted vertex tuple)."""
    t = np.asarray(tet2vert, dtype=np.int64)
    tri_of_tet = np.stack([t[:, [0, 2, 1]], t[:, [0, 1, 3]], t[:, [1, 2, 3]], t[:, [2, 0, 3]]], axis=1)
    tri_keys = np.sort(tri_of_tet.reshape(-1, 3), axis=1)
    tris, rf2f = np.unique(tri_keys, axis=0, return_inverse=True)
    edge_of_tri = np.stack([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [0, 2]]], axis=1)
    edge_keys = np.sort(edge_of_tri.reshape(-1, 2), axis=1)
    edges, fe2e = np.unique(edge_keys, axis=0, return_inverse=True)
    return (edges.astype(np.int32), fe2e.reshape(-1).astype(np.int32), rf2f.reshape(-1).astype(np.int32))


def save_osh(path: str, coords, tet2vert, version: int = 10, compressed: bool = True,
             tag_layout: str = "direct", extra_tags: bool = True, bare_stream: bool = False,
             version_in_stream: bool = False, family_byte: bool = None) -> None:
    """Write an Omega_h-style binary mesh directory (``nparts``, ``version``, ``0.osh``) following
    the stream layout documented in ``csrc/osh_reader.cpp``.  The layout is a restatement from
    memory of Omega_h's published format (no Omega_h here to check against): this writer exists so
    the reader's parsing, zlib handling and tet->tri->edge->vert composition are exercised.
    Alignment codes are written as zeros (the reader does not consume them).

    tag_layout: "direct" (name, ncomps, type, array), "class_ids" (an i32 class-id count and
    optional id list before the array) or "flags" (two flag bytes, stream versions < 5).
    version_in_stream / family_byte force the two header details the reader is unsure about."""
    import struct
    import zlib

    coords = np.ascontiguousarray(coords, dtype=np.float64)
    ev2v, fe2e, rf2f = simplex_downward(tet2vert)
    nv = len(coords)

    def arr(a):
        a = np.ascontiguousarray(a)
        raw = a.tobytes()
        head = struct.pack("<i", a.size)
        if compressed:
            z = zlib.compress(raw, 1)
            return head + struct.pack("<q", len(z)) + z
        return head + raw

    def tag(name, ncomps, type_code, a, class_ids=None):
        b = struct.pack("<i", len(name)) + name.encode() + struct.pack("<bb", ncomps, type_code)
        if tag_layout == "class_ids":
            ids = np.asarray(class_ids if class_ids is not None else [], dtype=np.int32)
            b += struct.pack("<i", ids.size)
            if ids.size:
                b += arr(ids)
        elif tag_layout == "flags":
            b += struct.pack("<bb", 1, 1)
        return b + arr(a)

    s = bytes([0xA1, 0x1A])
    if bare_stream or version_in_stream:
        s += struct.pack("<i", version)
    s += struct.pack("<b", 1 if compressed else 0)
    if (version >= 7) if family_byte is None else family_byte:
        s += struct.pack("<b", 0)                       # family: simplex
    s += struct.pack("<biibib", 3, 1, 0, 0, 0, 0)       # dim, comm size/rank, parting, ghost layers, no hints
    s += struct.pack("<i", nv)
    s += arr(ev2v.reshape(-1))
    s += arr(fe2e) + arr(np.zeros(fe2e.size, dtype=np.int8))
    s += arr(rf2f) + arr(np.zeros(rf2f.size, dtype=np.int8))
    vtags = []
    if extra_tags:
        vtags.append(tag("class_dim", 1, 0, np.full(nv, 3, dtype=np.int8)))
        vtags.append(tag("global", 1, 3, np.arange(nv, dtype=np.int64), class_ids=[1, 2, 3]))
    vtags.append(tag("coordinates", 3, 5, coords.reshape(-1)))
    if extra_tags:
        vtags.append(tag("class_id", 1, 2, np.zeros(nv, dtype=np.int32)))
    s += struct.pack("<i", len(vtags)) + b"".join(vtags)
    for n in (len(ev2v), fe2e.size // 3, rf2f.size // 4):   # tags of edges, triangles, tets
        s += struct.pack("<i", 1) + tag("class_dim", 1, 0, np.full(n, 3, dtype=np.int8))
    if bare_stream:
        with open(path, "wb") as f:
            f.write(s)
        return
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "nparts"), "w") as f:
        f.write("1\n")
    with open(os.path.join(path, "version"), "w") as f:
        f.write("%d\n" % version)
    with open(os.path.join(path, "0.osh"), "wb") as f:
        f.write(s)
