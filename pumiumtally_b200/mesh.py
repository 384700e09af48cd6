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


def save_gmsh(path: str, coords, tet2vert, version: str = "2.2", node_id_offset: int = 1) -> None:
    """Write an ASCII Gmsh file (format 2.2 or 4.1) holding the nodes and the tets as element
    type 4, plus one boundary triangle and one point element that readers must skip."""
    coords = np.asarray(coords, dtype=np.float64)
    t2v = np.asarray(tet2vert, dtype=np.int64) + node_id_offset
    nv, nt = len(coords), len(t2v)
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
