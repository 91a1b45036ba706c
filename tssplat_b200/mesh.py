"""Tet-sphere mesh helpers: Vega ``.veg`` reader/writer and the seeded synthetic tet-sphere packs
the parity tests and ``bench.py`` run on.

Nothing here touches the GPU.  A *pack* is N tet-spheres concatenated into one mesh exactly the way
the reference does it (vertex arrays stacked, tet indices offset by the running vertex base:
``geometry/tetmesh_geometry.py:305-331``), so spheres share no vertices and every operator is
block-diagonal by sphere.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

__all__ = ["TetPack", "load_veg", "save_veg", "make_tet_sphere", "make_pack", "concat_spheres",
           "perturb", "mean_edge_length", "connected_components", "surface_vf", "surface_vf_gpu", "save_npy_spheres",
           "load_npy_spheres"]

_EDGES = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]])


@dataclass
class TetPack:
    """Concatenated tet-spheres.  ``verts`` fp32 [n,3] rest positions, ``tets`` int32 [nele,4]."""
    verts: np.ndarray
    tets: np.ndarray
    vert_offsets: np.ndarray      # [S+1] first vertex of each sphere
    tet_offsets: np.ndarray       # [S+1] first tet of each sphere

    @property
    def num_spheres(self) -> int:
        return len(self.vert_offsets) - 1

    @property
    def n(self) -> int:
        return int(self.verts.shape[0])

    @property
    def nele(self) -> int:
        return int(self.tets.shape[0])

    def algorithmic_bytes(self) -> int:
        """B_alg = 24 V + 68 T summed over spheres."""
        return 24 * self.n + 68 * self.nele

    def slice_spheres(self, lo: int, hi: int) -> "TetPack":
        """Spheres [lo, hi) as a self-contained pack (indices rebased) -- the per-rank shard."""
        v0, v1 = int(self.vert_offsets[lo]), int(self.vert_offsets[hi])
        t0, t1 = int(self.tet_offsets[lo]), int(self.tet_offsets[hi])
        return TetPack(self.verts[v0:v1].copy(), (self.tets[t0:t1] - v0).astype(np.int32),
                       self.vert_offsets[lo:hi + 1] - v0, self.tet_offsets[lo:hi + 1] - t0)


def load_veg(path: str):
    """Read a Vega ``.veg`` tet mesh (1-indexed ids; layout as ``tssplat_ext/a.veg:1-8,4507-4512``).

    Returns (verts float64 [n,3], tets int32 [nele,4], 0-based).
    """
    with open(path, "r") as f:
        lines = [ln.strip() for ln in f]
    i, n_lines = 0, len(lines)
    verts = tets = None
    while i < n_lines:
        ln = lines[i]
        if ln.startswith("*VERTICES"):
            nv, dim = (int(s) for s in lines[i + 1].split()[:2])
            if dim != 3:
                raise ValueError(".veg: only 3-D vertices supported")
            block = np.array([lines[i + 2 + j].split() for j in range(nv)], dtype=np.float64)
            first = int(block[0, 0])
            verts = np.empty((nv, 3))
            verts[block[:, 0].astype(np.int64) - first] = block[:, 1:4]
            base = first
            i += 2 + nv
        elif ln.startswith("*ELEMENTS"):
            if not lines[i + 1].upper().startswith("TET"):
                raise ValueError(".veg: only TET elements supported")
            ne, per = (int(s) for s in lines[i + 2].split()[:2])
            if per != 4:
                raise ValueError(".veg: tets need 4 vertices")
            block = np.array([lines[i + 3 + j].split() for j in range(ne)], dtype=np.int64)
            tets = block[:, 1:5]
            i += 3 + ne
        else:
            i += 1
    if verts is None or tets is None:
        raise ValueError(".veg: missing *VERTICES or *ELEMENTS")
    return verts, (tets - base).astype(np.int32)


def save_veg(path: str, verts: np.ndarray, tets: np.ndarray,
             density: float = 1000.0, E: float = 1e9, nu: float = 0.45) -> None:
    """Write the same ``.veg`` dialect the reference exports (``tssplat_ext/a.veg:26631-26636``)."""
    verts = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    tets = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    with open(path, "w") as f:
        f.write("# Vega mesh file.\n# %d vertices, %d elements\n\n" % (len(verts), len(tets)))
        f.write("*VERTICES\n%d 3 0 0\n" % len(verts))
        for i, p in enumerate(verts):
            f.write("%d %.15g %.15g %.15g\n" % (i + 1, p[0], p[1], p[2]))
        f.write("\n*ELEMENTS\nTET\n%d 4 0\n" % len(tets))
        for i, t in enumerate(tets):
            f.write("%d %d %d %d %d\n" % (i + 1, t[0] + 1, t[1] + 1, t[2] + 1, t[3] + 1))
        f.write("\n*MATERIAL defaultMaterial\nENU, %g, %g, %g\n\n*REGION\nallElements, defaultMaterial\n"
                % (density, E, nu))


def _signed_volumes(P: np.ndarray, tets: np.ndarray) -> np.ndarray:
    a = P[tets[:, 1]] - P[tets[:, 0]]
    b = P[tets[:, 2]] - P[tets[:, 0]]
    c = P[tets[:, 3]] - P[tets[:, 0]]
    return np.einsum("ij,ij->i", np.cross(a, b), c) / 6.0


def make_tet_sphere(seed: int, n_tets: int = 4096, n_points: Optional[int] = None):
    """One seeded unit tet-sphere with exactly ``n_tets`` positively oriented tets.

    Fibonacci shells + jittered interior points, Delaunay, sliver removal, then the worst-shaped
    tets are trimmed until ``n_tets`` remain; unreferenced vertices are dropped.
    Returns (verts float64 [V,3], tets int32 [n_tets,4]).
    """
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    if n_points is None:
        n_points = max(16, int(round(n_tets / 4.9)))      # a.veg: 22120 tets / 4500 verts
    while True:
        # shells: radii so that points are roughly uniform in the ball
        n_shell = max(2, int(round((n_points / 4.2) ** (1.0 / 3.0))))
        radii = (np.arange(1, n_shell + 1) / n_shell)
        w = radii ** 2
        counts = np.maximum(4, np.round(w / w.sum() * (n_points - 1)).astype(int))
        pts = [np.zeros((1, 3))]
        for r, m in zip(radii, counts):
            k = np.arange(m) + 0.5
            phi = np.arccos(1.0 - 2.0 * k / m)
            th = np.pi * (1.0 + 5.0 ** 0.5) * k + rng.uniform(0, 2 * np.pi)
            p = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1) * r
            if r < 1.0:
                p += rng.normal(0.0, 0.12 / n_shell, p.shape)
            pts.append(p)
        P = np.concatenate(pts, 0)
        tets = Delaunay(P).simplices.astype(np.int64)
        vol = _signed_volumes(P, tets)
        flip = vol < 0
        tets[flip] = tets[flip][:, [0, 1, 3, 2]]
        vol = np.abs(vol)
        Pe = P[tets]
        el = np.linalg.norm(Pe[:, _EDGES[:, 0]] - Pe[:, _EDGES[:, 1]], axis=2)
        quality = vol / (el.max(axis=1) ** 3 + 1e-300)      # scale-free shape measure
        keep = (vol > 1e-7) & (quality > 2e-3)
        tets, quality = tets[keep], quality[keep]
        if len(tets) >= n_tets:
            break
        n_points = int(n_points * 1.08) + 8
    order = np.argsort(-quality, kind="stable")[:n_tets]
    tets = tets[np.sort(order)]
    used = np.unique(tets)
    remap = np.full(P.shape[0], -1, dtype=np.int64)
    remap[used] = np.arange(used.size)
    return P[used], remap[tets].astype(np.int32)


def concat_spheres(spheres: Sequence) -> TetPack:
    """Concatenate (verts, tets) pairs the way ``geometry/tetmesh_geometry.py:310-331`` does."""
    vs, ts, vo, to = [], [], [0], [0]
    for v, t in spheres:
        ts.append(np.asarray(t, dtype=np.int64) + vo[-1])
        vs.append(np.asarray(v, dtype=np.float32))
        vo.append(vo[-1] + len(v))
        to.append(to[-1] + len(t))
    return TetPack(np.concatenate(vs, 0).astype(np.float32), np.concatenate(ts, 0).astype(np.int32),
                   np.asarray(vo, dtype=np.int64), np.asarray(to, dtype=np.int64))


def make_pack(num_spheres: int, n_tets: int = 4096, seed: int = 0,
              unique: Optional[int] = None) -> TetPack:
    """``num_spheres`` tet-spheres (sphere i: ``seed = 1000 + seed + i``), each scaled by a radius
    U(0.1,0.3), randomly rotated and translated by U(-0.7,0.7)^3.  ``unique`` bounds how many
    distinct Delaunay meshes are generated (the rest are re-posed copies) to keep setup fast.
    """
    rng = np.random.default_rng(7919 + seed)
    n_unique = num_spheres if unique is None else min(unique, num_spheres)
    templates = [make_tet_sphere(1000 + seed + i, n_tets) for i in range(n_unique)]
    spheres = []
    for i in range(num_spheres):
        v, t = templates[i % n_unique]
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        r = rng.uniform(0.1, 0.3)
        c = rng.uniform(-0.7, 0.7, 3)
        spheres.append(((v @ q.T) * r + c, t))
    return concat_spheres(spheres)


def mean_edge_length(verts: np.ndarray, tets: np.ndarray) -> float:
    P = np.asarray(verts, dtype=np.float64)[np.asarray(tets, dtype=np.int64)]
    return float(np.linalg.norm(P[:, _EDGES[:, 0]] - P[:, _EDGES[:, 1]], axis=2).mean())


def perturb(pack_or_verts, tets=None, sigma_rel: float = 0.02, seed: int = 0) -> np.ndarray:
    """x = X + N(0, (sigma_rel * h)^2) per sphere, h = that sphere's mean rest edge length.

    ``sigma_rel=0.02`` is the benign case, ``0.35`` the inverted one.
    Returns fp32 [n,3].
    """
    rng = np.random.default_rng(seed)
    if isinstance(pack_or_verts, TetPack):
        pk = pack_or_verts
        x = pk.verts.astype(np.float64).copy()
        for s in range(pk.num_spheres):
            v0, v1 = pk.vert_offsets[s], pk.vert_offsets[s + 1]
            t0, t1 = pk.tet_offsets[s], pk.tet_offsets[s + 1]
            h = mean_edge_length(pk.verts, pk.tets[t0:t1])
            x[v0:v1] += rng.normal(0.0, sigma_rel * h, (v1 - v0, 3))
        return x.astype(np.float32)
    X = np.asarray(pack_or_verts, dtype=np.float64)
    h = mean_edge_length(X, tets)
    return (X + rng.normal(0.0, sigma_rel * h, X.shape)).astype(np.float32)


def connected_components(n: int, tets: np.ndarray) -> np.ndarray:
    """Vertex -> component label (0..C-1, in order of first vertex) via the tet connectivity."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components as cc
    tets = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    r = np.repeat(tets[:, 0], 3)
    c = tets[:, 1:].reshape(-1)
    g = sp.coo_matrix((np.ones(r.size, dtype=np.int8), (r, c)), shape=(n, n))
    _, lab = cc(g, directed=False)
    first = np.full(lab.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, lab, np.arange(n))
    rank = np.argsort(np.argsort(first))
    return rank[lab]


def surface_vf(tets: np.ndarray):
    """Surface of a tet mesh without pypgo: (surface_vertices, surface_triangles).

    Same outputs, in the same order and with the same orientation, as the reference's
    ``get_surface_vf`` (``geometry/mesh_utils.py:5-35``): the faces that belong to exactly one tet, listed
    in lexicographic order of their sorted vertex triple, each written in the orientation its tet gives it
    (face k opposite local vertex k: (1,2,3), (0,3,2), (0,1,3), (0,2,1)) and re-indexed into the sorted
    list of surface vertex ids.  ``reset()`` / ``permute_surface_v()`` of the geometry re-run this
    (``geometry/tetmesh_geometry.py:164-170,369-371``).
    """
    tets = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    T = len(tets)
    corner = np.array([[1, 2, 3], [0, 3, 2], [0, 1, 3], [0, 2, 1]])
    oriented = tets[:, corner].transpose(1, 0, 2).reshape(4 * T, 3)      # block k = faces opposite vertex k
    key = np.sort(oriented, axis=1)
    order = np.lexsort((np.arange(4 * T), key[:, 2], key[:, 1], key[:, 0]))   # ties: first occurrence first
    ks = key[order]
    new = np.ones(4 * T, dtype=bool)
    new[1:] = np.any(ks[1:] != ks[:-1], axis=1)
    start = np.flatnonzero(new)
    count = np.diff(np.append(start, 4 * T))
    once = start[count == 1]
    faces = oriented[order[once]]
    verts = np.unique(faces)
    return verts, np.searchsorted(verts, faces)


def surface_vf_gpu(tets: np.ndarray, n_vertices: Optional[int] = None, device: int = 0):
    """``surface_vf`` on the GPU (``tsb_surface_extract``: two radix sorts, a select and a scan): the same
    ``(surface_vertices, surface_triangles)`` as the reference's ``get_surface_vf``
    (``geometry/mesh_utils.py:5-35``), for ``reset()`` / ``permute_surface_v()``-style re-runs on large packs.
    Needs a CUDA device (no CPU path: use ``surface_vf`` for the numpy restatement)."""
    import ctypes as C

    from . import _capi
    t = np.ascontiguousarray(np.asarray(tets).reshape(-1, 4), dtype=np.int32)
    n = int(n_vertices) if n_vertices is not None else (int(t.max()) + 1 if t.size else 0)
    nsv, nsf = C.c_int32(0), C.c_int32(0)
    pv, pf = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
    rc = _capi.lib.tsb_surface_extract(t.ctypes.data, len(t), n, int(device), C.byref(nsv), C.byref(nsf), C.byref(pv), C.byref(pf))
    if rc:
        raise RuntimeError(f"surface_vf_gpu: {(_capi.lib.tsb_setup_last_error() or b'').decode()} (code {rc})")
    try:
        verts = np.ctypeslib.as_array(pv, shape=(nsv.value,)).astype(np.int64) if nsv.value else np.zeros(0, np.int64)
        faces = np.ctypeslib.as_array(pf, shape=(nsf.value, 3)).astype(np.int64) if nsf.value else np.zeros((0, 3), np.int64)
    finally:
        _capi.lib.tsb_free_host(pv)
        _capi.lib.tsb_free_host(pf)
    return verts, faces


def save_npy_spheres(pack: "TetPack", path: str, filename: str, verts: Optional[np.ndarray] = None) -> List[str]:
    """The reference's array export without pypgo: ``<filename>_vtx.npy`` / ``_elem.npy`` of the whole mesh
    (``geometry/tetrahedron_mesh.py:82-91``) and, per sphere i, ``<filename>_sp{i}_vtx.npy`` (its vertices) and
    ``<filename>_sp{i}_elem.npy`` (its tets with sphere-local indices) as
    ``geometry/tetmesh_geometry.py:373-382`` writes them.  ``verts`` overrides the positions (e.g. the optimised
    ``tet_v``).  Returns the written paths."""
    import os
    os.makedirs(path, exist_ok=True)
    V = pack.verts if verts is None else np.asarray(verts).reshape(pack.verts.shape)
    out = []

    def put(name, arr):
        f = os.path.join(path, filename + name)
        np.save(f, arr)
        out.append(f)
    put("_vtx.npy", V)
    put("_elem.npy", pack.tets)
    for i in range(pack.num_spheres):
        v0, v1 = int(pack.vert_offsets[i]), int(pack.vert_offsets[i + 1])
        t0, t1 = int(pack.tet_offsets[i]), int(pack.tet_offsets[i + 1])
        put(f"_sp{i}_vtx.npy", V[v0:v1])
        put(f"_sp{i}_elem.npy", (pack.tets[t0:t1] - v0).astype(np.int32))
    return out


def load_npy_spheres(path: str, filename: str) -> "TetPack":
    """Read back ``<filename>_sp{i}_vtx.npy`` / ``_sp{i}_elem.npy`` (any number of spheres) into a pack,
    concatenated the way ``geometry/tetmesh_geometry.py:305-331`` does."""
    import os
    spheres = []
    i = 0
    while os.path.exists(os.path.join(path, f"{filename}_sp{i}_vtx.npy")):
        v = np.load(os.path.join(path, f"{filename}_sp{i}_vtx.npy"))
        t = np.load(os.path.join(path, f"{filename}_sp{i}_elem.npy"))
        spheres.append((v, np.asarray(t).reshape(-1, 4)))
        i += 1
    if not spheres:
        raise FileNotFoundError(f"no {filename}_sp*_vtx.npy under {path}")
    return concat_spheres(spheres)
