"""CPU oracle for the tssplat geometry-energy hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(``tssplat_b200`` + ``tet_spheres``) never routes through it and fails loudly without its CUDA
library.

PARITY UNPINNED.  The reference ships no golden vectors, no known-answer tests and no fixtures with
expected values for this path, its extension cannot be built here (needs
libpgo: ``tssplat_ext/tet_spheres/tet_spheres.cpp:3``), and the weights of the tet Laplacian ``L``
live in the un-vendored, un-pinned third-party libpgo (github.com/bohanwang/libpgo, version not
pinned by the reference: ``README.md:38``; call site
``tssplat_ext/tet_spheres/tet_spheres.cpp:148``:
``pgo_create_tet_biharmonic_gradient_matrix(tetMeshGeo, 1, 0)``).  The working assumption, isolated
in :func:`tet_laplacian`, is: ``L`` = graph Laplacian of the tet face-adjacency graph (first
argument 1 = face neighbours), unscaled (second argument 0: ``L_tt = #face-neighbours``,
``L_ts = -1``), applied identically to each of the 9 entries of ``F``.

What this file restates, with the reference file:line each function follows
(all paths relative to /root/reference):

* ``F = G x`` -- per-tet deformation gradient ``F_t = Ds_t Dm_t^-1``; ``G`` is built exactly as
  ``geometry/mesh_utils.py:38-69`` (``compute_G_matrix``) builds it (row-major 3x3 flatten), and as
  ``tssplat_ext/tet_spheres/tet_spheres.cpp:149`` obtains it from libpgo.
* ``M = G^T L^T L G`` -- ``tet_spheres.cpp:148``.
* forward -- ``tssplat_ext/tet_spheres/tet_spheres_cuda.cu:118-195``:
  ``E = c1 * 0.5 * x^T (M x) + c2 * sum_t max(-det F_t, 0)^order``  (0.5 at ``:157``, c1/c2 at
  ``:191``; per-tet barrier ``cuda_forward_det`` ``:48-66``; ``det`` ``:21-30``).
* backward -- ``tet_spheres_cuda.cu:197-263``:
  ``dE/dx = gradH * ( c1 * M x + c2 * G^T D )``, ``D_t = -p (-J)^(p-1) cof(F_t)`` if ``J<0`` else 0
  (``cuda_backward_det`` ``:68-102``; cofactor ``ddetA_dA`` ``:32-46``; gradH scale ``:257-258``).
* ``order`` other than 2/4 gives zero energy and zero gradient in the reference (``:57-63,83-89``);
  the oracle reproduces that, the product ABI rejects it.

Two arithmetic modes:

* ``dtype=np.float64`` (default): operators and arithmetic in fp64 -- the parity target.
* ``dtype=np.float32``: operators built in fp64 then truncated to fp32 as the reference does
  (``tet_spheres.cpp:41-45``) and every product evaluated in fp32 -- shows the reference's own
  round-off floor (its ``x^T M x`` form cancels badly near the rest state).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = [
    "face_adjacency",
    "rest_inverse",
    "build_G",
    "tet_laplacian",
    "deformation_gradients",
    "ReferenceEnergyOracle",
]

# Face k of a tet is the face opposite to its local vertex k.
_FACE_OF = np.array([[1, 2, 3], [0, 3, 2], [0, 1, 3], [0, 2, 1]], dtype=np.int64)


def face_adjacency(tets: np.ndarray) -> np.ndarray:
    """nbr[t, k] = tet sharing the face of ``t`` opposite to local vertex ``k``, or -1.

    Raises ValueError on a face shared by more than two tets (non-manifold).
    """
    tets = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    T = tets.shape[0]
    faces = np.sort(tets[:, _FACE_OF].reshape(T * 4, 3), axis=1)
    order = np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))
    fs = faces[order]
    same = np.all(fs[1:] == fs[:-1], axis=1)
    if np.any(same[1:] & same[:-1]):
        raise ValueError("non-manifold tet mesh: a face is shared by more than two tets")
    nbr = np.full(T * 4, -1, dtype=np.int64)
    i = np.nonzero(same)[0]
    a, b = order[i], order[i + 1]
    nbr[a] = b // 4
    nbr[b] = a // 4
    return nbr.reshape(T, 4)


def rest_inverse(X: np.ndarray, tets: np.ndarray) -> np.ndarray:
    """B_t = Dm_t^-1 with Dm = [X1-X0, X2-X0, X3-X0] as columns (mesh_utils.py:50-54), fp64."""
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    tets = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    P = X[tets]                                    # T x 4 x 3
    Dm = np.transpose(P[:, 1:, :] - P[:, :1, :], (0, 2, 1))   # columns = edges
    return np.linalg.inv(Dm)


def deformation_gradients(x: np.ndarray, tets: np.ndarray, B: np.ndarray) -> np.ndarray:
    """F_t = Ds_t B_t  (T x 3 x 3)."""
    x = np.asarray(x).reshape(-1, 3)
    P = x[np.asarray(tets, dtype=np.int64).reshape(-1, 4)]
    Ds = np.transpose(P[:, 1:, :] - P[:, :1, :], (0, 2, 1))
    return Ds @ B


def build_G(X: np.ndarray, tets: np.ndarray) -> sp.csr_matrix:
    """Sparse (9T x 3n) gradient operator: vec_rowmajor(F_t) = (G x)[9t:9t+9].

    Same operator as ``geometry/mesh_utils.py:38-69`` (dense T x 9 x 12 there) scattered to global
    dof columns ``3*v + r``; 4 non-zeros per row.
    """
    tets = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    T = tets.shape[0]
    n = np.asarray(X).reshape(-1, 3).shape[0]
    B = rest_inverse(X, tets)                      # T x 3 x 3, rows = hat gradients of v1..v3
    a = np.concatenate([-B.sum(axis=1, keepdims=True), B], axis=1)   # T x 4 x 3 : a[t, k, c]
    # F[r, c] = sum_k x[v_k, r] * a[k, c]
    t_idx = np.arange(T)[:, None, None, None]
    r_idx = np.arange(3)[None, :, None, None]
    c_idx = np.arange(3)[None, None, :, None]
    k_idx = np.arange(4)[None, None, None, :]
    rows = np.broadcast_to(9 * t_idx + 3 * r_idx + c_idx, (T, 3, 3, 4))
    cols = np.broadcast_to(3 * tets[:, None, None, :] + r_idx, (T, 3, 3, 4))
    vals = np.broadcast_to(np.transpose(a, (0, 2, 1))[:, None, :, :], (T, 3, 3, 4))
    del k_idx
    return sp.csr_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(9 * T, 3 * n))


def tet_laplacian(tets: np.ndarray, face_neighbor: int = 1, scale: int = 0) -> sp.csr_matrix:
    """T x T Laplacian over tets -- THE ISOLATED libpgo ASSUMPTION (see module docstring).

    ``face_neighbor=1, scale=0`` are the arguments the reference passes
    (``tet_spheres.cpp:148``).  ``scale=1`` divides each row by its neighbour count.
    """
    if face_neighbor != 1:
        raise NotImplementedError("only the face-neighbour Laplacian the reference requests")
    nbr = face_adjacency(tets)
    T = nbr.shape[0]
    deg = (nbr >= 0).sum(axis=1).astype(np.float64)
    t, k = np.nonzero(nbr >= 0)
    L = sp.csr_matrix((-np.ones(t.size), (t, nbr[t, k])), shape=(T, T)) + sp.diags(deg)
    if scale:
        w = np.where(deg > 0, 1.0 / np.maximum(deg, 1.0), 0.0)
        L = sp.diags(w) @ L
    return L.tocsr()


def _det3(F):
    return (F[:, 0, 0] * (F[:, 1, 1] * F[:, 2, 2] - F[:, 1, 2] * F[:, 2, 1])
            - F[:, 0, 1] * (F[:, 1, 0] * F[:, 2, 2] - F[:, 1, 2] * F[:, 2, 0])
            + F[:, 0, 2] * (F[:, 1, 0] * F[:, 2, 1] - F[:, 1, 1] * F[:, 2, 0]))


def _cof3(F):
    """Cofactor matrix = d det / dF (tet_spheres_cuda.cu:32-46)."""
    C = np.empty_like(F)
    C[:, 0, 0] = F[:, 1, 1] * F[:, 2, 2] - F[:, 1, 2] * F[:, 2, 1]
    C[:, 0, 1] = F[:, 1, 2] * F[:, 2, 0] - F[:, 1, 0] * F[:, 2, 2]
    C[:, 0, 2] = F[:, 1, 0] * F[:, 2, 1] - F[:, 1, 1] * F[:, 2, 0]
    C[:, 1, 0] = F[:, 0, 2] * F[:, 2, 1] - F[:, 0, 1] * F[:, 2, 2]
    C[:, 1, 1] = F[:, 0, 0] * F[:, 2, 2] - F[:, 0, 2] * F[:, 2, 0]
    C[:, 1, 2] = F[:, 0, 1] * F[:, 2, 0] - F[:, 0, 0] * F[:, 2, 1]
    C[:, 2, 0] = F[:, 0, 1] * F[:, 1, 2] - F[:, 0, 2] * F[:, 1, 1]
    C[:, 2, 1] = F[:, 0, 2] * F[:, 1, 0] - F[:, 0, 0] * F[:, 1, 2]
    C[:, 2, 2] = F[:, 0, 0] * F[:, 1, 1] - F[:, 0, 1] * F[:, 1, 0]
    return C


class ReferenceEnergyOracle:
    """Sparse-operator restatement of ``TetSpheres`` + forward/backward (see module docstring)."""

    def __init__(self, rest_vertices, tets, dtype=np.float64, laplacian_scale: int = 0):
        X = np.asarray(rest_vertices, dtype=np.float32).astype(np.float64).reshape(-1, 3)
        # ^ the reference receives float32 rest positions and widens them (tet_spheres.cpp:251-254)
        self.tets = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
        self.n = X.shape[0]
        self.nele = self.tets.shape[0]
        self.dtype = np.dtype(dtype)
        G = build_G(X, self.tets)
        L = tet_laplacian(self.tets, 1, laplacian_scale)
        L9 = sp.kron(L, sp.identity(9, format="csr"), format="csr")
        LG = (L9 @ G).tocsr()
        M = (LG.T @ LG).tocsr()                    # G^T L^T L G   (tet_spheres.cpp:148)
        self.G = G.astype(self.dtype)              # fp64 -> fp32 truncation: tet_spheres.cpp:43-45
        self.M = M.astype(self.dtype)
        self.LG = LG.astype(self.dtype)

    # tet_spheres_cuda.cu:118-195
    def energy_terms(self, x, order: int):
        x = np.asarray(x, dtype=self.dtype).reshape(-1)
        Mx = self.M @ x
        sm = self.dtype.type(0.5) * np.dot(Mx, x)
        F = (self.G @ x).reshape(-1, 3, 3)
        J = np.maximum(-_det3(F), 0)
        if order == 2:
            tetJ = J * J
        elif order == 4:
            tetJ = J * J * J * J
        else:
            tetJ = np.zeros_like(J)
        return sm, np.sum(np.abs(tetJ))            # Sasum, cu:185

    def forward(self, x, c1: float, c2: float, order: int):
        sm, bar = self.energy_terms(x, order)
        return self.dtype.type(sm * self.dtype.type(c1) + bar * self.dtype.type(c2))

    # tet_spheres_cuda.cu:197-263
    def backward(self, gradH, x, c1: float, c2: float, order: int):
        x = np.asarray(x, dtype=self.dtype).reshape(-1)
        g = self.dtype.type(c1) * (self.M @ x)
        F = (self.G @ x).reshape(-1, 3, 3)
        J = _det3(F)
        inv = J < 0
        m = np.where(inv, -J, 0)
        if order == 2:
            coef = 2.0 * m
        elif order == 4:
            coef = 4.0 * m * m * m
        else:
            coef = np.zeros_like(m)
        D = (-coef)[:, None, None] * _cof3(F)
        D[~inv] = 0
        g = g + self.dtype.type(c2) * (self.G.T @ D.reshape(-1).astype(self.dtype))
        return (self.dtype.type(gradH) * g).reshape(-1, 3)

    # ---- AMIPS (BASELINE.json north_star names it; the reference has NO such term: SURVEY.md F1) ----------
    # No reference oracle exists for this term.  Definition used by the product (default off, c3 = 0), the
    # conformal AMIPS energy of the tet-meshing literature (Fu et al. 2015; TetWild's mesh-quality energy):
    #     psi_t = tr(F_t^T F_t) / (3 det(F_t)^(2/3)) - 1   for det F_t > 0,    0 otherwise
    # (>= 0, zero exactly for similarity maps; inverted tets are left to the barrier term), summed over tets.
    def amips_terms(self, x):
        F = (self.G @ np.asarray(x, dtype=self.dtype).reshape(-1)).reshape(-1, 3, 3)
        J = _det3(F)
        ok = J > 0
        tr = (F * F).sum(axis=(1, 2))
        psi = np.where(ok, tr / (3.0 * np.where(ok, J, 1.0) ** (2.0 / 3.0)) - 1.0, 0.0)
        return float(psi.sum()), F, J, ok, tr

    def amips_backward(self, gradH, x, c3: float):
        _, F, J, ok, tr = self.amips_terms(x)
        Js = np.where(ok, J, 1.0)
        P = (2.0 / (3.0 * Js ** (2.0 / 3.0)))[:, None, None] * (F - (tr / (3.0 * Js))[:, None, None] * _cof3(F))
        P[~ok] = 0
        g = self.dtype.type(c3) * (self.G.T @ P.reshape(-1).astype(self.dtype))
        return (self.dtype.type(gradH) * g).reshape(-1, 3)

    def inverted_fraction(self, x) -> float:
        F = (self.G @ np.asarray(x, dtype=self.dtype).reshape(-1)).reshape(-1, 3, 3)
        return float(np.mean(_det3(F) < 0))
