"""Torch restatement of the reference's surface gather + vertex-normal splat -- TEST INFRASTRUCTURE ONLY.

Follows ``geometry/tetmesh_geometry.py:33`` (``v_pos = tet_v[surface_vid]``) and ``:39-66``
(``_compute_vertex_normal``: face normals ``cross(v1 - v0, v2 - v0)``, splatted to the three vertices with
``scatter_add_``, zero/degenerate normals (``|n|^2 <= 1e-20``) replaced by (0,0,1), ``F.normalize``).  Written with
``index_add`` so it also runs in fp64 and differentiates through autograd: the checker for
``tssplat_b200.surface`` (forward and backward).  Pinned by ``tests/golden/ref_fixtures.npz``, whose normals were
produced by EXECUTING the reference's own function body (``tests/golden/make_ref_fixtures.py``)."""
import torch


def vertex_normals(tet_v: torch.Tensor, surface_vid: torch.Tensor, surface_f: torch.Tensor):
    v_pos = tet_v[surface_vid]
    i0, i1, i2 = surface_f[:, 0], surface_f[:, 1], surface_f[:, 2]
    fn = torch.cross(v_pos[i1] - v_pos[i0], v_pos[i2] - v_pos[i0], dim=1)
    n = torch.zeros_like(v_pos).index_add(0, i0, fn).index_add(0, i1, fn).index_add(0, i2, fn)
    keep = (n * n).sum(dim=1, keepdim=True) > 1e-20
    n = torch.where(keep, n, torch.tensor([0.0, 0.0, 1.0], dtype=n.dtype, device=n.device).expand_as(n))
    return v_pos, torch.nn.functional.normalize(n, dim=1)
