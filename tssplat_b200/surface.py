"""Surface gather + vertex normals on the GPU: what
``TetMeshGeometryForwardData`` computes on every geometry forward -- ``v_pos = tet_v[surface_vid]``
(``geometry/tetmesh_geometry.py:33``) and ``_compute_vertex_normal()`` (``:39-66``) -- as ONE kernel
through the C ABI (``tsb_surface_*``), with the analytic backward (two kernels), deterministic summation
order and no atomics.  Differentiable through ``torch.autograd.Function`` like the energy."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi

__all__ = ["SurfaceNormals", "SurfaceForwardData"]


def _stream(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


class SurfaceNormals:
    """Owns the surface topology on the device.  ``surface_vid`` int [nsv] (tet-mesh vertex of each surface
    vertex), ``surface_f`` int [nsf,3] (triangles over surface-vertex ids): the outputs of
    ``mesh_utils.get_surface_vf`` / :func:`tssplat_b200.mesh.surface_vf`."""

    def __init__(self, surface_vid, surface_f, n_tet_vertices: int, device=None):
        self._h = None
        if not torch.cuda.is_available():
            raise RuntimeError("tssplat_b200.surface needs a CUDA device (B200); there is no CPU path")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        sv = np.ascontiguousarray(np.asarray(surface_vid).reshape(-1), dtype=np.int32)
        sf = np.ascontiguousarray(np.asarray(surface_f).reshape(-1), dtype=np.int32)
        if sf.size % 3:
            raise RuntimeError("surface_f must hold triangles")
        self.nsv, self.nsf, self.n = int(sv.size), int(sf.size // 3), int(n_tet_vertices)
        h = C.c_void_p()
        rc = _capi.lib.tsb_surface_create(sv.ctypes.data, self.nsv, sf.ctypes.data, self.nsf, self.n, self.device.index, C.byref(h))
        if rc:
            raise RuntimeError(f"SurfaceNormals: {(_capi.lib.tsb_surface_last_error(None) or b'').decode()} (code {rc})")
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _capi.lib.tsb_surface_destroy(h)
            except Exception:
                pass

    def _check(self, rc):
        if rc:
            raise RuntimeError(f"tssplat_b200.surface: {(_capi.lib.tsb_surface_last_error(self._h) or b'').decode()} (code {rc})")

    def forward(self, tet_v: torch.Tensor, want_normals: bool = True):
        tv = tet_v.detach()
        if not (tv.is_cuda and tv.dtype == torch.float32 and tv.numel() == 3 * self.n):
            raise RuntimeError("tet_v must be a float32 CUDA tensor [n,3]")
        tv = tv if tv.is_contiguous() else tv.contiguous()
        v_pos = torch.empty((self.nsv, 3), dtype=torch.float32, device=self.device)
        v_nrm = torch.empty((self.nsv, 3), dtype=torch.float32, device=self.device) if want_normals else None
        self._check(_capi.lib.tsb_surface_forward(self._h, tv.data_ptr(), v_pos.data_ptr(), v_nrm.data_ptr() if want_normals else None,
                                                  _stream(self.device)))
        return v_pos, v_nrm

    def backward(self, tet_v: torch.Tensor, g_pos, g_nrm) -> torch.Tensor:
        tv = tet_v.detach()
        tv = tv if tv.is_contiguous() else tv.contiguous()
        out = torch.empty((self.n, 3), dtype=torch.float32, device=self.device)
        gp = g_pos.contiguous() if g_pos is not None else None
        gn = g_nrm.contiguous() if g_nrm is not None else None
        self._check(_capi.lib.tsb_surface_backward(self._h, tv.data_ptr(), gp.data_ptr() if gp is not None else None,
                                                   gn.data_ptr() if gn is not None else None, out.data_ptr(), _stream(self.device)))
        return out

    def __call__(self, tet_v: torch.Tensor):
        """Differentiable (v_pos, v_nrm)."""
        return _SurfaceFunc.apply(tet_v, self)


class _SurfaceFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tet_v, surf):
        ctx.surf = surf
        ctx.save_for_backward(tet_v)
        v_pos, v_nrm = surf.forward(tet_v)
        return v_pos, v_nrm

    @staticmethod
    def backward(ctx, g_pos, g_nrm):
        (tet_v,) = ctx.saved_tensors
        return ctx.surf.backward(tet_v, g_pos, g_nrm).reshape(tet_v.shape), None


class SurfaceForwardData:
    """The surface part of ``TetMeshGeometryForwardData`` (``geometry/tetmesh_geometry.py:25-66``): ``v_pos``,
    ``t_pos_idx`` and ``_compute_vertex_normal()``, computed by the fused kernel."""

    def __init__(self, tet_v: torch.Tensor, surf: SurfaceNormals, surface_f: torch.Tensor):
        self.tet_v = tet_v
        self.v_pos, self._v_nrm = surf(tet_v)
        self.t_pos_idx = surface_f

    def _compute_vertex_normal(self):
        return self._v_nrm
