"""Drop-in for the reference's ``utils/optimizer.py`` ``AdamUniform``:
same constructor arguments, ``step()`` / ``reset()`` semantics and state names (``step``, ``g1``,
``g2``), but each parameter update is two CUDA launches through the C ABI
(``tsb_adam_uniform_step``) with no host synchronisation, instead of ~12 torch kernels and the
``if s > m`` host sync of ``utils/optimizer.py:83-86``.

Adam with the second moment replaced by its global maximum (``utils/optimizer.py:74``) and an
optional clamp of the largest step component to ``grad_limit_values[ptr]`` (``:76-86``).
"""
from __future__ import annotations

import torch

from . import _capi
from . import tet_spheres_ext as _ext

__all__ = ["AdamUniform"]


class AdamUniform(torch.optim.Optimizer):
    def __init__(self, params, grad_limit=False, grad_limit_values=(0.05, 0.01), grad_limit_iters=(4000,),
                 lr=0.1, betas=(0.9, 0.999)):
        defaults = dict(lr=lr, betas=betas)
        self.grad_limit = grad_limit
        super().__init__(params, defaults)
        self.cc = 0                                            # utils/optimizer.py:17
        if grad_limit:
            self.grad_limit_values = list(grad_limit_values)
            self.grad_limit_iters = list(grad_limit_iters)
            self.grad_limit_ptr = 0

    @torch.no_grad()
    def reset(self):                                           # utils/optimizer.py:27-35
        for group in self.param_groups:
            for p in group["params"]:
                state = self.state[p]
                state["step"] = 0
                state["g1"] = torch.zeros_like(p.data)
                state["g2"] = torch.zeros_like(p.data)

    @torch.no_grad()
    def step(self):                                            # utils/optimizer.py:37-89
        for group in self.param_groups:
            lr = group["lr"]
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("AdamUniform (tssplat_b200) needs contiguous float32 CUDA parameters")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["g1"] = torch.zeros_like(p.data)
                    state["g2"] = torch.zeros_like(p.data)
                if "_work" not in state:
                    state["_work"] = torch.zeros(4, dtype=torch.float32, device=p.device)
                state["step"] += 1
                limit = 0.0
                if self.grad_limit:
                    limit = float(self.grad_limit_values[self.grad_limit_ptr])      # :77
                    if self.grad_limit_ptr < len(self.grad_limit_iters):            # :79-81
                        if self.cc >= self.grad_limit_iters[self.grad_limit_ptr]:
                            self.grad_limit_ptr += 1
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                # the C ABI selects the device that owns p (no torch.cuda.device context needed)
                rc = _capi.lib.tsb_adam_uniform_step(
                    p.data_ptr(), grad.data_ptr(), state["g1"].data_ptr(), state["g2"].data_ptr(), p.numel(),
                    float(lr), float(b1), float(b2), int(state["step"]), limit, state["_work"].data_ptr(),
                    _ext._stream_ptr(p.device))
                if rc:
                    _capi.check(rc, None, "AdamUniform.step")
                _ext.note_parameters_changed()         # p.data changed without bumping p._version
                self.cc += 1                                   # :89
