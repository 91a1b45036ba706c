"""Host-side mirror of the reference's energy module (``energies/smooth_barrier.py``): same class
names, constructor arguments, scheduler and order switch, so ``geometry/tetmesh_geometry.py:155-189``
can use it unchanged.  The reference file itself also works as is once ``tet_spheres`` resolves to
this repo (its only other import, ``pypgo``, is unused by the module).

All arithmetic happens in the CUDA library behind ``tet_spheres_ext``; nothing here computes.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from tet_spheres import tet_spheres_ext

__all__ = ["SmoothnessBarrierFunc", "SmoothnessBarrierEnergy"]


#: route SmoothnessBarrierEnergy through the C++ autograd bridge when it has been built (same launches, same
#: semantics, no Python between autograd and the C ABI); False forces the Python Function below
use_native_autograd = True


class SmoothnessBarrierFunc(torch.autograd.Function):
    """autograd bridge (``energies/smooth_barrier.py:9-31``): forward returns the 0-dim energy,
    backward returns ``(dE/dx * grad_output, None, None, None, None)`` and short-circuits on a
    ``None`` grad_output."""

    @staticmethod
    def forward(ctx, x_cur, tet_sp, c1, c2, order):      # ctx-style like the reference's (no per-call signature binding)
        ctx.save_for_backward(x_cur)
        ctx.constants = (tet_sp, c1, c2, order)
        return tet_spheres_ext.forward(x_cur, tet_sp, c1, c2, order)

    @staticmethod
    def backward(ctx, grad_output):
        if grad_output is None:
            return (None,) * 5
        (x_cur,) = ctx.saved_tensors
        tet_sp, c1, c2, order = ctx.constants
        grad = tet_spheres_ext.backward(grad_output, x_cur, tet_sp, c1, c2, int(order))
        return grad, None, None, None, None


class SmoothnessBarrierEnergy(torch.nn.Module):
    """``SmoothnessBarrierEnergy(tet_v, tet_f, FLAGS)`` (``energies/smooth_barrier.py:34-67``).

    ``tet_v``: numpy [n,3] REST positions; ``tet_f``: numpy [nele,4]; ``FLAGS``: mapping or object
    with ``smooth_eng_coeff``, ``barrier_coeff``, ``increase_order_iter`` (``config/gso.yaml:9-11``).
    """

    def __init__(self, tet_v, tet_f, FLAGS) -> None:
        super().__init__()
        v_flat = np.asarray(tet_v).flatten().astype(np.float32)
        f_flat = np.asarray(tet_f).flatten().astype(np.int32)
        self.tet_sp = tet_spheres_ext.TetSpheres(v_flat, f_flat)
        self.FLAGS = SimpleNamespace(**FLAGS) if isinstance(FLAGS, dict) else FLAGS
        self.smooth_eng_func = SmoothnessBarrierFunc          # the reference instantiates it; .apply is static

    def coeff_scheduler(self, it):
        """Both coefficients times ``2 ** (4 |sin(min(it/2400 * pi, pi/2))|)`` in [1, 16]
        (``energies/smooth_barrier.py:47-58``)."""
        phase = min(it / 300.0 / 4 * 0.5 * math.pi, 0.5 * math.pi)
        multiplier = math.pow(2, abs(math.sin(phase)) * 4)
        return self.FLAGS.smooth_eng_coeff * multiplier, self.FLAGS.barrier_coeff * multiplier

    def forward(self, x, it, c1, c2):
        order = 4 if it > self.FLAGS.increase_order_iter else 2     # smooth_barrier.py:61-63
        if (use_native_autograd and self.smooth_eng_func is SmoothnessBarrierFunc and tet_spheres_ext.fuse_backward_into_forward
                and not tet_spheres_ext.return_cpu_scalar):
            ns = self.tet_sp.native_state()        # C++ torch::autograd::Function over the same C ABI (csrc/torch_binding.cpp)
            if ns is not None:
                return ns[0].energy(x, ns[1], float(c1), float(c2), order, self.tet_sp)
        return self.smooth_eng_func.apply(x, self.tet_sp, c1, c2, order)
