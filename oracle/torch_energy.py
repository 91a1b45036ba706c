"""Vanilla-PyTorch CPU restatement of the reference's energy pipeline -- TEST / BASELINE
INFRASTRUCTURE ONLY (never imported by the product).

The reference advertises a "vanilla PyTorch version" of its energies (``README.md:37,111``) but
does not ship it (``energies/smooth_barrier.py:7`` is a commented-out import).  This file is what
such a path would compute, following the extension's own algorithm step by step with torch
sparse ops in fp32 and autograd for the gradient:

    xTemp = GTLTLG @ x ; sm = 0.5 * dot(xTemp, x)          tet_spheres_cuda.cu:131,154,157
    F     = G @ x                                            tet_spheres_cuda.cu:167
    tetJ  = max(-det F, 0) ** order ; bar = sum(tetJ)        tet_spheres_cuda.cu:48-66,185
    E     = sm * c1 + bar * c2                               tet_spheres_cuda.cu:191

It is the ``cpu_baseline`` / ``--impl reference`` leg of bench.py ("kind": "port"): a restatement
of the reference's math, not the reference itself (which needs libpgo + a GPU).  The operators come
from oracle/tet_energy_oracle.py (same libpgo assumption for L; PARITY UNPINNED).
"""
from __future__ import annotations

import numpy as np
import torch

from .tet_energy_oracle import ReferenceEnergyOracle


def _to_torch_csr(m, dtype):
    m = m.tocsr()
    return torch.sparse_csr_tensor(torch.from_numpy(m.indptr.astype(np.int64)),
                                   torch.from_numpy(m.indices.astype(np.int64)),
                                   torch.from_numpy(m.data.astype(np.float64)).to(dtype),
                                   size=m.shape)


class TorchEnergy(torch.nn.Module):
    def __init__(self, rest_vertices, tets, dtype=torch.float32):
        super().__init__()
        orc = ReferenceEnergyOracle(rest_vertices, tets)       # fp64 operators
        self.n, self.nele = orc.n, orc.nele
        self.M = _to_torch_csr(orc.M, dtype)                   # truncated like tet_spheres.cpp:43-45
        self.G = _to_torch_csr(orc.G, dtype)
        self.dtype = dtype

    def forward(self, x: torch.Tensor, c1: float, c2: float, order: int) -> torch.Tensor:
        xf = x.reshape(-1, 1).to(self.dtype)
        sm = 0.5 * (torch.sparse.mm(self.M, xf) * xf).sum()
        F = torch.sparse.mm(self.G, xf).reshape(-1, 3, 3)
        J = torch.clamp(-torch.linalg.det(F), min=0)
        bar = (J ** order).sum() if order in (2, 4) else J.sum() * 0
        return sm * c1 + bar * c2


def time_fwd_bwd(rest_vertices, tets, x_np, c1, c2, order, iters=10, warmup=2, threads=None):
    """Median seconds per forward+backward on the CPU with ``threads`` torch threads."""
    import time
    if threads:
        torch.set_num_threads(int(threads))
    mod = TorchEnergy(rest_vertices, tets)
    x = torch.from_numpy(np.asarray(x_np, dtype=np.float32)).clone().requires_grad_(True)
    ts = []
    for i in range(warmup + iters):
        x.grad = None
        t0 = time.perf_counter()
        e = mod(x, c1, c2, order)
        e.backward()
        t1 = time.perf_counter()
        if i >= warmup:
            ts.append(t1 - t0)
    return float(np.median(ts)), float(e.detach()), x.grad.detach().numpy().copy()
