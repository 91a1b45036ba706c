"""Where the host-side time of the two end-to-end paths goes (64 x 4096 pack): CPU time per call of
tsb_energy_grad_host through ctypes vs the device-side rate, and a cProfile of the autograd surface step.
Usage: python tools/profile_hostpaths.py [spheres]"""
import cProfile
import os
import pstats
import sys
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tssplat_b200 import _capi  # noqa: E402
from tssplat_b200 import tet_spheres_ext as ext  # noqa: E402
from tssplat_b200.energies import SmoothnessBarrierEnergy, SmoothnessBarrierFunc  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pack = make_pack(S, 4096, seed=0, unique=8)
sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1))
n = sp.n
x_host = torch.from_numpy(perturb(pack, sigma_rel=0.02, seed=0)).pin_memory()
g_host = torch.empty((n, 3)).pin_memory()
e_host = torch.empty(4).pin_memory()
c1, c2 = 2e-4 / S, 2e-4


def rate(fn, steps, label):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{label}: CPU issue {1e6 * (t1 - t0) / steps:.1f} us/call, device {1e3 * e0.elapsed_time(e1) / steps:.1f} us/step, "
          f"wall incl. drain {1e6 * (t2 - t0) / steps:.1f} us/step", flush=True)


rate(lambda: ext.energy_grad_host(sp, x_host, c1, c2, 2, 1.0, e_host, g_host), 2000, "energy_grad_host (python wrapper)")
st = torch.cuda.current_stream().cuda_stream
xp, ep, gp, h, fn = x_host.data_ptr(), e_host.data_ptr(), g_host.data_ptr(), sp._h, _capi.lib.tsb_energy_grad_host
rate(lambda: fn(h, xp, c1, c2, 2, 1.0, ep, gp, st), 2000, "tsb_energy_grad_host (bare ctypes)")
rate(lambda: fn(h, xp, c1, c2, 2, 1.0, ep, None, st), 2000, "tsb_energy_grad_host, energy only")


def sync_step():
    fn(h, xp, c1, c2, 2, 1.0, ep, gp, st)
    torch.cuda.current_stream().synchronize()


rate(sync_step, 1000, "tsb_energy_grad_host + stream sync every call")

x_dev = x_host.cuda()
en = torch.zeros(3, device="cuda")
gr = torch.empty((n, 3), device="cuda")
rate(lambda: _capi.lib.tsb_energy_grad(h, x_dev.data_ptr(), c1, c2, 2, 1.0, None, en.data_ptr(), gr.data_ptr(), st), 4000,
     "tsb_energy_grad (device buffers, eager launches)")
rate(lambda: x_dev.copy_(x_host, non_blocking=True), 2000, "H2D copy alone (torch)")
rate(lambda: g_host.copy_(gr, non_blocking=True), 2000, "D2H copy alone (torch)")

eng = SmoothnessBarrierEnergy.__new__(SmoothnessBarrierEnergy)
torch.nn.Module.__init__(eng)
eng.tet_sp, eng.FLAGS = sp, SimpleNamespace(smooth_eng_coeff=c1, barrier_coeff=c2, increase_order_iter=10 ** 9)
eng.smooth_eng_func = SmoothnessBarrierFunc
tet_v = torch.nn.Parameter(torch.empty((n, 3), device="cuda"))
e0_host = torch.empty(()).pin_memory()


def autograd_step():
    tet_v.grad = None
    with torch.no_grad():
        tet_v.copy_(x_host, non_blocking=True)
    e = eng(tet_v, 0, c1, c2)
    e.backward()
    g_host.copy_(tet_v.grad, non_blocking=True)
    e0_host.copy_(e.detach(), non_blocking=True)


rate(autograd_step, 1000, "autograd surface step")


def fwd_bwd_only():
    tet_v.grad = None
    e = eng(tet_v, 0, c1, c2)
    e.backward()


rate(fwd_bwd_only, 1000, "autograd forward+backward only (no copies)")
pr = cProfile.Profile()
pr.enable()
for _ in range(500):
    autograd_step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)


class _Floor(torch.autograd.Function):
    """What torch.autograd itself costs: a Function whose forward and backward only hand back tensors."""

    @staticmethod
    def forward(ctx, x, e, g):
        ctx.g = g
        return e[0]

    @staticmethod
    def backward(ctx, go):
        return ctx.g, None, None


e_pre, g_pre = torch.zeros(3, device="cuda"), torch.zeros((n, 3), device="cuda")


def floor_step():
    tet_v.grad = None
    _Floor.apply(tet_v, e_pre, g_pre).backward()


rate(floor_step, 1000, "torch.autograd floor (custom Function that launches nothing, incl. backward's ones_like)")
