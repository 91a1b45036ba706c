"""Energy-only trainer loop: BASELINE configs[2] minus the rasterizer (nvdiffrast, data, TetWild are
absent -- SURVEY.md F7).  What `trainer.py:71-132` does around the energy, with the same pieces:
coefficient scheduler + order switch (energies/smooth_barrier.py), the autograd surface,
AdamUniform (utils/optimizer.py) with grad_limit, CosineAnnealingLR -- all on the B200 path.

    python tools/energy_only_loop.py [spheres] [iters]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tssplat_b200.energies import SmoothnessBarrierEnergy  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402
from tssplat_b200.optimizer import AdamUniform  # noqa: E402


def run(spheres=64, iters=1500, seed=0, log_every=0):
    pack = make_pack(spheres, 4096, seed=seed, unique=8)
    flags = dict(smooth_eng_coeff=2e-4 / spheres, barrier_coeff=2e-4, increase_order_iter=1000)   # gso.yaml:9-11, tetmesh_geometry.py:243
    eng = SmoothnessBarrierEnergy(pack.verts, pack.tets, flags)
    tet_v = torch.nn.Parameter(torch.from_numpy(perturb(pack, sigma_rel=0.35, seed=1)).cuda())
    opt = AdamUniform([tet_v], grad_limit=True, grad_limit_values=[0.01, 0.01], grad_limit_iters=[1500], lr=0.2)   # gso.yaml:35-39
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=iters)
    torch.cuda.synchronize()
    e_first = e_last = None
    t0 = time.perf_counter()
    for it in range(iters):
        c1, c2 = eng.coeff_scheduler(it)
        reg_loss = eng(tet_v, it, c1, c2)
        opt.zero_grad(set_to_none=True)
        reg_loss.backward()
        opt.step()
        sched.step()
        if it == 0:
            e_first = reg_loss.detach()
        e_last = reg_loss.detach()
        if log_every and it % log_every == 0:
            print(f"  it {it:5d} reg_loss {float(reg_loss):.6g}")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return iters / dt, float(e_first), float(e_last)


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    run(S, 50)
    rate, e0, e1 = run(S, N, log_every=max(1, N // 6))
    print(f"energy-only trainer loop, {S} spheres x 4096 tets: {rate:.0f} it/s (wall clock, Python-bound), "
          f"reg_loss {e0:.4g} -> {e1:.4g}")
