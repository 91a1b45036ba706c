"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped otherwise): sphere-per-rank
``ShardedEnergy`` over NCCL must reproduce the single-GPU energy (after the scalar all-reduce) and
the concatenated per-rank gradients must equal the single-GPU gradient -- SURVEY.md section 8(e)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tssplat_b200.mesh import make_pack, perturb

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from tssplat_b200.sharding import ShardedEnergy
        pack = make_pack(6, 1024, seed=21)
        x = perturb(pack, sigma_rel=0.3, seed=3)
        sh = ShardedEnergy(pack, device=torch.device("cuda", rank))
        lo, hi = sh.sphere_range
        v0, v1 = int(pack.vert_offsets[lo]), int(pack.vert_offsets[hi])
        energy, grad = sh.energy_grad(torch.from_numpy(x[v0:v1]).cuda(rank), 2e-4 / 6, 2e-4, 2)
        sh.wait()
        torch.cuda.synchronize()
        q.put((rank, v0, v1, energy.cpu().numpy().copy(), grad.cpu().numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_sharded_energy_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world, port = 2, 29600 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    from tssplat_b200 import tet_spheres_ext as ext
    pack = make_pack(6, 1024, seed=21)
    x = perturb(pack, sigma_rel=0.3, seed=3)
    sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1))
    e1, g1 = sp.energy_grad(torch.from_numpy(x).cuda(), 2e-4 / 6, 2e-4, 2)
    e1, g1 = e1.cpu().numpy(), g1.cpu().numpy()
    for _, _, _, e, _ in out:
        assert np.allclose(e, e1, rtol=1e-5)
    g = np.concatenate([o[4] for o in out])
    assert out[0][2] == out[1][1] and g.shape == g1.shape
    assert np.linalg.norm(g - g1) <= 1e-5 * np.linalg.norm(g1)
