"""N>1 host logic on CPU: sphere partition + the scalar all-reduce over gloo (world_size 2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _helpers import COracle
from tssplat_b200.mesh import make_pack, perturb
from tssplat_b200.sharding import allreduce_energy, partition_spheres, shard_pack


def test_partition_covers_and_balances():
    for sizes, ws in (([4096] * 64, 8), ([4096] * 7, 3), ([100, 900, 50, 50, 400], 2), ([10], 4), ([5, 5], 2)):
        parts = partition_spheres(sizes, ws)
        assert len(parts) == ws and parts[0][0] == 0 and parts[-1][1] == len(sizes)
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:])) and all(lo <= hi for lo, hi in parts)
    assert partition_spheres([4096] * 64, 8) == [(8 * r, 8 * r + 8) for r in range(8)]
    loads = [sum([100, 900, 50, 50, 400][lo:hi]) for lo, hi in partition_spheres([100, 900, 50, 50, 400], 2)]
    assert max(loads) <= 1000


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pack = make_pack(5, 256, seed=11)
        x = perturb(pack, sigma_rel=0.3, seed=2)
        local, (lo, hi) = shard_pack(pack, rank, world)
        v0, v1 = int(pack.vert_offsets[lo]), int(pack.vert_offsets[hi])
        # stand-in for the per-rank GPU launch: the oracle on this rank's spheres
        e, terms, g = COracle(local.verts, local.tets).energy_grad(x[v0:v1], 2e-4, 3e-4, 2)
        energy = torch.tensor([e, terms[0], terms[1]], dtype=torch.float64)
        work = allreduce_energy(energy, async_op=True)
        work.wait()
        q.put((rank, lo, hi, energy.numpy().copy(), g))
    finally:
        dist.destroy_process_group()


def test_sphere_per_rank_allreduce_gloo():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pack = make_pack(5, 256, seed=11)
    x = perturb(pack, sigma_rel=0.3, seed=2)
    e, terms, g = COracle(pack.verts, pack.tets).energy_grad(x, 2e-4, 3e-4, 2)
    for _, _, _, energy, _ in out:                                   # every rank holds the global sums
        assert energy[0] == pytest.approx(e, rel=1e-12) and energy[1] == pytest.approx(terms[0], rel=1e-12)
    g_cat = np.concatenate([o[4] for o in out])                      # gradients need no exchange
    assert np.abs(g_cat - g).max() <= 1e-12 * np.abs(g).max()
    assert out[0][2] == out[1][1] and out[0][1] == 0 and out[1][2] == 5
