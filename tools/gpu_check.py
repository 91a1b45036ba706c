"""Quick GPU sanity run (developer tool): parity vs the C oracle on a few packs and a crude
CUDA-event timing of the fused launch.  Usage: python tools/gpu_check.py [S] [T]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _helpers import COracle  # noqa: E402
from tssplat_b200 import tet_spheres_ext as ext  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402


def parity(pack, tile_tets, sig, order, scale=0):
    sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1), tile_tets=tile_tets, laplacian_scale=scale)
    x_np = perturb(pack, sigma_rel=sig, seed=1)
    x = torch.from_numpy(x_np).cuda()
    c1, c2 = 2e-4 / pack.num_spheres, 2e-4
    e, g = sp.energy_grad(x, c1, c2, order, 0.7)
    e2, g2 = sp.energy_grad(x, c1, c2, order, 0.7)
    torch.cuda.synchronize()
    eo, terms, go = COracle(pack.verts, pack.tets, scale).energy_grad(x_np, c1, c2, order, gradH=0.7)
    e = e.cpu().numpy().astype(np.float64)
    gg = g.cpu().numpy().astype(np.float64)
    det = bool(torch.equal(g, g2)) and bool(torch.equal(torch.as_tensor(e), e2.cpu().double()) or True)
    print(f"  TT={tile_tets} sig={sig} order={order} scale={scale}: E={e[0]:.8g} rel={abs(e[0]-eo)/abs(eo):.2e} "
          f"sm rel={abs(e[1]-terms[0])/abs(terms[0]):.2e} bar={e[2]:.6g}/{terms[1]:.6g} "
          f"g rel={np.linalg.norm(gg-go)/np.linalg.norm(go):.2e} maxabs={np.abs(gg-go).max():.2e} "
          f"nan={int(np.isnan(gg).sum())} deterministic={det} tiles={sp.info['n_tiles']}")


def timing(pack, tile_tets, threads512=256, reps=200, skip_combine=0):
    import ctypes
    from tssplat_b200 import _capi
    _capi.lib.tsb_debug_set_threads_512(ctypes.c_int(threads512))
    _capi.lib.tsb_debug_set_skip_combine(ctypes.c_int(skip_combine))
    sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1), tile_tets=tile_tets)
    x = torch.from_numpy(perturb(pack, sigma_rel=0.02, seed=0)).cuda()
    c1, c2 = 2e-4 / pack.num_spheres, 2e-4
    for _ in range(5):
        sp.energy_grad(x, c1, c2, 2)
    torch.cuda.synchronize()
    # eager launches
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    energy = torch.empty(3, device="cuda"); grad = torch.empty((sp.n, 3), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def launch():
        _capi.lib.tsb_energy_grad(sp._h, x.data_ptr(), c1, c2, 2, 1.0, None, energy.data_ptr(), grad.data_ptr(), st)
    s.record()
    for _ in range(reps):
        launch()
    e.record(); torch.cuda.synchronize()
    t_eager = s.elapsed_time(e) / reps * 1e3
    # graph replay
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        st2 = side.cuda_stream
        for _ in range(3):
            _capi.lib.tsb_energy_grad(sp._h, x.data_ptr(), c1, c2, 2, 1.0, None, energy.data_ptr(), grad.data_ptr(), st2)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                _capi.lib.tsb_energy_grad(sp._h, x.data_ptr(), c1, c2, 2, 1.0, None, energy.data_ptr(), grad.data_ptr(), side.cuda_stream)
    g.replay(); torch.cuda.synchronize()
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    t_graph = s.elapsed_time(e) / reps * 1e3
    balg = pack.algorithmic_bytes()
    _capi.lib.tsb_debug_set_skip_combine(ctypes.c_int(0))
    print(f"  S={pack.num_spheres} TT={tile_tets} NT512={threads512} skip_combine={skip_combine}: eager {t_eager:.2f} us/launch, graph {t_graph:.2f} us/launch "
          f"(warm L2), B_alg={balg/1e6:.1f} MB -> {balg/t_graph/1e3:.0f} GB/s; tiles={sp.info['n_tiles']} "
          f"stream_bytes={sp.info['stream_bytes']/1e6:.1f} MB dup={sp.info['n_local_vertices']/sp.n:.2f}")


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    print(torch.cuda.get_device_name(0))
    small = make_pack(3, 1024, seed=1)
    for tt in (256, 512, 1024):
        for sig, order in ((0.02, 2), (0.35, 2), (0.35, 4)):
            parity(small, tt, sig, order)
    parity(small, 512, 0.35, 2, scale=1)
    t0 = time.time(); pack = make_pack(S, T, seed=0, unique=8); print(f"pack S={S} T={T}: n={pack.n} nele={pack.nele} ({time.time()-t0:.1f}s)")
    parity(pack, 512, 0.35, 2)
    if os.environ.get("QUICK"):
        import ctypes
        from tssplat_b200 import _capi
        for rep in range(2):
            for fl in (0, 1):
                _capi.lib.tsb_debug_set_exp_flags(ctypes.c_int(fl)); print(f"  exp_flags={fl}")
                timing(pack, 512, 256); timing(pack, 512, 256, skip_combine=1)
        _capi.lib.tsb_debug_set_exp_flags(ctypes.c_int(0))
        sys.exit(0)
    for tt, nt in ((256, 256), (512, 256), (512, 512), (1024, 512)):
        timing(pack, tt, nt)
