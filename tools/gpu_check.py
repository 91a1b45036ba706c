"""Quick GPU check: parity of the fused kernel against the C oracle over the kernel variants, then
CUDA-graph timings at several pack sizes.  Usage: python tools/gpu_check.py [parity|time|all] [sizes...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _helpers import GOLDEN, COracle  # noqa: E402
from tssplat_b200 import _capi  # noqa: E402
from tssplat_b200 import tet_spheres_ext as ext  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402


def parity(name, verts, tets, sig, order, scale=0, **kw):
    sp = ext.TetSpheres(np.ascontiguousarray(verts, dtype=np.float32).reshape(-1),
                        np.ascontiguousarray(tets, dtype=np.int32).reshape(-1), laplacian_scale=scale, **kw)
    x_np = perturb(verts, tets, sig, 1)
    x = torch.from_numpy(x_np).cuda()
    e, g = sp.energy_grad(x, 2e-4, 3e-4, order, 0.7)
    e2, g2 = sp.energy_grad(x, 2e-4, 3e-4, order, 0.7)
    torch.cuda.synchronize()
    eo, terms, go = COracle(verts, tets, scale).energy_grad(x_np, 2e-4, 3e-4, order, gradH=0.7)
    e = e.cpu().numpy().astype(np.float64)
    g = g.cpu().numpy().astype(np.float64)
    rel_e = abs(e[0] - eo) / max(abs(eo), 1e-30)
    rel_g = np.linalg.norm(g - go) / np.linalg.norm(go)
    rep = float((g2.cpu().numpy() - g).__abs__().max())
    i = sp.info
    print(f"  {name:14s} sig={sig} order={order} scale={scale} {kw}: E={e[0]:.8g} relE={rel_e:.2e} relG={rel_g:.2e} "
          f"repeat_maxdiff={rep:.1e} | grid={i['grid']} nw={i['warps_per_cta']} global={i['mode_global']} smem={i['smem_bytes']} "
          f"segs={i['n_segments']} pad={i['nnz_padded'] / max(i['nnz'], 1):.3f}", flush=True)
    return rel_e < 1e-5 and rel_g < 1e-5


def run_parity():
    ok = True
    pk = make_pack(3, 1024, seed=1)
    d = np.load(os.path.join(GOLDEN, "a_veg_mesh.npz"))
    pk2 = make_pack(2, 1500, seed=4)
    for kw in ({}, {"warps_per_cta": 8}, {"force_global": True}, {"warps_per_cta": 8, "force_global": True}, {"ring_slots": 2}):
        for sig, order in ((0.02, 2), (0.35, 2), (0.35, 4)):
            ok &= parity("pack3x1024", pk.verts, pk.tets, sig, order, **kw)
        ok &= parity("pack2x1500", pk2.verts, pk2.tets, 0.35, 2, scale=1, **kw)
        ok &= parity("a_veg", d["verts"], d["tets"], 0.35, 2, **kw)
    pk3 = make_pack(16, 4096, seed=0, unique=4)
    for kw in ({}, {"warps_per_cta": 8}):
        ok &= parity("pack16x4096", pk3.verts, pk3.tets, 0.02, 2, **kw)
        ok &= parity("pack16x4096", pk3.verts, pk3.tets, 0.35, 4, **kw)
    print("PARITY", "OK" if ok else "FAILED", flush=True)
    return ok


def timing(S, reps=20, **kw):
    pack = make_pack(S, 4096, seed=0, unique=8)
    t0 = time.time()
    sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1), **kw)
    t_create = time.time() - t0
    x = torch.from_numpy(perturb(pack, sigma_rel=0.02, seed=0)).cuda()
    energy = torch.zeros(3, device="cuda")
    grad = torch.empty((pack.n, 3), device="cuda")
    st = torch.cuda.Stream()
    c1, c2 = 2e-4 / S, 2e-4

    def launch():
        rc = _capi.lib.tsb_energy_grad(sp._h, x.data_ptr(), c1, c2, 2, 1.0, None, energy.data_ptr(), grad.data_ptr(), st.cuda_stream)
        assert rc == 0, _capi.last_error(sp._h)
    with torch.cuda.stream(st):
        for _ in range(5):
            launch()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                launch()
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nrep = max(3, int(20000 / (reps * max(S / 16, 1))))
        e0.record(st)
        for _ in range(nrep):
            g.replay()
        e1.record(st)
        st.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (nrep * reps)
    b_alg = pack.algorithmic_bytes()
    i = sp.info
    print(f"  S={S:5d} {kw}: {us:8.2f} us/step  B_alg {b_alg / 1e6:.1f} MB -> {b_alg / us / 1e3:7.1f} GB/s ({b_alg / us / 1e3 / 6573.2:.3f} of HBM peak); "
          f"plan {i['stream_bytes'] / 1e6:.1f} MB/step; create {t_create:.2f}s grid={i['grid']} segs={i['n_segments']} "
          f"pad={i['nnz_padded'] / max(i['nnz'], 1):.3f} nnz/row={i['nnz'] / i['n']:.1f}", flush=True)
    return us


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    sizes = [int(a) for a in sys.argv[2:]] or [16, 64, 256, 1024]
    print(torch.cuda.get_device_name(0), flush=True)
    if what in ("parity", "all"):
        run_parity()
    if what in ("time", "all"):
        for S in sizes:
            for kw in ({}, {"warps_per_cta": 8, "ring_slots": 2}):
                try:
                    timing(S, **kw)
                except Exception as ex:  # keep going: a variant may not fit
                    print(f"  S={S} {kw}: {type(ex).__name__}: {ex}", flush=True)
