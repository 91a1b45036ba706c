"""Per-phase timeline of the tile kernel under warm graph replay (developer tool).
Usage: python tools/phase_timing.py [S] [T] [tile_tets] [nt512]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tssplat_b200 import _capi, tet_spheres_ext as ext  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
TT = int(sys.argv[3]) if len(sys.argv) > 3 else 512
NT512 = int(sys.argv[4]) if len(sys.argv) > 4 else 256
_capi.lib.tsb_debug_set_threads_512(ctypes.c_int(NT512))
pack = make_pack(S, T, seed=0, unique=8)
sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1), tile_tets=TT)
x = torch.from_numpy(perturb(pack, sigma_rel=0.02, seed=0)).cuda()
nt = sp.info["n_tiles"]
dbg = torch.zeros((nt, 16), dtype=torch.int64, device="cuda")
_capi.lib.tsb_debug_set_timing.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
for _ in range(20):
    sp.energy_grad(x, 2e-4 / S, 2e-4, 2)
torch.cuda.synchronize()
_capi.lib.tsb_debug_set_timing(sp._h, dbg.data_ptr())
sp.energy_grad(x, 2e-4 / S, 2e-4, 2)
torch.cuda.synchronize()
_capi.lib.tsb_debug_set_timing(sp._h, None)
d = dbg.cpu().numpy()
names = ["start->bar-init sync", "wait V blob", "phase0 stores", "sync", "wait T blob", "phase1", "energy reduce+sync",
         "wait ELL", "phase2"]
print(f"S={S} T={T} TT={TT} tiles={nt} fill={sp.info['fill']}")
seg = np.diff(d[:, :10], axis=1)
for i, nm in enumerate(names):
    print(f"  {nm:24s} mean {seg[:, i].mean():8.0f}  p10 {np.percentile(seg[:, i], 10):8.0f}  p90 {np.percentile(seg[:, i], 90):8.0f} cycles")
tot = d[:, 9] - d[:, 0]
print(f"  CTA total                mean {tot.mean():8.0f}  min {tot.min()}  max {tot.max()} cycles")
t0 = d[:, 14].min()
start, end, sm = d[:, 14] - t0, d[:, 15] - t0, d[:, 13]
print(f"  kernel span {end.max()} ns; first-wave CTAs start within {np.sort(start)[min(295, nt - 1)]} ns; "
      f"last CTA starts at {start.max()} ns; CTAs per SM min/max {np.bincount(sm.astype(int)).min()}/{np.bincount(sm.astype(int)).max()}")
order = np.argsort(start)
print("  start(ns) of every 50th CTA:", [int(start[i]) for i in order[::50]])
print("  dur(ns)   of every 50th CTA:", [int(end[i] - start[i]) for i in order[::50]])
