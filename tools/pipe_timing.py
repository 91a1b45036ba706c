"""Event timeline of the pipelined tile kernel (developer tool): per CTA and tile, clock64 stamps of
TMA issue, x-gather start/done, tet-math start/end, row-gather start/end."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tssplat_b200 import _capi, tet_spheres_ext as ext  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pack = make_pack(S, 4096, seed=0, unique=8)
sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1), tile_tets=512)
x = torch.from_numpy(perturb(pack, sigma_rel=0.02, seed=0)).cuda()
dbg = torch.zeros((148, 64), dtype=torch.int64, device="cuda")
_capi.lib.tsb_debug_set_timing.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
for _ in range(20):
    sp.energy_grad(x, 2e-4 / S, 2e-4, 2)
torch.cuda.synchronize()
_capi.lib.tsb_debug_set_timing(sp._h, dbg.data_ptr())
sp.energy_grad(x, 2e-4 / S, 2e-4, 2)
torch.cuda.synchronize()
_capi.lib.tsb_debug_set_timing(sp._h, None)
d = dbg.cpu().numpy()
t0 = d[:, 63:64]
names = ["x_start", "x_done", "p1_start", "p1_end", "a2_start", "a2_end", "tma_issue"]
print("tiles", sp.info["n_tiles"], "fill", sp.info["fill"])
for k in range(5):
    row = d[:, k * 8:k * 8 + 7] - t0
    ok = d[:, k * 8 + 2] > 0
    if not ok.any():
        break
    print(f"tile {k} ({ok.sum()} CTAs): " + "  ".join(f"{n}={np.median(row[ok, i]):.0f}" for i, n in enumerate(names)))
