"""Profiling target: N fused launches on an S x T pack (developer tool, run under ncu).
Usage: python tools/prof_target.py [S] [T] [tile_tets] [launches] [energy_only]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tssplat_b200 import tet_spheres_ext as ext  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
TT = int(sys.argv[3]) if len(sys.argv) > 3 else 512
N = int(sys.argv[4]) if len(sys.argv) > 4 else 8
EO = int(sys.argv[5]) if len(sys.argv) > 5 else 0
pack = make_pack(S, T, seed=0, unique=8)
sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1), tile_tets=TT)
x = torch.from_numpy(perturb(pack, sigma_rel=0.02, seed=0)).cuda()
for _ in range(N):
    e, g = sp.energy_grad(x, 2e-4 / S, 2e-4, 2, want_grad=not EO)
torch.cuda.synchronize()
print("done", float(e[0]), sp.info)
