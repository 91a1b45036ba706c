"""ncu target: a few fused energy+gradient launches on one pack.
Usage: python tools/prof_target.py [S] [launches] [sigma]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tssplat_b200 import tet_spheres_ext as ext  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 6
SIG = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
pack = make_pack(S, 4096, seed=0, unique=8)
sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1))
x = torch.from_numpy(perturb(pack, sigma_rel=SIG, seed=0)).cuda()
for _ in range(N):
    e, g = sp.energy_grad(x, 2e-4 / S, 2e-4, 2)
torch.cuda.synchronize()
print(S, sp.info["grid"], float(e[0]))
