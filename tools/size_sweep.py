"""Step / tile-kernel time vs pack size for the compiled tile variants (developer tool)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tssplat_b200 import _capi, tet_spheres_ext as ext  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402
lib = _capi.lib


def bench(sp, x, S, reps, rotate=None):
    c1, c2 = 2e-4 / S, 2e-4
    energy = torch.empty(3, device="cuda"); grad = torch.empty((sp.n, 3), device="cuda")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            lib.tsb_energy_grad(sp._h, x.data_ptr(), c1, c2, 2, 1.0, None, energy.data_ptr(), grad.data_ptr(), side.cuda_stream)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                lib.tsb_energy_grad(sp._h, x.data_ptr(), c1, c2, 2, 1.0, None, energy.data_ptr(), grad.data_ptr(), side.cuda_stream)
        g.replay(); side.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(side); g.replay(); e.record(side); side.synchronize()
    return s.elapsed_time(e) / reps * 1e3


for S in [int(a) for a in sys.argv[1:]] or [16, 64, 256, 1024]:
    pack = make_pack(S, 4096, seed=0, unique=8)
    x = torch.from_numpy(perturb(pack, sigma_rel=0.02, seed=0)).cuda()
    balg = pack.algorithmic_bytes()
    for v4, tt in ((1, 512), (1, 256), (1, 1024)):
        sp = ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1), tile_tets=tt)
        reps = max(4, min(200, int(4000 / S)))
        res = []
        for skip in (0, 1):
            lib.tsb_debug_set_skip_combine(ctypes.c_int(skip))
            res.append(bench(sp, x, S, reps))
        lib.tsb_debug_set_skip_combine(ctypes.c_int(0))
        print(f"S={S:5d} TT={tt:4d} tiles={sp.info['n_tiles']:6d} fill={sp.info['fill']:4d}: step {res[0]:8.2f} us "
              f"(tile kernel only {res[1]:8.2f} us)  B_alg {balg/1e6:7.1f} MB -> {balg/res[0]/1e3:6.0f} GB/s = {balg/res[0]/1e3/6573.2:.3f} of HBM peak", flush=True)
        del sp
