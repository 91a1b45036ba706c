"""Which data being HBM-cold costs the 64 x 4096 step its ~1.5 us: the plan stream, or x / grad?
A: one handle, one x (everything L2-warm).  B: 9 handles rotating, one shared x/grad (plan cold).
C: one handle, 220 rotating x/grad buffers (x/grad cold, plan warm).  D: both rotating (bench.py's setup)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tssplat_b200 import _capi  # noqa: E402
from tssplat_b200 import tet_spheres_ext as ext  # noqa: E402
from tssplat_b200.mesh import make_pack, perturb  # noqa: E402

S, NH, NX = 64, 9, 220
KW = {"warps_per_cta": 8, "ring_slots": 2} if len(sys.argv) > 1 and sys.argv[1] == "nw8" else {}
pack = make_pack(S, 4096, seed=0, unique=8)
hs = [ext.TetSpheres(pack.verts.reshape(-1), pack.tets.reshape(-1), **KW) for _ in range(NH)]
print(KW, {k: hs[0].info[k] for k in ("grid", "warps_per_cta", "ctas_per_sm", "smem_bytes", "ring_slots")})
x0 = torch.from_numpy(perturb(pack, sigma_rel=0.02, seed=0)).cuda()
xs = [x0.clone() for _ in range(NX)]
gs = [torch.empty_like(x0) for _ in range(NX)]
en = torch.zeros((NX, 3), device="cuda")
st = torch.cuda.Stream()
c1, c2 = 2e-4 / S, 2e-4


def run(label, hsel, xsel, count):
    def launch(k):
        h, j = hs[hsel(k)], xsel(k)
        rc = _capi.lib.tsb_energy_grad(h._h, xs[j].data_ptr(), c1, c2, 2, 1.0, None, en[j].data_ptr(), gs[j].data_ptr(), st.cuda_stream)
        assert rc == 0
    with torch.cuda.stream(st):
        for k in range(3):
            launch(k)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for k in range(count):
                launch(k)
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, 20000 // count)
        e0.record(st)
        for _ in range(reps):
            g.replay()
        e1.record(st)
        st.synchronize()
    print(f"  {label}: {e0.elapsed_time(e1) * 1e3 / (reps * count):.2f} us/step", flush=True)


for _ in range(1):
    run("A all warm                 ", lambda k: 0, lambda k: 0, 90)
    run("B plan cold (9 handles)    ", lambda k: k % NH, lambda k: 0, 90)
    run("C x/grad cold (220 buffers)", lambda k: 0, lambda k: k % NX, NX)
    run("D both cold                ", lambda k: k % NH, lambda k: k % NX, 3 * NX)
