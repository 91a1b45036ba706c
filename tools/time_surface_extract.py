import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from tssplat_b200.mesh import make_pack, surface_vf, surface_vf_gpu
for S in (64, 1024):
    pk = make_pack(S, 4096, seed=0, unique=8)
    surface_vf_gpu(pk.tets[:4096], pk.n)
    t0 = time.perf_counter(); a = surface_vf_gpu(pk.tets, pk.n); t1 = time.perf_counter(); b = surface_vf(pk.tets); t2 = time.perf_counter()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    print(f"S={S}: surface extraction GPU {1e3*(t1-t0):.1f} ms (incl. copies), numpy {1e3*(t2-t1):.1f} ms; {len(a[0])} surface vertices, {len(a[1])} faces")
