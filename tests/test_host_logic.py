"""CPU tests of the product's host logic: the C-ABI library loads and exports what the header
declares, the tile plan is consistent (re-enacted in numpy against the oracle), mesh helpers,
error paths, and loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from _helpers import GOLDEN, ROOT, COracle, build_host_plan, emulate_kernel
from tssplat_b200 import _capi
from tssplat_b200.mesh import (concat_spheres, connected_components, load_veg, make_pack, make_tet_sphere, perturb,
                               save_veg)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tssplat_b200.h")).read()
    declared = set(re.findall(r"\b(tsb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(_capi.lib, sym), sym


def test_info_struct_matches_header():
    hdr = open(os.path.join(ROOT, "include", "tssplat_b200.h")).read()
    body = hdr[hdr.index("typedef struct {\n  int32_t n;"):hdr.index("} tsb_info_t;")]
    fields = re.findall(r"int(?:32|64)_t\s+([a-z_]+);", body)
    assert fields == [f for f, _ in _capi.tsb_info_t._fields_]


@pytest.mark.parametrize("tile_tets,balance", [(256, 0), (512, 0), (512, 5), (1024, 0)])
def test_plan_reenactment_matches_oracle(tile_tets, balance):
    pack = make_pack(3, 768, seed=2)
    plan = build_host_plan(pack.verts, pack.tets, tile_tets, balance_sms=balance)
    assert sorted(plan["tet_order"].tolist()) == list(range(pack.nele))      # every tet exactly once
    assert sum(int(t["ntet"]) for t in plan["tiles"]) == pack.nele
    assert plan["n_components"] == 3
    orc = COracle(pack.verts, pack.tets)
    for sig, order in ((0.02, 2), (0.35, 4)):
        x = perturb(pack, sigma_rel=sig, seed=1)
        E, es, eb, g = emulate_kernel(plan, x, 2e-4, 3e-4, order, gradH=0.7)
        Eo, terms, go = orc.energy_grad(x, 2e-4, 3e-4, order, gradH=0.7)
        assert E == pytest.approx(Eo, rel=2e-6)          # fp32 rest inverses, fp64 arithmetic
        assert np.linalg.norm(g - go) <= 2e-6 * np.linalg.norm(go)


def test_plan_laplacian_scale_and_unreferenced_vertices():
    v, t = make_tet_sphere(1201, 300)
    v = np.concatenate([v, [[5.0, 5.0, 5.0], [6.0, 6.0, 6.0]]])          # two vertices no tet uses
    plan = build_host_plan(v, t, 256, laplacian_scale=1)
    x = perturb(v, t, 0.3, 7)
    E, _, _, g = emulate_kernel(plan, x, 1e-3, 1e-3, 2)
    Eo, _, go = COracle(v, t, 1).energy_grad(x, 1e-3, 1e-3, 2)
    assert E == pytest.approx(Eo, rel=2e-6)
    assert np.all(g[-2:] == 0.0) and np.all(go[-2:] == 0.0)
    assert np.linalg.norm(g - go) <= 2e-6 * np.linalg.norm(go)


def test_plan_real_mesh_a_veg():
    d = np.load(os.path.join(GOLDEN, "a_veg_mesh.npz"))
    plan = build_host_plan(d["verts"], d["tets"], 512, balance_sms=0)
    x = perturb(d["verts"], d["tets"], 0.35, 1)
    E, es, eb, g = emulate_kernel(plan, x, 3.2e-3, 3.2e-3, 2, gradH=0.5)
    gold = np.load(os.path.join(GOLDEN, "golden_energy.npz"))
    assert E == pytest.approx(float(gold["a_veg/inverted_o2/energy"]), rel=2e-6)
    assert np.linalg.norm(g) == pytest.approx(float(gold["a_veg/inverted_o2/grad_l2"]), rel=2e-6)


def test_tiny_meshes():
    """One tet (no neighbours: smoothness identically 0) and two tets sharing a face."""
    v1 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
    t1 = np.array([[0, 1, 2, 3]], dtype=np.int32)
    x = (v1 * np.array([1, 1, -1.5])).astype(np.float32)
    E, es, eb, g = emulate_kernel(build_host_plan(v1, t1, 256), x, 1.0, 1.0, 2)
    assert es == 0.0 and eb == pytest.approx(1.5 ** 2)
    v2 = np.concatenate([v1, [[1.0, 1.0, 1.0]]])
    t2 = np.array([[0, 1, 2, 3], [1, 2, 3, 4]], dtype=np.int32)
    t2[1] = [1, 3, 2, 4]                                                   # positive orientation
    x2 = perturb(v2, t2, 0.3, 1)
    E2, _, _, g2 = emulate_kernel(build_host_plan(v2, t2, 256), x2, 0.7, 0.3, 2)
    Eo, _, go = COracle(v2, t2).energy_grad(x2, 0.7, 0.3, 2)
    assert E2 == pytest.approx(Eo, rel=1e-5) and np.abs(g2 - go).max() < 1e-5 * np.abs(go).max()


def _plan_error(v, t):
    with pytest.raises(RuntimeError) as ei:
        build_host_plan(np.asarray(v, dtype=np.float64), np.asarray(t, dtype=np.int32), 256)
    return str(ei.value)


def test_plan_rejects_bad_meshes():
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], dtype=np.float64)
    assert "out of range" in _plan_error(v, [[0, 1, 2, 7]])
    assert "zero rest volume" in _plan_error(np.zeros((4, 3)), [[0, 1, 2, 3]])
    assert "repeats a vertex" in _plan_error(v, [[0, 1, 1, 3]])
    three = [[0, 1, 2, 3], [0, 2, 1, 4], [0, 1, 2, 4]]                     # face (0,1,2) used three times
    vv = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
    assert "non-manifold" in _plan_error(vv, three)


def test_unsupported_tile_size():
    v, t = make_tet_sphere(1202, 64)
    with pytest.raises(RuntimeError):
        build_host_plan(v, t, 300)


def test_veg_round_trip(tmp_path):
    v, t = make_tet_sphere(1203, 200)
    p = str(tmp_path / "m.veg")
    save_veg(p, v, t)
    v2, t2 = load_veg(p)
    assert np.array_equal(t, t2) and np.abs(v - v2).max() < 1e-14
    if os.path.exists("/root/reference/tssplat_ext/a.veg"):               # only in the build container
        va, ta = load_veg("/root/reference/tssplat_ext/a.veg")
        d = np.load(os.path.join(GOLDEN, "a_veg_mesh.npz"))
        assert np.array_equal(ta, d["tets"]) and np.array_equal(va, d["verts"])


def test_synthetic_pack_properties():
    from oracle.tet_energy_oracle import face_adjacency
    from tssplat_b200.mesh import _signed_volumes
    pk = make_pack(4, 512, seed=9)
    assert pk.nele == 4 * 512 and pk.num_spheres == 4
    assert np.all(_signed_volumes(pk.verts.astype(np.float64), pk.tets.astype(np.int64)) > 0)
    face_adjacency(pk.tets)                                                # manifold (raises otherwise)
    lab = connected_components(pk.n, pk.tets)
    for s in range(4):
        assert len(set(lab[pk.vert_offsets[s]:pk.vert_offsets[s + 1]].tolist())) == 1
    assert len(set(lab.tolist())) == 4
    assert pk.algorithmic_bytes() == 24 * pk.n + 68 * pk.nele
    sub = pk.slice_spheres(1, 3)
    assert sub.nele == 1024 and sub.tets.min() == 0 and sub.tets.max() == sub.n - 1


def test_product_fails_loudly_without_gpu():
    import torch
    from tssplat_b200 import tet_spheres_ext as ext
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    v, t = make_tet_sphere(1204, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.TetSpheres(v.astype(np.float32).reshape(-1), t.reshape(-1))
    h = C.c_void_p()
    vf = np.ascontiguousarray(v, dtype=np.float32)
    rc = _capi.lib.tsb_create(vf.ctypes.data, t.ctypes.data, len(v), len(t), None, 0, C.byref(h))
    assert rc == _capi.TSB_E_CUDA and not h.value                           # no CPU fallback
    with pytest.raises(RuntimeError):
        ext.TetSpheres(np.zeros(9), np.zeros(4, dtype=np.int32))            # wrong dtype (float64)


def test_drop_in_import_and_scheduler():
    """`from tet_spheres import tet_spheres_ext` (energies/smooth_barrier.py:6) and the coefficient
    scheduler / order switch (energies/smooth_barrier.py:47-66)."""
    import math
    from tet_spheres import tet_spheres_ext
    for name in ("TetSpheres", "forward", "backward", "random_x", "grad_limit"):
        assert hasattr(tet_spheres_ext, name)
    from tssplat_b200.energies import SmoothnessBarrierEnergy
    eng = SmoothnessBarrierEnergy.__new__(SmoothnessBarrierEnergy)
    from types import SimpleNamespace
    eng.FLAGS = SimpleNamespace(smooth_eng_coeff=2e-4 / 64, barrier_coeff=2e-4, increase_order_iter=1000)
    assert eng.coeff_scheduler(0) == pytest.approx((2e-4 / 64, 2e-4))
    c1, c2 = eng.coeff_scheduler(600)
    m = 2 ** (4 * abs(math.sin(600 / 2400 * math.pi)))
    assert c1 == pytest.approx(2e-4 / 64 * m) and c2 == pytest.approx(2e-4 * m)
    assert eng.coeff_scheduler(5000) == pytest.approx((2e-4 / 64 * 16, 2e-4 * 16))
