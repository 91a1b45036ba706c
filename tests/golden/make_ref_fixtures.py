"""Generates tests/golden/ref_fixtures.npz from REFERENCE-HELD code imported from /root/reference
(run in the build container only; the GPU box never reads /root/reference):

* geometry/mesh_utils.py:38-69  compute_G_matrix   -> per-tet deformation gradients F = G x and det F
* geometry/mesh_utils.py:5-35   get_surface_vf     -> surface vertex ids + surface triangles
* utils/optimizer.py:37-89      AdamUniform        -> a 10-step parameter trajectory incl. the grad_limit schedule

These pin the restated pieces (oracle F / det F, the product's barrier term, the AdamUniform kernels, the
surface extraction) to the reference's own Python.  Only the tet Laplacian L (libpgo) stays an assumption.

    python tests/golden/make_ref_fixtures.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from tssplat_b200.mesh import make_pack, perturb
    mu = _load("ref_mesh_utils", "geometry/mesh_utils.py")
    opt = _load("ref_optimizer", "utils/optimizer.py")
    out = {}

    d = np.load(os.path.join(HERE, "a_veg_mesh.npz"))
    pk = make_pack(3, 1024, seed=1)
    meshes = {"a_veg": (d["verts"].astype(np.float32), d["tets"]), "pack3x1024": (pk.verts, pk.tets)}
    for name, (v32, t) in meshes.items():
        # the reference widens float32 rest positions to double (tet_spheres.cpp:251-254)
        V = v32.astype(np.float64)
        G = mu.compute_G_matrix(V, t.astype(np.int64))                   # T x 9 x 12, reference code
        for case, (sig, seed) in {"benign": (0.02, 0), "inverted": (0.35, 1)}.items():
            x = perturb(v32, t, sig, seed)                               # float32 positions, committed generator
            xl = x.astype(np.float64)[t].reshape(len(t), 12)             # tet-local dofs (v0xyz, v1xyz, ...)
            F = np.einsum("tij,tj->ti", G, xl)                           # T x 9, row-major 3x3 (mesh_utils.py:66-67)
            det = np.linalg.det(F.reshape(-1, 3, 3))
            k = f"{name}/{case}"
            out[k + "/x"] = x
            out[k + "/detF"] = det
            out[k + "/F_sample"] = F[:: max(1, len(t) // 512)][:512]
            out[k + "/barrier_o2"] = np.sum(np.maximum(-det, 0.0) ** 2)  # tet_spheres_cuda.cu:48-66 on the reference's F
            out[k + "/barrier_o4"] = np.sum(np.maximum(-det, 0.0) ** 4)
        sv, sf = mu.get_surface_vf(t.astype(np.int64))
        out[name + "/surface_vid"] = sv.astype(np.int64)
        out[name + "/surface_f"] = sf.astype(np.int64)

    # vertex normals by EXECUTING the reference's own function body (geometry/tetmesh_geometry.py:39-66; the module
    # itself cannot be imported: it pulls in pypgo/trimesh at import time)
    import ast
    import types
    import torch.nn.functional as F_
    src = open(os.path.join(REF, "geometry/tetmesh_geometry.py")).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TetMeshGeometryForwardData")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "_compute_vertex_normal")
    ns = {"torch": torch, "F": F_, "dot": lambda a, b: torch.sum(a * b, -1, keepdim=True)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref:_compute_vertex_normal", "exec"), ns)
    for name, (v32, t) in meshes.items():
        sv, sf = out[name + "/surface_vid"], out[name + "/surface_f"]
        x = torch.from_numpy(out[name + "/inverted/x"])
        obj = types.SimpleNamespace(v_pos=x[torch.from_numpy(sv)], t_pos_idx=torch.from_numpy(sf))   # tetmesh_geometry.py:33
        out[name + "/v_nrm"] = ns["_compute_vertex_normal"](obj).numpy()

    # AdamUniform trajectory (CPU float32 torch, exactly the reference class)
    torch.manual_seed(0)
    n = 600
    p = torch.nn.Parameter(torch.randn(n, 3))
    p0 = p.detach().clone()
    o = opt.AdamUniform([p], grad_limit=True, grad_limit_values=[0.05, 0.01], grad_limit_iters=[4], lr=0.2, betas=(0.9, 0.999))
    grads, traj = [], []
    for step in range(10):
        g = torch.randn(n, 3) * (5.0 if step % 3 == 0 else 0.02)
        p.grad = g.clone()
        o.step()
        grads.append(g.numpy().copy())
        traj.append(p.detach().numpy().copy())
    st = o.state[p]
    out["adam/p0"] = p0.numpy()
    out["adam/grads"] = np.stack(grads)
    out["adam/traj"] = np.stack(traj)
    out["adam/g1"] = st["g1"].numpy()
    out["adam/g2"] = st["g2"].numpy()
    out["adam/hyper"] = np.array([0.2, 0.9, 0.999, 0.05, 0.01, 4.0])
    np.savez_compressed(os.path.join(HERE, "ref_fixtures.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_fixtures.npz"), {k: np.asarray(v).shape for k, v in list(out.items())[:6]})


if __name__ == "__main__":
    main()
