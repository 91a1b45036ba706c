"""Test-side helpers: the C oracle binding, plan inspection, and a numpy emulation of the CUDA
kernel's tile algorithm (CPU checks of the host logic only -- never a product path)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_SO = os.path.join(ROOT, "oracle", "libtet_energy_oracle.so")


class COracle:
    """oracle/tet_energy_oracle.c through ctypes."""

    def __init__(self, rest, tets, laplacian_scale=0):
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.tso_create.restype = C.c_void_p
        self.lib.tso_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.tso_destroy.argtypes = [C.c_void_p]
        self.lib.tso_energy_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                             C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        self.rest = np.ascontiguousarray(np.asarray(rest, dtype=np.float32).reshape(-1, 3))
        self.tets = np.ascontiguousarray(np.asarray(tets, dtype=np.int32).reshape(-1, 4))
        self.n, self.nele = len(self.rest), len(self.tets)
        self.h = self.lib.tso_create(self.rest.ctypes.data, self.tets.ctypes.data, self.n, self.nele,
                                     int(laplacian_scale))
        if not self.h:
            raise ValueError("C oracle rejected the mesh")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.tso_destroy(self.h)
            self.h = None

    def energy_grad(self, x, c1, c2, order, gradH=1.0, nthreads=0, want_grad=True):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1, 3))
        terms = np.zeros(2)
        g = np.zeros((self.n, 3)) if want_grad else None
        self.lib.tso_energy_grad(self.h, x.ctypes.data, float(c1), float(c2), int(order), float(gradH),
                                 terms.ctypes.data, g.ctypes.data if want_grad else None, int(nthreads))
        return float(c1) * terms[0] + float(c2) * terms[1], terms, g


_TILE_DT = np.dtype([(k, np.int32) for k in
                     ("ntet", "nvert", "vert_off", "ngrp", "grp_off", "ell_off", "cg_off", "ncg")])
_ARRAYS = {"tiles": _TILE_DT, "idx8": np.uint16, "Bsoa": np.float32, "vlist": np.int32, "Xloc": np.float32,
           "dest": np.int32, "ell": np.uint16, "ell_grp_ptr": np.int32, "cg_list": np.int32, "need": np.int32,
           "gsv_ptr": np.int32, "sv_vid": np.int32, "sv_slot_ptr": np.int32, "tet_order": np.int32}


def build_host_plan(rest, tets, tile_tets=512, laplacian_scale=0):
    """Run the product's host plan builder (no CUDA) and copy its arrays out as numpy."""
    from tssplat_b200 import _capi
    lib = _capi.lib
    lib.tsb_debug_plan_build.restype = C.c_int
    lib.tsb_debug_plan_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.POINTER(C.c_void_p)]
    lib.tsb_debug_plan_array.restype = C.c_int
    lib.tsb_debug_plan_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int32)]
    lib.tsb_debug_plan_scalars.argtypes = [C.c_void_p, C.c_void_p]
    lib.tsb_debug_plan_free.argtypes = [C.c_void_p]
    rest = np.ascontiguousarray(np.asarray(rest, dtype=np.float32).reshape(-1))
    tets = np.ascontiguousarray(np.asarray(tets, dtype=np.int32).reshape(-1))
    d = C.c_void_p()
    rc = lib.tsb_debug_plan_build(rest.ctypes.data, tets.ctypes.data, rest.size // 3, tets.size // 4,
                                  int(tile_tets), int(laplacian_scale), C.byref(d))
    if rc != 0:
        raise RuntimeError(_capi.last_error(None))
    try:
        plan = {}
        for name, dt in _ARRAYS.items():
            ptr, cnt, eb = C.c_void_p(), C.c_int64(), C.c_int32()
            assert lib.tsb_debug_plan_array(d, name.encode(), C.byref(ptr), C.byref(cnt), C.byref(eb)) == 0, name
            nbytes = cnt.value * eb.value
            buf = (C.c_char * nbytes).from_address(ptr.value) if nbytes else b""
            plan[name] = np.frombuffer(bytes(buf), dtype=dt).copy()
        sc = np.zeros(8, np.int32)
        lib.tsb_debug_plan_scalars(d, sc.ctypes.data)
        for k, v in zip(("n", "nele", "tile_tets", "max_local_vertices", "n_tiles", "n_components",
                         "n_shared_vertices", "n_slots"), sc):
            plan[k] = int(v)
        plan["laplacian_scale"] = int(laplacian_scale)
    finally:
        lib.tsb_debug_plan_free(d)
    return plan


def emulate_kernel(plan, x, c1, c2, order, gradH=1.0, dtype=np.float64):
    """numpy re-enactment of tsb_kernels.cu's four phases on the host plan (same formulas, same data
    structures, tile by tile).  Returns (energy_total, smooth, barrier, grad[n,3])."""
    TT = plan["tile_tets"]
    x = np.asarray(x, dtype=np.float32).reshape(-1, 3).astype(dtype)
    n = plan["n"]
    grad = np.full((n, 3), np.nan, dtype=dtype)
    scratch = np.full((max(plan["n_slots"], 1), 3), np.nan, dtype=dtype)
    idx8 = plan["idx8"].reshape(-1, 8)
    Bs = plan["Bsoa"].reshape(-1, 9, TT)
    es_tot = eb_tot = 0.0
    for ti, td in enumerate(plan["tiles"]):
        nt, nv, vo = int(td["ntet"]), int(td["nvert"]), int(td["vert_off"])
        vl = plan["vlist"][vo:vo + nv]
        xs = x[vl]                                              # phase 0
        Xs = plan["Xloc"].reshape(-1, 3)[vo:vo + nv].astype(dtype)
        ids = idx8[ti * TT: ti * TT + nt].astype(np.int64)
        own, opp = ids[:, :4], ids[:, 4:]
        assert own.max() < nv
        B = np.transpose(Bs[ti, :, :nt], (1, 0)).reshape(nt, 3, 3).astype(dtype)   # rows a1..a3
        a = np.concatenate([-B.sum(axis=1, keepdims=True), B], axis=1)              # nt x 4 x 3
        p = xs[own]                                             # nt x 4 x 3
        e = p[:, 1:, :] - p[:, :1, :]                           # e[j][r]
        F = np.einsum("tjr,tjc->trc", e, a[:, 1:, :])
        J = np.linalg.det(F)
        inv = J < 0
        m = np.where(inv, -J, 0.0)
        if order == 2:
            eb = m * m; coef = 2 * m
        else:
            eb = m ** 4; coef = 4 * m ** 3
        cof = np.linalg.inv(np.where(np.abs(J)[:, None, None] > 0, F, np.eye(3)))
        cof = np.transpose(cof, (0, 2, 1)) * J[:, None, None]   # cof(F) = det(F) F^-T
        Pm = (-c2 * coef)[:, None, None] * cof
        Pm[~inv] = 0
        z = np.einsum("trc,tjc->tjr", Pm, a)                    # nt x 4 x 3
        H = np.zeros((nt, 3, 3), dtype=dtype)
        lam = np.zeros((nt, 4, 4), dtype=dtype)
        rho = np.zeros((nt, 4), dtype=dtype)
        valid = opp != 0xFFFF
        deg = valid.sum(axis=1)
        d_all = np.zeros((nt, 4, 3), dtype=dtype)
        for k in range(4):
            v = valid[:, k]
            ok = np.where(v, opp[:, k], 0)
            r = Xs[ok] - Xs[own[:, 0]]
            l123 = np.einsum("tjc,tc->tj", B, r)
            l0 = 1 - l123.sum(axis=1)
            lk = np.concatenate([l0[:, None], l123], axis=1)
            lam[:, k, :] = np.where(v[:, None], lk, 0)
            lkk = lk[:, k]
            rho[:, k] = np.where(v, -1.0 / np.where(v, lkk, 1.0), 0.0)
            d = (xs[ok] - p[:, 0, :]) - np.einsum("tj,tjr->tr", l123, e)
            d_all[:, k, :] = np.where(v[:, None], d, 0)
            H += (rho[:, k, None] * d_all[:, k, :])[:, :, None] * a[:, k, None, :]
        w = np.where(plan["laplacian_scale"] & (deg > 0), 1.0 / np.maximum(deg, 1), 1.0)
        H *= w[:, None, None]
        es = 0.5 * (H * H).sum(axis=(1, 2))
        y = (c1 * w)[:, None, None] * rho[:, :, None] * np.einsum("trc,tkc->tkr", H, a)   # nt x 4 x 3
        y = np.where(valid[:, :, None], y, 0)
        z = z - np.einsum("tkj,tkr->tjr", lam, y)
        outb = np.full((24, TT), np.nan, dtype=dtype)           # phase 1 table
        for j in range(4):
            for r in range(3):
                outb[j * 3 + r, :nt] = z[:, j, r]
                col = np.where(valid[:, j], y[:, j, r], np.nan)
                outb[(4 + j) * 3 + r, :nt] = col
        es_tot += es.sum(); eb_tot += eb.sum()
        # phase 2: sliced-ELL gather
        gp = plan["ell_grp_ptr"][td["grp_off"]: td["grp_off"] + td["ngrp"] + 1]
        acc = np.zeros((td["ngrp"] * 32, 3), dtype=dtype)
        for g in range(td["ngrp"]):
            blk = plan["ell"][td["ell_off"] + gp[g]: td["ell_off"] + gp[g + 1]].reshape(-1, 32).astype(np.int64)
            for lane in range(32):
                ent = blk[:, lane]
                ent = ent[ent != 0xFFFF]
                for c in range(3):
                    acc[g * 32 + lane, c] = outb[(ent & 7) * 3 + c, ent >> 3].sum()
        dest = plan["dest"][vo:vo + nv]
        for pidx in range(nv):
            dd = int(dest[pidx])
            if dd >= 0:
                assert np.isnan(grad[dd, 0]), "exclusive vertex written twice"
                grad[dd] = gradH * acc[pidx]
            else:
                assert np.isnan(scratch[-1 - dd, 0]), "scratch slot written twice"
                scratch[-1 - dd] = acc[pidx]
    # phase 3: combine shared vertices
    for o in range(plan["n_tiles"]):
        for sv in range(plan["gsv_ptr"][o], plan["gsv_ptr"][o + 1]):
            s0, s1 = plan["sv_slot_ptr"][sv], plan["sv_slot_ptr"][sv + 1]
            vid = int(plan["sv_vid"][sv])
            assert np.isnan(grad[vid, 0]), "shared vertex also written as exclusive"
            grad[vid] = gradH * scratch[s0:s1].sum(axis=0)
    return c1 * es_tot + c2 * eb_tot, es_tot, eb_tot, grad
