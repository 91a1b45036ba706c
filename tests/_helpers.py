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


_ARRAYS = {"vblob": np.uint8, "tblob": np.uint8, "ell": np.uint16, "slot_ptr": np.int32,
           "tet_order": np.int32, "tile_first": np.int32}
_HDR = ("ntet", "nvert", "nrow", "ell_off", "nell")
ROW_CAP = 16


def rows_cap(tt, nv):
    return nv + 8 * tt // ROW_CAP


def vblob_bytes(tt, nv):
    nr = rows_cap(tt, nv)
    return 64 + 16 * nv + 4 * nr + 4 * (nr // 32 + 4)


def build_host_plan(rest, tets, tile_tets=512, laplacian_scale=0, balance_sms=0):
    """Run the product's host plan builder (no CUDA) and copy its arrays out as numpy."""
    from tssplat_b200 import _capi
    lib = _capi.lib
    lib.tsb_debug_plan_build.restype = C.c_int
    lib.tsb_debug_plan_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.POINTER(C.c_void_p)]
    lib.tsb_debug_plan_array.restype = C.c_int
    lib.tsb_debug_plan_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int32)]
    lib.tsb_debug_plan_scalars.argtypes = [C.c_void_p, C.c_void_p]
    lib.tsb_debug_plan_free.argtypes = [C.c_void_p]
    rest = np.ascontiguousarray(np.asarray(rest, dtype=np.float32).reshape(-1))
    tets = np.ascontiguousarray(np.asarray(tets, dtype=np.int32).reshape(-1))
    d = C.c_void_p()
    rc = lib.tsb_debug_plan_build(rest.ctypes.data, tets.ctypes.data, rest.size // 3, tets.size // 4,
                                  int(tile_tets), int(laplacian_scale), int(balance_sms), C.byref(d))
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
        sc = np.zeros(10, np.int32)
        lib.tsb_debug_plan_scalars(d, sc.ctypes.data)
        for k, v in zip(("n", "nele", "tile_tets", "max_local_vertices", "n_tiles", "n_components",
                         "n_shared_vertices", "n_slots", "fill", "ell_cap"), sc):
            plan[k] = int(v)
        plan["laplacian_scale"] = int(laplacian_scale)
    finally:
        lib.tsb_debug_plan_free(d)
    # unpack the per-tile blobs
    TT, NV = plan["tile_tets"], plan["max_local_vertices"]
    VB, TB, NR = vblob_bytes(TT, NV), 52 * TT, rows_cap(TT, NV)
    tiles = []
    for t in range(plan["n_tiles"]):
        vb = plan["vblob"][t * VB:(t + 1) * VB]
        tb = plan["tblob"][t * TB:(t + 1) * TB]
        hdr = dict(zip(_HDR, vb[:20].view(np.int32)))
        nv, nt = int(hdr["nvert"]), int(hdr["ntet"])
        tiles.append(dict(
            hdr, vlist=vb[64:64 + 4 * NV].view(np.int32)[:nv],
            X=np.stack([vb[64 + 4 * NV:64 + 8 * NV].view(np.float32)[:nv],
                        vb[64 + 8 * NV:64 + 16 * NV].view(np.float32).reshape(-1, 2)[:nv, 0],
                        vb[64 + 8 * NV:64 + 16 * NV].view(np.float32).reshape(-1, 2)[:nv, 1]], axis=1),
            slot=vb[64 + 16 * NV:64 + 16 * NV + 4 * NR].view(np.int32)[:int(hdr["nrow"])],
            grp_ptr=vb[64 + 16 * NV + 4 * NR:].view(np.int32)[:(int(hdr["nrow"]) + 31) // 32 + 1],
            idx8=tb[:16 * TT].view(np.uint16).reshape(-1, 8)[:nt],
            B=tb[16 * TT:].view(np.float32).reshape(-1, 9)[:nt]))
    plan["tiles"] = tiles
    return plan


def emulate_kernel(plan, x, c1, c2, order, gradH=1.0, dtype=np.float64):
    """numpy re-enactment of tsb_kernels.cu (tile kernel phases 0-2 + combine kernel) on the host
    plan: same formulas, same data structures, tile by tile.
    Returns (energy_total, smooth, barrier, grad[n,3])."""
    TT = plan["tile_tets"]
    TTP = TT + 4
    x = np.asarray(x, dtype=np.float32).reshape(-1, 3).astype(dtype)
    n = plan["n"]
    scratch = np.full((max(plan["n_slots"], 1), 3), np.nan, dtype=dtype)
    es_tot = eb_tot = 0.0
    for ti, td in enumerate(plan["tiles"]):
        nt, nv = int(td["ntet"]), int(td["nvert"])
        vl = td["vlist"]
        xs = x[vl]                                              # phase 0
        Xs = td["X"].astype(dtype)
        ids = td["idx8"].astype(np.int64)
        own, oppr = ids[:, :4], ids[:, 4:]
        valid = (oppr & 0x8000) != 0
        opp = oppr & 0x7FFF
        assert own.max() < nv and opp.max() < nv
        assert np.all(opp[~valid] == own[~valid]), "boundary faces must point at the own vertex"
        B = td["B"].reshape(nt, 3, 3).astype(dtype)             # rows a1..a3
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
        deg = valid.sum(axis=1)
        d_all = np.zeros((nt, 4, 3), dtype=dtype)
        for k in range(4):
            v = valid[:, k]
            ok = opp[:, k]
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
        outb = np.full((24, TTP), np.nan, dtype=dtype)          # phase 1 table
        outb[0:3, TT] = 0.0                                     # zero column
        for j in range(4):
            for r in range(3):
                outb[j * 3 + r, :nt] = z[:, j, r]
                outb[(4 + j) * 3 + r, :nt] = y[:, j, r]
        es_tot += es.sum(); eb_tot += eb.sum()
        # phase 2: sliced-ELL gather (entries are word offsets into the [24][TT+4] table, stored as
        # [k/2][lane][2] pairs; padding points at the zero column)
        gp = td["grp_ptr"]
        nrow = int(td["nrow"])
        ngrp = (nrow + 31) // 32
        flat = outb.reshape(-1)
        ell = plan["ell"][td["ell_off"]: td["ell_off"] + td["nell"]]
        n_real = 0
        for r in range(nrow):
            g, lane = r >> 5, r & 31
            blk = ell[gp[g]: gp[g + 1]].reshape(-1, 32, 2).astype(np.int64)
            assert blk.shape[0] * 2 <= ROW_CAP
            ent = blk[:, lane, :].reshape(-1)
            n_real += int((ent != TT).sum())
            sl = int(td["slot"][r])
            assert np.isnan(scratch[sl, 0]), "scratch slot written twice"
            scratch[sl] = [flat[ent + c * TTP].sum() for c in range(3)]
        assert n_real == int(4 * nt + valid.sum()), "gather table must list every (tet, slot) exactly once"
    # combine kernel
    sp = plan["slot_ptr"]
    assert sp[0] == 0 and sp[-1] == plan["n_slots"] and not np.isnan(scratch[:plan["n_slots"]]).any()
    grad = np.stack([gradH * scratch[sp[v]:sp[v + 1]].sum(axis=0) for v in range(n)])
    return c1 * es_tot + c2 * eb_tot, es_tot, eb_tot, grad
