"""Test-side helpers: the C oracle binding, plan inspection, and a numpy emulation of the CUDA
kernel's tile algorithm (CPU checks of the host logic only -- never a product path)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_SO = os.path.join(ROOT, "oracle", "libtet_energy_oracle.so")


def host_has_avx2_fma() -> bool:
    try:
        flags = open("/proc/cpuinfo").read()
        return " avx2" in flags and " fma" in flags
    except OSError:
        return False


class COracle:
    """oracle/tet_energy_oracle.c through ctypes.  variant: "" = the fp64 checker; "fast" / "fast32" = the
    AVX2 timing builds (fp64 / fp32 arithmetic) used by bench.py's CPU arms only."""

    def __init__(self, rest, tets, laplacian_scale=0, variant=""):
        so = ORACLE_SO if not variant else ORACLE_SO.replace(".so", f"_{variant}.so")
        self.lib = C.CDLL(so)
        self.lib.tso_create.restype = C.c_void_p
        self.lib.tso_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.tso_destroy.argtypes = [C.c_void_p]
        self.lib.tso_energy_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                             C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        self.lib.tso_energy_grad_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int,
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

    def energy_grad_ex(self, x, c1, c2, c3, order, gradH=1.0, nthreads=0, want_grad=True):
        """With the AMIPS term (c3): returns (total, terms[3] = smooth/barrier/amips, grad)."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1, 3))
        terms = np.zeros(3)
        g = np.zeros((self.n, 3)) if want_grad else None
        self.lib.tso_energy_grad_ex(self.h, x.ctypes.data, float(c1), float(c2), float(c3), int(order), float(gradH),
                                    terms.ctypes.data, g.ctypes.data if want_grad else None, int(nthreads))
        return float(c1) * terms[0] + float(c2) * terms[1] + float(c3) * terms[2], terms, g


PLAN_DEBUG_SO = os.path.join(ROOT, "tests", "native", "libtsb_plan_debug.so")
_ARRAYS = {"stream": np.uint8, "X4": np.float32, "vlist": np.int32, "segs": np.int32, "cta_seg": np.int32,
           "wdesc": np.uint32, "wseg": np.uint16, "orphans": np.int32, "pos16": np.uint16, "pos_gid": np.int32}
_SCALARS = ("n", "nele", "n_components", "n_boundary_faces", "laplacian_scale", "mode_global", "nw", "grid", "vh",
            "area_verts", "max_comp_verts", "contiguous", "nnz", "nnz_padded", "n_rb", "n_tetcells",
            "gather_wf", "gather_wf_ideal", "tet_wf", "tet_wf_ideal")
_SEG = ("comp", "vbase", "nv", "x4off", "expected", "whole", "npos", "p4off")


def build_host_plan(rest, tets, nw=16, grid=148, laplacian_scale=0, force_global=0, vh_cap=0, area_cap=0,
                    tet_cost=0.0):
    """Run the product's host plan builder (tssplat_b200/csrc/tsb_plan.cpp, no CUDA) through the
    test-only inspection library and copy its arrays out as numpy."""
    lib = C.CDLL(PLAN_DEBUG_SO)
    lib.tsbdbg_build.restype = C.c_int
    lib.tsbdbg_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_int32, C.c_float, C.POINTER(C.c_void_p)]
    lib.tsbdbg_array.restype = C.c_int
    lib.tsbdbg_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int32)]
    lib.tsbdbg_scalars.argtypes = [C.c_void_p, C.c_void_p]
    lib.tsbdbg_free.argtypes = [C.c_void_p]
    lib.tsbdbg_last_error.restype = C.c_char_p
    rest = np.ascontiguousarray(np.asarray(rest, dtype=np.float32).reshape(-1))
    tets = np.ascontiguousarray(np.asarray(tets, dtype=np.int32).reshape(-1))
    d = C.c_void_p()
    rc = lib.tsbdbg_build(rest.ctypes.data, tets.ctypes.data, rest.size // 3, tets.size // 4, int(nw), int(grid),
                          int(laplacian_scale), int(force_global), int(vh_cap), int(area_cap), float(tet_cost),
                          C.byref(d))
    if rc != 0:
        raise RuntimeError(lib.tsbdbg_last_error().decode())
    try:
        plan = {}
        for name, dt in _ARRAYS.items():
            ptr, cnt, eb = C.c_void_p(), C.c_int64(), C.c_int32()
            assert lib.tsbdbg_array(d, name.encode(), C.byref(ptr), C.byref(cnt), C.byref(eb)) == 0, name
            nbytes = cnt.value * eb.value
            buf = (C.c_char * nbytes).from_address(ptr.value) if nbytes else b""
            plan[name] = np.frombuffer(bytes(buf), dtype=dt).copy()
        sc = np.zeros(20, np.int64)
        lib.tsbdbg_scalars(d, sc.ctypes.data)
        for k, v in zip(_SCALARS, sc):
            plan[k] = int(v)
    finally:
        lib.tsbdbg_free(d)
    plan["segs"] = [dict(zip(_SEG, row)) for row in plan["segs"].reshape(-1, 8).tolist()]
    return plan


def emulate_kernel(plan, x, c1, c2, order, gradH=1.0, dtype=np.float64):
    """numpy re-enactment of energy_grad_kernel (tsb_kernels.cu) on the host plan: every CTA, every
    warp walks its cell stream exactly as the kernel does (row blocks, then tet cells, segment by
    segment), with the same formulas.  Returns (energy_total, smooth, barrier, grad[n,3])."""
    G, NW, glob = plan["grid"], plan["nw"], bool(plan["mode_global"])
    IB = 4 if glob else 2
    idt = np.uint32 if glob else np.uint16
    CELL = 1024 if glob else 768
    TPL = 1 if glob else 2
    st = plan["stream"]
    x = np.asarray(x, dtype=np.float32).reshape(-1, 3)
    n = plan["n"]
    grad = np.full((n, 3), np.nan, dtype=dtype)
    grad[plan["orphans"]] = 0.0
    bar_add = np.zeros((n, 3), dtype=dtype)
    es = eb = 0.0
    X4 = plan["X4"].reshape(-1, 4)
    wdesc = plan["wdesc"].reshape(G, NW, 2)
    wseg = plan["wseg"].reshape(-1, NW, 2)
    cta_seg = plan["cta_seg"].reshape(G, 2)
    rows_done = np.zeros(plan["n_components"], dtype=np.int64)
    rows_seen = 0
    lanes = np.arange(32)
    for b in range(G):
        pos = [int(wdesc[b, w, 0]) * 16 for w in range(NW)]
        end = [pos[w] + int(wdesc[b, w, 1]) for w in range(NW)]
        assert all(int(wdesc[b, w, 1]) % CELL == 0 for w in range(NW))
        for s in range(cta_seg[b, 0], cta_seg[b, 1]):
            h = plan["segs"][s]
            nv = h["nv"]
            if glob:
                gids = np.arange(n)
                U = (x - X4[:, :3]).astype(dtype)                       # float32 subtraction, like the kernel
                Pp = x.astype(dtype)
                to_local = lambda a, base=None: a.astype(np.int64)
            else:
                vg = (h["vbase"] + np.arange(nv)) if h["vbase"] >= 0 else plan["vlist"][h["x4off"]:h["x4off"] + nv]
                npos = h["npos"]
                spos = plan["pos16"][h["x4off"]:h["x4off"] + nv].astype(np.int64)
                assert len(set(spos.tolist())) == nv and spos.max() < npos, "staging positions must be distinct"
                gids = plan["pos_gid"][h["p4off"]:h["p4off"] + npos].astype(np.int64)      # position -> global id
                assert np.array_equal(gids[spos], vg)
                U = np.full((npos, 3), np.nan, dtype=dtype)                                 # unused positions hold garbage
                Pp = np.full((npos, 3), np.nan, dtype=dtype)
                U[spos] = (x[vg] - X4[h["x4off"]:h["x4off"] + nv, :3]).astype(dtype)
                Pp[spos] = x[vg].astype(dtype)
                assert npos <= 2047 and npos <= plan["area_verts"] and (h["whole"] or npos <= plan["vh"])
                nv = npos

                li = s - cta_seg[b, 0]
                ub = 0 if h["whole"] else (li & 1) * 2 * plan["vh"]
                xb = h["npos"] if h["whole"] else ub + plan["vh"]

                def to_local(a, base=None):
                    a = a.astype(np.int64)
                    assert np.all(a % 16 == 0)
                    r = a // 16 - (ub if base is None else base)
                    assert r.min() >= 0 and r.max() < nv
                    return r
            for w in range(NW):
                nrb, ntc = (int(v) for v in wseg[s, w])
                p = pos[w]
                for _ in range(nrb):
                    acc = np.zeros((32, 3), dtype=dtype)
                    e = np.zeros(32, dtype=dtype)
                    q, len4 = 0, 1
                    while q < len4:
                        idx = to_local(st[p:p + 128 * IB].view(idt).reshape(32, 4))
                        wbits = st[p + 128 * IB:p + 128 * IB + 512].view(np.uint32).reshape(32, 4)
                        wq = wbits.view(np.float32).astype(dtype)
                        p += CELL
                        if q == 0:
                            hdr = wbits[:, 0]
                            len4 = int((hdr[0] >> 24) & 63)
                            llog = int(hdr[0] >> 30)
                            rid = (hdr & 0xFFFFFF).astype(np.int64)
                            active = rid != 0xFFFFFF
                            assert np.all(((hdr >> 24) & 63) == len4) and np.all((hdr >> 30) == llog) and 1 <= len4 <= 62
                            assert np.all(np.isfinite(wq[:, 0])), "header must read as a finite weight"
                            r = idx[:, 0]
                            ui = U[r]
                        d = U[idx] - ui[:, None, :]
                        assert np.all(d[:, 0 if q == 0 else slice(0, 0)] == 0)
                        assert np.all(wq[~active][:, 1:] == 0)
                        acc += np.einsum("lk,lkr->lr", wq, d)
                        e += np.einsum("lk,lk->l", wq, (d * d).sum(axis=2))
                        q += 1
                    L = 1 << llog
                    assert np.all(r.reshape(-1, L) == r.reshape(-1, L)[:, :1]), "the lanes of a row must be adjacent"
                    tot = acc.reshape(-1, L, 3).sum(axis=1)            # shuffle reduction over the L lanes
                    lead = (lanes % L == 0) & active
                    ra = r[lead]
                    assert np.array_equal(gids[ra], rid[lead]), "header row id must be the global id of the lane's row"
                    assert np.isnan(grad[gids[ra]]).all(), "a vertex row has two writers"
                    grad[gids[ra]] = gradH * c1 * tot[lead[::L]]
                    uref = U[h["x4off"]] if glob else U[0]
                    es += 0.5 * np.einsum("lr,lr->", ui[lead] - uref, tot[lead[::L]])
                    rows_seen += int(lead.sum())
                assert ntc == 0 or w < NW - 1 or NW == 1, "the signalling warp must own no tets"
                for _ in range(ntc):
                    idx = st[p:p + 128 * IB * TPL].view(idt).reshape(32 * TPL, 4)
                    idx = to_local(idx) if glob else to_local(idx, xb)
                    idet = st[p + 128 * IB * TPL:p + 128 * IB * TPL + 128 * TPL].view(np.float32).astype(dtype)
                    p += CELL
                    q = Pp[idx]
                    e1, e2, e3 = q[:, 1] - q[:, 0], q[:, 2] - q[:, 0], q[:, 3] - q[:, 0]
                    c23 = np.cross(e2, e3)
                    J = np.einsum("lr,lr->l", e1, c23) * idet
                    inv = J < 0
                    m = np.where(inv, -J, 0.0)
                    eb += (m ** order).sum()
                    coef = order * m ** (order - 1)
                    k = (-coef * idet * c2 * gradH)[:, None]
                    g1, g2, g3 = k * c23, k * np.cross(e3, e1), k * np.cross(e1, e2)
                    for g, col in ((-(g1 + g2 + g3), 0), (g1, 1), (g2, 2), (g3, 3)):
                        np.add.at(bar_add, gids[idx[inv, col]], g[inv])
                pos[w] = p
            rows_done[h["comp"]] += 1
        assert pos == end, "a warp did not consume exactly its stream"
    for sg in plan["segs"]:
        assert rows_done[sg["comp"]] == sg["expected"], "rows-done counter would never reach `expected`"
    assert rows_seen == n - len(plan["orphans"]) and not np.isnan(grad).any()
    return c1 * es + c2 * eb, es, eb, grad + bar_add
