#!/usr/bin/env python
"""bench.py -- geometry-energy + gradient iterations/sec at 64 tet-spheres x 4096 tets per GPU.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A *step* is one fused energy+gradient pass (tile kernel + combine kernel, 2 launches) over one
synthetic pack of 64 tet-spheres x 4096 tets (BASELINE.json metric; the kernel-only form of
configs[2], whose rasterizer/trainer dependencies are absent -- SURVEY.md F7).  Weak scaling: every
rank owns its own 64-sphere pack (spheres share nothing, so there is no data-path collective); the
scalar energies are all-reduced asynchronously once per graph replay, off the critical path.

Timing rules honoured: W >= 3 warm-up steps; the timed steps rotate over R distinct packs whose
combined footprint exceeds the 126 MB L2, so every step streams its tile data from HBM; CUDA events
on the launching stream, barrier + synchronize on both sides, max over ranks; SM clocks and
throttle reasons sampled through NVML during the timed region.

The reference arm and the cpu_baseline run a CPU *restatement* of the reference's math (the
reference extension needs libpgo + cuSPARSE + a GPU and ships no CPU path: SURVEY.md F2/F4): the
fp64 matrix-free C oracle (oracle/tet_energy_oracle.c, OpenMP over all host threads) -- the fastest
CPU form we have, so the GPU/CPU ratio is not flattered.  The slower "vanilla PyTorch" restatement
of the reference's SpMV pipeline (oracle/torch_energy.py) is timed too and reported in `extras`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SPHERES, TETS = 64, 4096
METRIC = "geometry_energy_grad_iters_per_sec_64x4k"
UNIT = "iters/s"
N_ROTATE = 8           # distinct packs per rank; 8 x ~32 MB of tile data > 126 MB L2
ORDER = 2


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_step"])
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                break
            time.sleep(0.001)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": (float(np.median(self.samples)) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _cpu_restatement_time(pack, x, c1, c2, n_spheres, iters, warmup, threads):
    """Seconds per fwd+bwd of the torch restatement on the first n_spheres spheres of pack."""
    from oracle.torch_energy import time_fwd_bwd
    sub = pack.slice_spheres(0, n_spheres)
    v1 = int(pack.vert_offsets[n_spheres])
    t, e, _ = time_fwd_bwd(sub.verts, sub.tets, x[:v1], c1, c2, ORDER, iters=iters, warmup=warmup, threads=threads)
    return t


def _c_oracle(pack):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _helpers import COracle
    return COracle(pack.verts, pack.tets)


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _best_threads(co, x, c1, c2):
    """Pick the OpenMP thread count that makes the CPU port fastest on this host (all the threads it
    can USE: on a 128-thread box the 64-sphere pack stops scaling well before 128)."""
    avail = _host_threads()
    best = None
    for th in sorted({t for t in (8, 16, 32, 64, avail) if t <= avail} | {avail}):
        co.energy_grad(x, c1, c2, ORDER, nthreads=th)
        t0 = time.perf_counter()
        for _ in range(3):
            co.energy_grad(x, c1, c2, ORDER, nthreads=th)
        dt = (time.perf_counter() - t0) / 3
        if best is None or dt < best[0]:
            best = (dt, th)
    return best[1]


def run_reference(args):
    """CPU arm: the oracle port on the host cores (rank 0 only), all OpenMP threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tssplat_b200.mesh import make_pack, perturb
    pack = make_pack(SPHERES, TETS, seed=0, unique=8)
    x = perturb(pack, sigma_rel=0.02, seed=0)
    c1, c2 = 2e-4 / SPHERES, 2e-4
    co = _c_oracle(pack)
    cores = _best_threads(co, x, c1, c2)
    t0 = time.perf_counter()
    co.energy_grad(x, c1, c2, ORDER, nthreads=cores)
    t_full = time.perf_counter() - t0
    budget = 120.0
    ns = int(max(1, min(SPHERES, SPHERES * budget / max(t_full * (args.steps + args.warmup), 1e-9))))
    if ns < SPHERES:
        sub = pack.slice_spheres(0, ns)
        x = x[: int(pack.vert_offsets[ns])]
        co = _c_oracle(sub)
    for _ in range(args.warmup):
        co.energy_grad(x, c1, c2, ORDER, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.energy_grad(x, c1, c2, ORDER, nthreads=cores)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = 1.0 / (dt * SPHERES / ns)          # block-diagonal by sphere: work is linear in spheres
    sample = (f"{ns} of {SPHERES} spheres per step ({ns * TETS} tets), energy+gradient, extrapolated linearly; "
              f"fp64 matrix-free C oracle, OpenMP, {cores} threads (fastest of the counts tried, {_host_threads()} available)")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * SPHERES / ns,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{SPHERES} tet-spheres x {TETS} tets, energy+gradient, CPU port of the reference's "
                                  "math (the reference itself needs libpgo + a GPU)", "order": ORDER},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tile-tets", type=int, default=0)
    ap.add_argument("--spheres-per-gpu", type=int, default=SPHERES,
                    help="weak scaling (default): spheres owned by every rank")
    ap.add_argument("--total-spheres", type=int, default=0,
                    help="strong scaling: total spheres split sphere-per-rank (BASELINE configs[3], [4])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from tssplat_b200 import _capi
    from tssplat_b200 import tet_spheres_ext as ext
    from tssplat_b200.energies import SmoothnessBarrierEnergy
    from tssplat_b200.mesh import make_pack, perturb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tssplat_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: R distinct packs per rank, resident in HBM ---------------------------------------
    strong = args.total_spheres > 0
    if strong:
        from tssplat_b200.sharding import partition_spheres
        lo, hi = partition_spheres([TETS] * args.total_spheres, world)[rank]
        n_sph = hi - lo
    else:
        n_sph = args.spheres_per_gpu
    total_spheres = args.total_spheres if strong else n_sph * world
    c1, c2 = 2e-4 / total_spheres, 2e-4
    packs, handles, xs = [], [], []
    L2_BYTES = 126e6
    n_rotate = None
    i = 0
    while n_rotate is None or i < n_rotate:
        pk = make_pack(n_sph, TETS, seed=1000 * rank + 17 * i, unique=8)
        packs.append(pk)
        handles.append(ext.TetSpheres(pk.verts.reshape(-1), pk.tets.reshape(-1), tile_tets=args.tile_tets))
        xs.append(torch.from_numpy(perturb(pk, sigma_rel=0.02, seed=i)).to(dev))
        if n_rotate is None:   # enough distinct packs that one rotation streams > 1.5 x L2 through the GPU
            n_rotate = int(min(64, max(2, -(-1.5 * L2_BYTES // handles[0].info["stream_bytes"]))))
            n_rotate = max(n_rotate, N_ROTATE if n_sph <= 128 else 2)
        i += 1
    n = handles[0].n
    info = handles[0].info
    b_alg = float(np.mean([pk.algorithmic_bytes() for pk in packs]))
    footprint = sum(h.info["stream_bytes"] for h in handles)
    energies = torch.zeros((n_rotate, 3), device=dev)
    grads = [torch.empty((h.n, 3), device=dev) for h in handles]
    stream = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev)
    lib = _capi.lib

    def launch(i, st):
        rc = lib.tsb_energy_grad(handles[i]._h, xs[i].data_ptr(), c1, c2, ORDER, 1.0, None,
                                 energies[i].data_ptr(), grads[i].data_ptr(), st)
        if rc != 0:
            raise RuntimeError(_capi.last_error(handles[i]._h))

    with torch.cuda.stream(stream):
        for w in range(args.warmup):
            launch(w % n_rotate, stream.cuda_stream)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for i in range(n_rotate):
                launch(i, stream.cuda_stream)
        graph.replay()
        stream.synchronize()

    def timed_region(steps, rotate=True):
        """Exactly `steps` steps; returns seconds (device time, this rank)."""
        reps, rem = divmod(steps, n_rotate)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(reps):
                graph.replay()
                if world > 1:       # scalar loss all-reduce, once per replay, on the side stream
                    comm.wait_stream(stream)
                    with torch.cuda.stream(comm):
                        dist.all_reduce(energies, op=dist.ReduceOp.SUM, async_op=True)
            for i in range(rem):
                launch(i, stream.cuda_stream)
            stream.wait_stream(comm)
            ev1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return ev0.elapsed_time(ev1) * 1e-3

    timed_region(min(args.steps, 10 * n_rotate))                      # settle clocks / NCCL
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_local = timed_region(args.steps)
    clocks = sampler.stop()
    t_all = torch.tensor([t_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    t = float(t_all.item())
    ms_per_step = t / args.steps * 1e3
    # weak: every rank steps its own pack -> aggregate pack-iterations/s; strong: all ranks together
    # advance ONE pack of total_spheres per step.  Reported in units of the 64-sphere metric pack.
    value = args.steps / t * (total_spheres / SPHERES)

    # ---- warm-L2 variant (one pack replayed): explains the launch-latency floor ---------------
    with torch.cuda.stream(stream):
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, stream=stream):
            for _ in range(n_rotate):
                launch(0, stream.cuda_stream)
        g1.replay(); stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(max(1, args.steps // n_rotate)):
            g1.replay()
        e1.record(stream)
        stream.synchronize()
    warm_ms = e0.elapsed_time(e1) / (max(1, args.steps // n_rotate) * n_rotate)

    # ---- end to end with HOST buffers: (1) through the C-ABI host entry point, (2) through the
    # reference-facing autograd surface (SmoothnessBarrierEnergy) -- copies inside the timed region
    x_host = xs[0].cpu().pin_memory()
    g_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    e_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def timed_e2e(step_fn, steps):
        for _ in range(5):
            step_fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for _ in range(steps):
            step_fn()
        ee1.record()
        torch.cuda.synchronize()
        te = torch.tensor([ee0.elapsed_time(ee1) * 1e-3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * steps / float(te.item())

    e2e_scale = (n_sph / SPHERES) if not strong else (total_spheres / SPHERES) / world
    e2e_steps = int(min(args.steps, 500))
    e2e_value = e2e_scale * timed_e2e(lambda: ext.energy_grad_host(handles[0], x_host, c1, c2, ORDER, 1.0, e_host, g_host), e2e_steps)

    eng = SmoothnessBarrierEnergy.__new__(SmoothnessBarrierEnergy)
    torch.nn.Module.__init__(eng)
    from types import SimpleNamespace
    eng.tet_sp, eng.FLAGS = handles[0], SimpleNamespace(smooth_eng_coeff=c1, barrier_coeff=c2, increase_order_iter=10 ** 9)
    from tssplat_b200.energies import SmoothnessBarrierFunc
    eng.smooth_eng_func = SmoothnessBarrierFunc
    tet_v = torch.nn.Parameter(torch.empty((n, 3), device=dev))
    e0_host = torch.empty((), dtype=torch.float32).pin_memory()

    def autograd_step():
        tet_v.grad = None
        with torch.no_grad():
            tet_v.copy_(x_host, non_blocking=True)                   # H2D of this step's input
        e = eng(tet_v, 0, c1, c2)                                     # forward (fused launch)
        e.backward()                                                  # backward (rescale of cached grad)
        g_host.copy_(tet_v.grad, non_blocking=True)                  # D2H of the result
        e0_host.copy_(e.detach(), non_blocking=True)

    e2e_autograd = e2e_scale * timed_e2e(autograd_step, int(min(args.steps, 300)))

    if rank == 0:
        peak, peak_src = _peaks()
        achieved = b_alg / (t / args.steps) / 1e9                   # GB/s per GPU, whole step (both launches)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n_sph} tet-spheres x {TETS} tets per GPU ({total_spheres} in total), fused energy+grad "
                                   "(tile kernel + combine kernel); kernel-only form of BASELINE configs[2] (rasterizer "
                                   "deps absent); value is in 64-sphere-pack iterations/s",
                       "spheres_per_gpu": n_sph, "total_spheres": total_spheres, "tets_per_sphere": TETS, "vertices": int(n), "order": ORDER,
                       "x": "rest + N(0,(0.02 h)^2), no inverted tets", "parallelism": f"sphere-per-rank x{world}",
                       "l2": f"inputs larger than L2: rotating {n_rotate} distinct packs, {footprint / 1e6:.0f} MB "
                             "of per-step data > 126 MB L2", "tile_tets": int(info["tile_tets"]),
                       "tiles": int(info["n_tiles"]), "graph": f"CUDA graph of {n_rotate} steps replayed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_step": b_alg,
                         "note": "achieved = B_alg (24V+68T per sphere) / CUDA-event step time incl. both launches"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 12),
                    "d2h_bytes_per_step": int(n * 12 + 12), "steps": e2e_steps,
                    "path": "C-ABI tsb_energy_grad_host: pinned host x -> H2D -> fused launch -> D2H grad + energy[3]"},
            "gpu_launches": int(2 * args.steps),
            "clocks": clocks,
            "extras": {"e2e_autograd_surface_iters_per_s": e2e_autograd,
                       "e2e_autograd_surface_note": "pinned host x -> H2D -> SmoothnessBarrierEnergy.forward/backward "
                                                    "(torch.autograd.Function) -> D2H grad+energy; Python/autograd-bound",
                       "warm_l2_ms_per_step": warm_ms, "warm_l2_iters_per_s": 1e3 / warm_ms,
                       "stream_bytes_per_step": int(info["stream_bytes"])},
        }
        if world == 1 and not args.no_cpu_baseline and n_sph == SPHERES:
            x0 = xs[0].cpu().numpy()
            co = _c_oracle(packs[0])
            cores = _best_threads(co, x0, c1, c2)
            reps = 0
            t0 = time.perf_counter()
            while reps < 200 and time.perf_counter() - t0 < 15.0:
                co.energy_grad(x0, c1, c2, ORDER, nthreads=cores)
                reps += 1
            tc = (time.perf_counter() - t0) / reps
            out["cpu_baseline"] = {"value": 1.0 / tc, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{reps} energy+gradient iterations of the full 64-sphere pack; fp64 "
                                             f"matrix-free C oracle (oracle/tet_energy_oracle.c), OpenMP, {cores} threads "
                                             f"(fastest of the counts tried, {_host_threads()} available)"}
            # the reference-shaped "vanilla PyTorch" pipeline (SpMV GTLTLG, SpMV G, autograd), best thread count
            best = None
            ns = 2
            avail = _host_threads()
            for th in sorted({min(avail, 8), min(avail, 32), avail}):
                tt_ = _cpu_restatement_time(packs[0], x0, c1, c2, ns, iters=4, warmup=1, threads=th)
                if best is None or tt_ < best[0]:
                    best = (tt_, th)
            out["extras"]["cpu_torch_restatement_iters_per_s"] = 1.0 / (best[0] * SPHERES / ns)
            out["extras"]["cpu_torch_restatement_note"] = (f"torch sparse fp32 + autograd restatement of the reference's SpMV "
                                                           f"pipeline, {ns} of {SPHERES} spheres extrapolated, best of thread "
                                                           f"counts -> {best[1]} threads")
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
