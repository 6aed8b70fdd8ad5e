#!/usr/bin/env python
"""bench.py — headline benchmark of the particle-deposition hot path (BASELINE.json metric: Mparticles/s mapped).

  python bench.py --gpus N --steps K --warmup W            (under torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    (the reference algorithm's CPU path, oracle port)

A step = one `sphMapping` / `healpix_map` / stencil pass over the whole synthetic particle set of the workload:
centre + filter + deposit (+ the sum of the partial images over the ranks for N > 1) + reduce_image.
Default workload "c2" = BASELINE.json configs[1]: 16 777 216 Gadget-like particles, 4096^2 map, WendlandC6(2),
calc_mean mass-weighted temperature map (q = T, w = rho, reduce_image = true).

value    : whole-job Mparticles/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks.
e2e      : the same step through the reference-facing C ABI call (s2g_sphmap / s2g_healpix_map / s2g_stencil_deposit)
           with PAGEABLE host arrays (what a Julia `Array` is): H2D of the inputs and D2H of the map inside the region.
roofline : the BINDING roof of the workload's dominant kernel — "fp64" (FP64 issue, tile-gather kernels) or "atomic"
           (L2 red.f64 throughput, scatter kernels); the HBM figure of the contract is the secondary key `hbm`.
parity   : GPU vs the CPU oracle on the first particles of the SAME stream at full image size (max per-pixel relative
           error, counters compared exactly); HEALPix against the extended-precision arbiter.
extra    : (default run only) the other BASELINE configs — C3 (+ CIC/TSC on the same set), C4, C5 — each with its
           own steps/warm-up, roofline and parity, so that one driver run carries every config.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: particles, npix, dims, kernel, n_ngb, seed   (SURVEY.md §8d); dims 0 = HEALPix; stencil = CIC(2)/TSC(3) order
    "c2": dict(n=16 * 1024 * 1024, npix=4096, dims=2, kernel="WendlandC6", n_ngb=295.0, seed=2,
               desc="synthetic 16M-particle box, 2D calc_mean mass-weighted T map, 4096^2, WendlandC6"),
    "c3": dict(n=64 * 1024 * 1024, npix=512, dims=3, kernel="Cubic", n_ngb=64.0, seed=3,
               desc="synthetic 64M particles, 3D sphMapping onto 512^3, Cubic"),
    "c3cic": dict(n=64 * 1024 * 1024, npix=512, dims=3, kernel="Cubic", n_ngb=64.0, seed=3, stencil=2,
                  desc="CIC deposit of the synthetic 64M-particle set onto 512^3"),
    "c3tsc": dict(n=64 * 1024 * 1024, npix=512, dims=3, kernel="Cubic", n_ngb=64.0, seed=3, stencil=3,
                  desc="TSC deposit of the synthetic 64M-particle set onto 512^3"),
    "c4": dict(n=128 * 1024 * 1024, npix=2048, dims=0, kernel="WendlandC4", n_ngb=200.0, seed=4,
               desc="synthetic 128M particles, healpix_map all-sky Nside=2048, WendlandC4, shell [0.05L,0.5L]"),
    "c4s": dict(n=4 * 1024 * 1024, npix=2048, dims=0, kernel="WendlandC4", n_ngb=200.0, seed=4,
                desc="4M particles of the c4 stream (hsml of the 128M set), Nside=2048 (debug)", n_stream=128 * 1024 * 1024),
    "c4t": dict(n=256 * 1024, npix=2048, dims=0, kernel="WendlandC4", n_ngb=200.0, seed=4,
                desc="256k particles of the c4 stream, Nside=2048 (profiling)", n_stream=128 * 1024 * 1024),
    "c3s": dict(n=8 * 1024 * 1024, npix=512, dims=3, kernel="Cubic", n_ngb=64.0, seed=3,
                desc="8M particles of the c3 stream, 512^3 (debug)", n_stream=64 * 1024 * 1024),
    "c5": dict(n=1024 * 1024 * 1024, npix=8192, dims=2, kernel="WendlandC6", n_ngb=295.0, seed=5,
               desc="synthetic 1B-particle box, 8192^2 2D map, WendlandC6"),
    "c5s": dict(n=32 * 1024 * 1024, npix=8192, dims=2, kernel="WendlandC6", n_ngb=295.0, seed=5,
                desc="32M particles of the c5 stream (hsml of the 1B set), 8192^2 (debug)", n_stream=1024 * 1024 * 1024),
    "tiny": dict(n=64 * 1024 * 1024, npix=4096, dims=2, kernel="WendlandC6", n_ngb=0.05, seed=6,
                 desc="64M particles with ~2-pixel kernels, 4096^2 (scatter regime)"),
    "c3big": dict(n=512 * 1024, npix=512, dims=3, kernel="Cubic", n_ngb=4096.0, seed=3,
                  desc="512k particles with ~13-cell kernels, 512^3 (3D large-footprint regime)", n_stream=64 * 1024 * 1024),
    "tinys": dict(n=4 * 1024 * 1024, npix=4096, dims=2, kernel="WendlandC6", n_ngb=0.05, seed=6,
                  desc="4M particles of the tiny stream (profiling)", n_stream=64 * 1024 * 1024),
    "small": dict(n=1 << 20, npix=1024, dims=2, kernel="WendlandC6", n_ngb=295.0, seed=2,
                  desc="1M particles, 1024^2 (debug)"),
}
SIGMA = 1.5
KERNEL_DIM = {2: 2, 3: 3, 0: 2}
HP_SHELL = (0.05, 0.5)
# the other BASELINE configs carried by the default run: (workload, steps, warm-up)
EXTRAS = [("c3", 3, 2), ("c3cic", 3, 2), ("c3tsc", 3, 2), ("tiny", 3, 2), ("c4", 1, 1), ("c5", 1, 1)]
# particles of the stream the parity check / CPU baseline run on: (main workload, extra entry)
PARITY_SAMPLE = {2: (1 << 17, 1 << 15), 3: (1 << 20, 1 << 18), 0: (1 << 16, 1 << 13), "stencil": (1 << 22, 1 << 20)}
CPU_SAMPLE = {2: 1 << 19, 3: 1 << 20, 0: 1 << 16, "stencil": 1 << 24}


def wl_key(wl):
    return "stencil" if wl.get("stencil") else wl["dims"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons every 200 ms while the timed region runs (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.sm = []
        self.reasons = set()
        self.max_mhz = None
        self.ok = False

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
            self.ok = True
            while not self.stop_flag:
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.2)
        except Exception:
            self.ok = False

    def summary(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------------------------------
# host-side particles of a stream (parity samples, CPU baseline, reference arm)
# ----------------------------------------------------------------------------------------------------------------------
_HOST_CACHE = {}


USE_DEVICE_STREAM = True   # the reference arm switches this off: nothing of the product is loaded there


def host_particles(wl, count):
    """First `count` particles of the workload's stream on the host.  The stream is DEFINED by the device generator
    (counter-based, keyed by the global particle id), so with a GPU the particles are generated there and copied back;
    without one (the CPU-only reference arm on a box without a device) a numpy stand-in with the same recipe."""
    key = (wl["seed"], wl.get("n_stream", wl["n"]), count)
    if key in _HOST_CACHE:
        return _HOST_CACHE[key]
    n_stream = wl.get("n_stream", wl["n"])
    s2g = None
    if USE_DEVICE_STREAM:
        import __graft_entry__ as ge
        s2g = ge.load_package()
    if s2g is not None and s2g.lib().s2g_device_count() > 0:
        import torch
        from sphtogrid_b200 import _lib
        ctx = s2g.default_context()
        dev = torch.device("cuda", ctx.device)
        t = [torch.empty(count * 3 if i == 0 else count, dtype=torch.float64, device=dev) for i in range(5)]
        _lib.check(s2g.lib().s2g_synth_particles_dev(ctx.handle, wl["seed"], 0, count, n_stream, 1.0, wl["n_ngb"],
                                                     SIGMA, 1, *[_lib.ptr(x.data_ptr()) for x in t]))
        ctx.sync()
        out = [x.cpu().numpy() for x in t]
        out[0] = out[0].reshape(count, 3)
        del t
    else:
        rng = np.random.default_rng(wl["seed"])
        pos = rng.random((count, 3))
        g = rng.normal(size=count)
        rho = np.exp(SIGMA * g - 0.5 * SIGMA ** 2)
        mass = np.full(count, 1.0 / n_stream)
        hsml = np.cbrt(3.0 * wl["n_ngb"] * mass / (4.0 * np.pi * rho))
        temp = 1e4 * rho ** (2.0 / 3.0) * np.exp(0.5 * rng.normal(size=count))
        out = [pos, hsml, mass, rho, temp]
    _HOST_CACHE[key] = tuple(out)
    return _HOST_CACHE[key]


def shell_sample(wl, count):
    """The first `count` particles of the stream inside the HEALPix shell around the box centre, recentred."""
    pos, hsml, m, rho, temp = host_particles(wl, 2 * count)
    pos = pos - 0.5
    r = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2 + pos[:, 2] ** 2)
    sel = np.flatnonzero((r >= HP_SHELL[0]) & (r <= HP_SHELL[1]))[:count]
    return np.ascontiguousarray(pos[sel]), hsml[sel], m[sel], rho[sel], temp[sel]


def oracle_map(orc, wl, sample, cores, exact=False):
    """The reference's CPU path (oracle port) on the first `sample` particles; returns (map(s), counters, seconds)."""
    dims, npix = wl["dims"], wl["npix"]
    if wl.get("stencil"):
        pos, hsml, m, rho, temp = host_particles(wl, sample)
        p = pos - 0.5
        t0 = time.perf_counter()
        img = orc.stencil_deposit(wl["stencil"], 3, p, rho, float(npix), npix, False)
        return img, None, time.perf_counter() - t0
    if dims == 0:
        pos, hsml, m, rho, temp = shell_sample(wl, sample)
        t0 = time.perf_counter()
        a, w, st = orc.healpix_deposit(pos, hsml, m, rho, temp, rho, npix, wl["kernel"], 2, True, n_workers=cores,
                                       exact="sens" if exact else False)
        return (a, w), st, time.perf_counter() - t0
    pos, hsml, m, rho, temp = host_particles(wl, sample)
    par = orc.mapping_parameters(center=[0.5, 0.5, 0.5], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=npix)
    t0 = time.perf_counter()
    p, par_c = orc.center_particles(pos.copy(), par)
    if dims == 2:
        flat, st = orc.cic_mapping_2d(p, hsml, m, rho, temp, rho, par_c.len2pix, npix, wl["kernel"], 2, True,
                                      n_workers=cores)
        out = orc.reduce_image_2d(flat, npix, npix, True)
    else:
        flat, st = orc.cic_mapping_3d(p, hsml, m, rho, rho, np.ones_like(rho), par_c.len2pix, npix, wl["kernel"], 3,
                                      False, n_workers=cores)
        out = orc.reduce_image_3d(flat, npix, True)
    return out, st, time.perf_counter() - t0


def cpu_workers(wl):
    """Host threads of the CPU legs: every core, bounded so that the per-worker private images (one full image + two
    scratch planes per worker, like the reference's Distributed workers) fit into 40 % of the host memory."""
    cores = os.cpu_count() or 1
    ncell = wl["npix"] ** (wl["dims"] or 2) * (12 if wl["dims"] == 0 else 1)
    per_worker = ncell * 8 * 4
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    return int(max(1, min(cores, (0.4 * avail) // per_worker)))


def reference_arm(args, wl):
    """The reference's own CPU implementation of the path (oracle port of cic_mapping_2D/3D with `parallel=true`
    slicing over all host threads; HEALPix: the particle loop over private maps), built -O3 with FMA contraction
    (oracle/libs2g_oracle_fast.so, rebuilt -march=native on the box when gcc is there), on a bounded sample."""
    from oracle import oracle as orc
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc.select_library("fast")
    global USE_DEVICE_STREAM
    USE_DEVICE_STREAM = False      # same recipe drawn with numpy: the CPU arm loads nothing of the product
    cores = cpu_workers(wl)
    sample = args.cpu_sample
    if not sample:
        # bounded sample: the largest power of two (<= CPU_SAMPLE) for which warm-up + steps end within ~5 minutes,
        # from a two-point calibration (the per-worker image allocation is a fixed cost, the rest is linear)
        top = CPU_SAMPLE[wl_key(wl)]
        c0, c1 = max(top >> 6, 256), max(top >> 5, 512)
        t0 = oracle_map(orc, wl, c0, cores)[2]
        t1 = oracle_map(orc, wl, c1, cores)[2]
        slope = max((t1 - t0) / (c1 - c0), 1e-12)
        fixed = max(t0 - slope * c0, 0.0)
        budget = 300.0 / max(1, args.warmup + args.steps)
        sample = top
        while sample > c1 and fixed + slope * sample > budget:
            sample >>= 1
    times = []
    for it in range(args.warmup + args.steps):
        _, _, dt = oracle_map(orc, wl, sample, cores)
        if it >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = sample / t / 1e6
    line = {"impl": "reference", "metric": "Mparticles/s mapped", "value": val, "unit": "Mparticles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "particles": wl["n"], "npix": wl["npix"], "kernel": wl["kernel"]},
            "cpu_baseline": {"value": val, "unit": "Mparticles/s", "cores": cores, "kind": "port",
                             "build": "gcc -O3, FMA contraction on (libs2g_oracle_fast.so)",
                             "sample": f"first {sample} particles of the same synthetic stream, full-size image"},
            "e2e": {"value": val, "unit": "Mparticles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
class Env:
    pass


def make_env(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    e = Env()
    e.torch, e.dist = torch, dist
    e.s2g = ge.load_package()
    from sphtogrid_b200 import _lib
    e._lib = _lib
    e.L = e.s2g.lib()
    e.world = int(os.environ.get("WORLD_SIZE", "1"))
    e.rank = int(os.environ.get("RANK", "0"))
    e.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(e.local_rank)
    e.dev = torch.device("cuda", e.local_rank)
    if e.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=e.dev)
    e.ctx = e.s2g.Context(e.local_rank, strategy=args.strategy)
    if args.accum != "f64":
        e.ctx.set_accumulate_mode(args.accum)
    e.stream = torch.cuda.current_stream(e.dev)
    e.ctx.set_stream(e.stream.cuda_stream)
    e.args = args
    e.peaks = None
    return e


def live_peaks(e):
    """FP64 DFMA rate and coalesced / random red.f64 rates, measured in this process at these clocks."""
    if e.peaks is None:
        r = C.c_double(0)
        out = {}
        e._lib.check(e.L.s2g_microbench(e.ctx.handle, 0, 0, 20000, C.byref(r)))
        out["fp64_gflops"] = r.value
        e._lib.check(e.L.s2g_microbench(e.ctx.handle, 1, 256 << 20, 2000, C.byref(r)))
        out["red_rows_g"] = r.value
        e._lib.check(e.L.s2g_microbench(e.ctx.handle, 2, 1 << 30, 2000, C.byref(r)))
        out["red_random_g"] = r.value
        e.peaks = out
    return e.peaks


def run_workload(e, name, steps, warmup, main):
    """Times one workload on this process group; returns the result dict on rank 0 (None elsewhere)."""
    torch, dist, s2g, _lib, L, ctx = e.torch, e.dist, e.s2g, e._lib, e.L, e.ctx
    args, world, rank, dev, stream = e.args, e.world, e.rank, e.dev, e.stream
    wl = WORKLOADS[name]
    n_total = wl["n"]
    s, en = s2g.domain_decomposition(n_total, world)[rank]
    n_loc = en - s
    dims, npix, stencil = wl["dims"], wl["npix"], wl.get("stencil", 0)
    healpix = dims == 0
    f64 = torch.float64
    P = lambda t: _lib.ptr(t.data_ptr())
    d_pos = torch.empty(n_loc * 3, dtype=f64, device=dev)
    d_h, d_m, d_rho, d_T = (torch.empty(n_loc, dtype=f64, device=dev) for _ in range(4))
    _lib.check(L.s2g_synth_particles_dev(ctx.handle, wl["seed"], s, n_loc, wl.get("n_stream", n_total), 1.0,
                                         wl["n_ngb"], SIGMA, 1, P(d_pos), P(d_h), P(d_m), P(d_rho), P(d_T)))
    if stencil:
        d_pos -= 0.5            # the stencils take positions relative to the image centre
        del d_h, d_m, d_T
        d_h = d_m = d_T = None
    d_one = torch.ones(n_loc, dtype=f64, device=dev) if (dims == 3 and not stencil) else None
    par = s2g.mappingParameters(center=[0.5, 0.5, 0.5], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=npix)
    par2 = s2g.recentred_parameters(par)
    kid = getattr(s2g, wl["kernel"])(KERNEL_DIM[dims]).kernel_id
    ncell = npix ** dims if not healpix else 12 * npix * npix
    planes = 2
    image = torch.empty(ncell * planes, dtype=f64, device=dev)
    out = torch.empty(ncell, dtype=f64, device=dev)
    q_t, w_t = (d_T, d_rho) if dims != 3 else (d_rho, d_one)
    shift, half = _lib.dbl3(par.center), _lib.dbl3(par2.halfsize)
    hp_center = (C.c_double * 3)(0.5, 0.5, 0.5)          # observer at the box centre
    hp_shell = (C.c_double * 2)(*HP_SHELL)               # radius_limits = [0.05 L, 0.5 L]
    hp_nsel = C.c_int64(0)

    def deposit_only():
        if stencil:
            _lib.check(L.s2g_stencil_deposit_dev(ctx.handle, stencil, 3, P(d_pos), P(d_rho), n_loc, 1, float(npix), npix,
                                                 0, 0, P(image)))
        elif healpix:
            # the whole healpix_map body on the device: Pos .-= center, shell filter, far-to-near selection
            # (filter_sort_particles incl. its sorted[mask] semantics), particle loop
            _lib.check(L.s2g_healpix_map_dev(ctx.handle, P(d_pos), P(d_h), P(d_m), P(d_rho), P(q_t), P(w_t), n_loc,
                                             hp_center, hp_shell, npix, kid, 1, 0, P(image), P(image[ncell:]),
                                             C.byref(hp_nsel)))
        else:
            _lib.check(L.s2g_sphmap_dev(ctx.handle, dims, P(d_pos), P(d_h), P(d_m), P(d_rho), P(q_t), P(w_t), n_loc, 1,
                                        1, shift, 0, -1.0, half, float(par2.len2pix), npix, kid, 1, 0, P(image)))

    from sphtogrid_b200 import distributed as sdist
    xwork = {}
    divide = sdist.device_divide(ctx)

    def exchange_only():
        if healpix or stencil:
            dist.reduce(image, dst=0)
            return
        flat = sdist.exchange_reduce(image, 1, ncell, dims, True, divide, work=xwork, gather="root")
        if rank == 0:
            if dims == 2:
                _lib.check(L.s2g_reduce_image_2d_dev(ctx.handle, P(flat), npix, npix, 1, 0, P(out)))
            else:
                out.copy_(flat)

    def step_device():
        deposit_only()
        if world > 1:
            # image = sum(fetch.(futures)) (cic_interpolation.jl:199): the one exchange step of the path — un-reduced
            # HEALPix / stencil maps are reduced to the master; 2D / 3D: reduce-scatter per plane, reduce_image division
            # on this rank's pixel slice, gather on the master, transposition to Array(N,N,1) memory
            exchange_only()
            return
        if healpix or stencil:
            return
        if dims == 2:
            _lib.check(L.s2g_reduce_image_2d_dev(ctx.handle, P(image), npix, npix, 1, 1, P(out)))
        else:
            _lib.check(L.s2g_reduce_image_3d_dev(ctx.handle, P(image), npix, 1, P(out)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps_, warmup_, sampler=None):
        for _ in range(warmup_):
            fn()
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps_):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        if sampler:
            sampler.stop_flag = True
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], dtype=f64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / steps_, float(t[1]) / steps_

    sampler = ClockSampler(e.local_rank) if (rank == 0 and main) else None
    ms_step, wall_step = timed(step_device, steps, warmup, sampler)
    # the exchange step alone (N > 1): device time of reduce-scatter + slice division + gather + transposition
    exchange_ms = None
    if world > 1:
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(stream)
        for _ in range(3):
            exchange_only()
        x1.record(stream)
        barrier()
        tx = torch.tensor([x0.elapsed_time(x1) / 3.0], dtype=f64, device=dev)
        dist.all_reduce(tx, op=dist.ReduceOp.MAX)
        exchange_ms = float(tx[0])
    # counters and phase times of one more deposit (the last library call of a step is the reduce)
    deposit_only()
    st = ctx.stats()
    n_in = int(hp_nsel.value) if healpix else n_loc
    cnt = torch.tensor([st["n_mapped"], st["footprint_pixels"], st["touched_pixels"], st["n_pairs"],
                        st["n_launches"] + (0 if healpix else 1), n_in], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    n_mapped, fpx_all, touched_all, pairs, launches, n_in_all = [int(x) for x in cnt.tolist()]
    fmax = torch.tensor([float(st["footprint_pixels"])], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(fmax, op=dist.ReduceOp.MAX)
    imb = (float(fmax[0]) / (fpx_all / world) - 1.0) if fpx_all > 0 else 0.0
    if stencil:
        n_mapped = n_total
    fpx, touched = int(st["footprint_pixels"]), int(st["touched_pixels"])
    value = n_mapped / (ms_step * 1e-3) / 1e6

    # ---- end to end through the C ABI with HOST arrays (pageable numpy = a Julia Array; pinned for comparison)
    e2e = None
    if not args.no_e2e and (main or name in ("c3", "c4")) and n_loc * 64 <= (24 << 30):
        host = [t.cpu().numpy() if t is not None else None for t in (d_pos, d_h, d_m, d_rho, q_t, w_t)]
        res = {}
        for kind in (("pageable", "pinned") if main else ("pageable",)):
            if kind == "pinned":
                keep = [torch.empty(a.shape, dtype=f64, pin_memory=True) if a is not None else None for a in host]
                for k_, a in zip(keep, host):
                    if a is not None:
                        k_.numpy()[...] = a
                hb = [k_.numpy() if k_ is not None else None for k_ in keep]
                h_out_t = torch.empty(ncell * (2 if healpix else 1), dtype=f64, pin_memory=True)
                h_out = h_out_t.numpy()
            else:
                hb = host
                h_out = np.empty(ncell * (2 if healpix else 1))
            hp = lambda a: _lib.ptr(a) if a is not None else None
            s_, e_ = (0, n_loc)

            def step_e2e():
                if world > 1:
                    # one process per GPU: every rank stages ITS slice from host memory, rank 0 receives the map
                    for dt_, ha in ((d_pos, hb[0]), (d_h, hb[1]), (d_m, hb[2]), (d_rho, hb[3])):
                        if dt_ is not None:
                            dt_.copy_(torch.from_numpy(ha), non_blocking=True)
                    step_device()
                    if rank == 0:
                        torch.from_numpy(h_out[:ncell]).copy_(out if not healpix else image[:ncell])
                    stream.synchronize()
                elif stencil:
                    _lib.check(L.s2g_stencil_deposit(ctx.handle, stencil, 3, hp(hb[0]), hp(hb[3]), n_loc, 1,
                                                     float(npix), npix, 0, hp(h_out2), None))
                elif healpix:
                    _lib.check(L.s2g_healpix_map(ctx.handle, hp(hb[0]), hp(hb[1]), hp(hb[2]), hp(hb[3]), hp(hb[4]),
                                                 hp(hb[5]), n_loc, hp_center, hp_shell, npix, kid, 1, None,
                                                 hp(h_out[:ncell]), hp(h_out[ncell:]), None))
                else:
                    _lib.check(L.s2g_sphmap(ctx.handle, dims, hp(hb[0]), hp(hb[1]), hp(hb[2]), hp(hb[3]), hp(hb[4]),
                                            hp(hb[5]), n_loc, 1, 1, shift, 0, -1.0, half, float(par2.len2pix), npix,
                                            kid, 1, 1, 0, None, hp(h_out), None))
            if stencil:
                h_out2 = np.empty(ncell * 2)
            ms_e, wall_e = timed(step_e2e, max(1, min(steps, 5)), 1)
            res[kind] = max(ms_e, wall_e)
            if world == 1:
                stx = ctx.stats()
                res[kind + "_phases"] = {k: round(float(stx.get(k, 0.0)), 2) for k in
                                         ("ms_h2d", "ms_compute", "ms_epilogue", "ms_d2h", "ms_total")}
        n_in_arrays = sum(1 for a in host if a is not None)
        e2e = {"value": n_mapped / (res["pageable"] * 1e-3) / 1e6, "unit": "Mparticles/s",
               "h2d_bytes_per_step": int(sum(a.nbytes for a in host if a is not None) * world),
               "d2h_bytes_per_step": int(ncell * 8 * (2 if (healpix or stencil) else 1)),
               "ms_per_step": res["pageable"], "host_memory": "pageable (numpy = Julia Array)",
               "pinned_ms_per_step": res.get("pinned"), "arrays": n_in_arrays,
               "phases_last_call": res.get("pageable_phases")}
        del host

    result = None
    if rank == 0:
        pk, pk_kind = peaks()
        lp = live_peaks(e)
        hbm_peak = float(pk.get("hbm_gbs", 6650.0))
        dep_ms = st["ms_deposit"] if st["ms_deposit"] > 0 else ms_step
        strategy_gather = dims == 2 and st["n_pairs"] > 0 and st["n_gather"] >= st["n_scatter"]
        hp_gather = healpix and st["n_pairs"] > 0
        kernel_name = ("k_stencil" if stencil else
                       {2: "k_gather2d" if strategy_gather else "k_scatter2d", 3: "k_scatter3d",
                        0: "k_hp_gather" if hp_gather else "k_healpix"}[dims])
        # launches of the dominant kernel per step on this rank: the gathers walk the shard in slices
        # (S2G_BATCH_PARTICLES, default 8 Mi particles), one launch each; 3D / stencil: one launch per step
        n_dom = max(1, -(-n_loc // int(os.environ.get("S2G_BATCH_PARTICLES", 8 << 20)))) if strategy_gather else 1
        launch_ms = dep_ms / n_dom
        # algorithmic bytes per launch (SURVEY §8d roof 1): every particle field of the slice read once + every image
        # plane written once
        in_bytes = 32 if stencil else 64
        alg_bytes = (n_loc // n_dom) * in_bytes + ncell * planes * 8
        hbm = {"achieved": alg_bytes / (launch_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "frac": alg_bytes / (launch_ms * 1e-3) / 1e9 / hbm_peak, "peak_kind": pk_kind,
               "alg_bytes_per_launch": alg_bytes}
        flop_per_px = {2: 40.0, 3: 45.0, 0: 80.0}[dims]   # SURVEY §8d roof 3: pass A + pass B per footprint pixel
        fp_ms = dep_ms + st["ms_norm"]
        fp64 = {"achieved": fpx * flop_per_px / (fp_ms * 1e-3) / 1e12, "peak": lp["fp64_gflops"] / 1e3,
                "unit": "TFLOP/s", "flop_per_footprint_pixel": flop_per_px, "footprint_pixels": fpx,
                "peak_kind": "live DFMA microbenchmark (s2g_microbench 0), same process and clocks"}
        fp64["frac"] = fp64["achieved"] / fp64["peak"]
        if stencil:
            n_red = planes * n_loc * stencil ** 3
            red_peak, red_kind = lp["red_rows_g"], ("live red.f64 microbenchmark, 32 consecutive doubles per warp (a "
                                                    "stencil's reds fall on 2-3 consecutive cells of %d rows: the "
                                                    "random-address rate, %.1f Gred/s over 1 GiB, is the other bracket)"
                                                    % (stencil ** 2, lp["red_random_g"]))
        else:
            n_red = planes * touched
            red_peak, red_kind = lp["red_rows_g"], "live red.f64 microbenchmark, 32 consecutive doubles per warp"
        atomic = {"achieved": n_red / (dep_ms * 1e-3) / 1e9, "peak": red_peak, "unit": "Gred/s", "reds": n_red,
                  "peak_kind": red_kind}
        atomic["frac"] = atomic["achieved"] / atomic["peak"]
        gather = strategy_gather or hp_gather
        bind = fp64 if gather else atomic
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if world == 1 and os.path.exists(tpath):
            tj = json.load(open(tpath)).get(name, {}).get(kernel_name)
            if tj:
                traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], tj.get("source")
        roofline = {"bound": "fp64" if gather else "atomic", "achieved": bind["achieved"], "peak": bind["peak"],
                    "unit": bind["unit"], "frac": bind["frac"], "traffic": traffic, "traffic_source": traffic_src,
                    "kernel": kernel_name, "launches_per_step": n_dom, "launch_ms": launch_ms,
                    "phase_ms": {"norm": st["ms_norm"], "deposit": dep_ms, "sort": st["ms_sort"], "prep": st["ms_prep"]},
                    "why": ("tile-gather: no global atomics in the inner loop, FP64 issue binds (DESIGN.md §4)" if gather
                            else "scatter: two red.global.add.f64 per touched pixel bind (DESIGN.md §4)"),
                    "hbm": hbm, "fp64": fp64, "atomic": atomic}

        # ---- parity on the first particles of the same stream, and the CPU baseline beside it
        parity, cpu_baseline = None, None
        if world == 1 and not args.no_parity:
            parity = parity_check(e, name, PARITY_SAMPLE[wl_key(wl)][0 if main else 1])
        if world == 1 and not args.no_cpu_baseline and (main or name in ("c3", "c4")):
            from oracle import oracle as orc
            orc.select_library("fast")
            cores = cpu_workers(wl)
            sample = args.cpu_sample or (CPU_SAMPLE[wl_key(wl)] if main else CPU_SAMPLE[wl_key(wl)] // 4)
            _, _, dt = oracle_map(orc, wl, sample, cores)
            orc.select_library("checker")
            cpu_baseline = {"value": sample / dt / 1e6, "unit": "Mparticles/s", "cores": cores, "kind": "port",
                            "build": "gcc -O3, FMA contraction on (libs2g_oracle_fast.so)",
                            "sample": f"first {sample} particles of the same stream, full-size image, {dt:.1f} s"}
        result = {"metric": "Mparticles/s mapped", "value": value, "unit": "Mparticles/s", "n_gpus": world,
                  "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
                  "scaling": "strong", "vs_baseline": None,
                  "dtype": "f64" if args.accum == "f64" else "f32 partial sums in k_gather2d, f64 elsewhere (1e-5 mode)",
                  "data": "synthetic",
                  "config": {"workload": wl["desc"], "name": name, "particles": n_total, "npix": npix,
                             "kernel": wl["kernel"], "strategy": args.strategy, "accumulate": args.accum,
                             "l2": "inputs (%.2f GB/rank) larger than L2" % (n_loc * in_bytes / 1e9),
                             "mapped_particles": n_mapped, "particles_in": n_in_all, "pairs": pairs,
                             "footprint_pixels_all_ranks": fpx_all, "touched_pixels_all_ranks": touched_all,
                             "exchange": (None if world == 1 else
                                          "reduce(NCCL) of the two un-reduced maps to rank 0" if (healpix or stencil) else
                                          "reduce_scatter(NCCL) per plane + reduce_image division of the rank's pixel "
                                          "slice + gather on rank 0 + transposition"),
                             "exchange_ms": exchange_ms,
                             "shard": "domain_decomposition by particle id (uniform synthetic stream); work imbalance "
                                      "max/mean - 1 = %.4f" % imb},
                  "clocks": sampler.summary() if sampler else None, "e2e": e2e, "gpu_launches": launches * steps,
                  "roofline": roofline, "parity": parity, "cpu_baseline": cpu_baseline,
                  "wall_ms_per_step": wall_step}
    del d_pos, d_h, d_m, d_rho, d_T, d_one, image, out, q_t, w_t
    torch.cuda.empty_cache()
    return result


def parity_check(e, name, sample):
    """GPU (public host-array API) vs the checker oracle on the first `sample` particles of the stream, full image."""
    from oracle import oracle as orc
    s2g = e.s2g
    orc.select_library("checker")
    wl = WORKLOADS[name]
    dims, npix = wl["dims"], wl["npix"]
    cores = cpu_workers(wl)
    ref, ost, dt = oracle_map(orc, wl, sample, cores, exact=True)

    def relerr(a, b):
        a = np.asarray(a).ravel(); b = np.asarray(b).ravel()
        den = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-14 * float(np.max(np.abs(b))))
        return float(np.max(np.abs(a - b) / den))

    out = {"sample": sample, "oracle_s": round(dt, 2), "tolerance": 1e-10}
    if wl.get("stencil"):
        pos, hsml, m, rho, temp = host_particles(wl, sample)
        par = s2g.mappingParameters(center=[0, 0, 0], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=npix)
        fn = s2g.cic_deposit if wl["stencil"] == 2 else s2g.tsc_deposit
        got = fn(pos - 0.5, rho, param=par, dimensions=3, average=False, periodic=False, ctx=e.ctx)
        out.update(max_rel_err=relerr(got, ref), tolerance=1e-12, counters_equal=None,
                   against="oracle stencil (semantics defined in DESIGN.md §6)")
    elif dims == 0:
        pos, hsml, m, rho, temp = shell_sample(wl, sample)
        a, w, st = s2g.healpix_deposit(pos, hsml, m, rho, temp, rho, npix, getattr(s2g, wl["kernel"])(2), True,
                                       ctx=e.ctx, return_stats=True)
        ea, ew = ref
        worst = 0.0
        nbad = 0
        for g, x, sens in ((w, ew, ost["sens"]), (a, ea, ost["sens_q"])):
            d = np.abs(g - x); den = np.maximum(np.abs(g), np.abs(x))
            allow = 1e-10 * den + 8 * 2.220446049250313e-16 * sens
            nbad += int(np.count_nonzero(d > allow))
            worst = max(worst, float(np.max((d - 8 * 2.220446049250313e-16 * sens) / np.where(den > 0, den, 1.0))))
        out.update(max_rel_err=max(worst, 0.0), pixels_over_bar=nbad,
                   counters_equal=all(st[k] == ost[k] for k in ("n_mapped", "touched_pixels", "n_fallback")),
                   against="extended-precision arbiter (long double), bar 1e-10 + 8 ulp * sensitivity, no floor")
    else:
        pos, hsml, m, rho, temp = host_particles(wl, sample)
        par = s2g.mappingParameters(center=[0.5, 0.5, 0.5], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=npix)
        if dims == 2:
            got, st = s2g.sphMapping(pos.copy(), hsml, m, rho, temp, rho, param=par, kernel=getattr(s2g, wl["kernel"])(2),
                                     calc_mean=True, show_progress=False, ctx=e.ctx, return_stats=True)
        else:
            got, st = s2g.sphMapping(pos.copy(), hsml, m, rho, rho, np.ones_like(rho), param=par,
                                     kernel=getattr(s2g, wl["kernel"])(3), dimensions=3, show_progress=False,
                                     ctx=e.ctx, return_stats=True)
        out.update(max_rel_err=relerr(got, ref),
                   counters_equal=all(st[k] == ost[k] for k in ("n_mapped", "footprint_pixels", "touched_pixels",
                                                                "n_fallback")),
                   against="oracle port, Float64 operation by operation (floor 1e-14 of the map maximum)")
    out["ok"] = bool(out["max_rel_err"] <= out["tolerance"] and out["counters_equal"] in (True, None)
                     and out.get("pixels_over_bar", 0) == 0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("S2G_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--strategy", default="auto", choices=["auto", "scatter", "gather"])
    ap.add_argument("--accum", default="f64", choices=["f64", "f32"],
                    help="f32: the optional FP32-accumulate mode of the 2D gather kernel (1e-5 bar); not the headline")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--extra", default=os.environ.get("S2G_BENCH_EXTRA", "auto"),
                    help="auto: the default c2 run also times c3, c3cic, c3tsc, c4, c5; none: only --workload; "
                         "or a comma list of workloads")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        reference_arm(args, wl)
        return

    e = make_env(args)
    line = run_workload(e, args.workload, args.steps, max(args.warmup, 0), True)
    if args.extra == "auto":
        extras = EXTRAS if (args.workload == "c2" and args.accum == "f64" and args.strategy == "auto") else []
    elif args.extra in ("none", ""):
        extras = []
    else:
        extras = [(x, 1 if WORKLOADS[x]["n"] > (1 << 27) else 3, 1 if WORKLOADS[x]["n"] > (1 << 27) else 2)
                  for x in args.extra.split(",")]
    ex_out = []
    for name, k, w in extras:
        try:
            r = run_workload(e, name, k, w, False)
        except Exception as ex:  # an extra must never cost the headline line
            r = {"config": {"name": name}, "error": f"{type(ex).__name__}: {ex}"}
            e.torch.cuda.empty_cache()
        if e.rank == 0:
            r.pop("clocks", None)
            ex_out.append(r)
    if e.rank == 0:
        line["extra"] = ex_out
        print(json.dumps(line), flush=True)
    if e.world > 1:
        e.dist.barrier()
        e.dist.destroy_process_group()
    e.ctx.close()


if __name__ == "__main__":
    main()
