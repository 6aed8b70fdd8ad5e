#!/usr/bin/env python
"""bench.py — headline benchmark of the particle-deposition hot path (BASELINE.json metric: Mparticles/s mapped).

  python bench.py --gpus N --steps K --warmup W            (under torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    (the reference algorithm's CPU path, oracle port)

A step = one `sphMapping` pass over the whole synthetic particle set of the workload:
centre + filter + Smac deposit (+ NCCL sum of the partial images for N > 1) + reduce_image.
Default workload "c2" = BASELINE.json configs[1]: 16 777 216 Gadget-like particles, 4096^2 map, WendlandC6(2),
calc_mean mass-weighted temperature map (q = T, w = rho, reduce_image = true).

value : whole-job Mparticles/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks.
e2e   : same step through the C ABI with pinned HOST buffers, H2D of the inputs and D2H of the map inside the region.
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
    # name: particles, npix, dims, kernel, n_ngb, seed   
    "c2": dict(n=16 * 1024 * 1024, npix=4096, dims=2, kernel="WendlandC6", n_ngb=295.0, seed=2,
               desc="synthetic 16M-particle box, 2D calc_mean mass-weighted T map, 4096^2, WendlandC6"),
    "c3": dict(n=64 * 1024 * 1024, npix=512, dims=3, kernel="Cubic", n_ngb=64.0, seed=3,
               desc="synthetic 64M particles, 3D sphMapping onto 512^3, Cubic"),
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


def reference_arm(args, wl):
    """The reference's own CPU implementation of the path (oracle port of cic_mapping_2D/3D with `parallel=true`
    slicing over all host threads), on a bounded sample of the same workload."""
    from oracle import oracle as orc
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or (131072 if wl["dims"] == 2 else 1 << 20)
    pos, hsml, m, rho, temp = host_particles(wl, sample)
    par = orc.mapping_parameters(center=[0.5, 0.5, 0.5], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=wl["npix"])
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        p = pos.copy()
        if wl["dims"] == 2:
            orc.sph_mapping(p, hsml, m, rho, temp, rho, param=par, kernel=wl["kernel"], parallel=True,
                            n_workers=cores, calc_mean=True, reduce_image=True)
        else:
            orc.sph_mapping(p, hsml, m, rho, rho, np.ones_like(rho), param=par, kernel=wl["kernel"], parallel=True,
                            n_workers=cores, dimensions=3, reduce_image=True)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = sample / t / 1e6
    line = {"impl": "reference", "metric": "Mparticles/s mapped", "value": val, "unit": "Mparticles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "particles": wl["n"], "npix": wl["npix"], "kernel": wl["kernel"]},
            "cpu_baseline": {"value": val, "unit": "Mparticles/s", "cores": cores, "kind": "port",
                             "sample": f"first {sample} particles of the same synthetic stream, full-size image"},
            "e2e": {"value": val, "unit": "Mparticles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


_HOST_CACHE = {}


def host_particles(wl, count):
    """First `count` particles of the workload's stream on the host.  Generated on the GPU when there is one (the
    stream is defined by the device generator); without a GPU, a numpy Philox-free stand-in with the same recipe."""
    key = (wl["seed"], count)
    if key in _HOST_CACHE:
        return _HOST_CACHE[key]
    import __graft_entry__ as ge
    s2g = ge.load_package()
    if s2g.lib().s2g_device_count() > 0:
        import torch
        from sphtogrid_b200 import _lib
        ctx = s2g.default_context()
        dev = torch.device("cuda", ctx.device)
        t = [torch.empty(count * 3 if i == 0 else count, dtype=torch.float64, device=dev) for i in range(5)]
        _lib.check(s2g.lib().s2g_synth_particles_dev(ctx.handle, wl["seed"], 0, count, wl["n"], 1.0, wl["n_ngb"],
                                                     SIGMA, 1, *[_lib.ptr(x.data_ptr()) for x in t]))
        ctx.sync()
        out = [x.cpu().numpy() for x in t]
        out[0] = out[0].reshape(count, 3)
    else:
        rng = np.random.default_rng(wl["seed"])
        pos = rng.random((count, 3))
        g = rng.normal(size=count)
        rho_bar = 1.0
        rho = rho_bar * np.exp(SIGMA * g - 0.5 * SIGMA ** 2)
        mass = np.full(count, 1.0 / wl["n"])
        hsml = np.cbrt(3.0 * wl["n_ngb"] * mass / (4.0 * np.pi * rho))
        temp = 1e4 * (rho / rho_bar) ** (2.0 / 3.0) * np.exp(0.5 * rng.normal(size=count))
        out = [pos, hsml, mass, rho, temp]
    _HOST_CACHE[key] = tuple(out)
    return _HOST_CACHE[key]


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
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    s2g = ge.load_package()
    from sphtogrid_b200 import _lib
    L = s2g.lib()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = s2g.Context(local_rank, strategy=args.strategy)
    if args.accum != "f64":
        ctx.set_accumulate_mode(args.accum)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)

    # ---- this rank's shard of the global particle stream (domain_decomposition over particle ids)
    n_total = wl["n"]
    s, e = s2g.domain_decomposition(n_total, world)[rank]
    n_loc = e - s
    dims, npix = wl["dims"], wl["npix"]
    f64 = torch.float64
    d_pos = torch.empty(n_loc * 3, dtype=f64, device=dev)
    d_h, d_m, d_rho, d_T = (torch.empty(n_loc, dtype=f64, device=dev) for _ in range(4))
    P = lambda t: _lib.ptr(t.data_ptr())
    _lib.check(L.s2g_synth_particles_dev(ctx.handle, wl["seed"], s, n_loc, wl.get("n_stream", n_total), 1.0,
                                         wl["n_ngb"], SIGMA, 1, P(d_pos), P(d_h), P(d_m), P(d_rho), P(d_T)))
    d_one = torch.ones(n_loc, dtype=f64, device=dev) if dims == 3 else None
    par = s2g.mappingParameters(center=[0.5, 0.5, 0.5], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=npix)
    par2 = s2g.recentred_parameters(par)
    kid = getattr(s2g, wl["kernel"])(KERNEL_DIM[dims]).kernel_id
    healpix = dims == 0
    ncell = npix ** dims if not healpix else 12 * npix * npix
    planes = 2
    image = torch.empty(ncell * planes, dtype=f64, device=dev)
    out = torch.empty(ncell, dtype=f64, device=dev)
    q_t, w_t = (d_T, d_rho) if dims != 3 else (d_rho, d_one)
    shift, half = _lib.dbl3(par.center), _lib.dbl3(par2.halfsize)

    hp_center = (C.c_double * 3)(0.5, 0.5, 0.5)          # observer at the box centre
    hp_shell = (C.c_double * 2)(0.05, 0.5)               # radius_limits = [0.05 L, 0.5 L]
    hp_nsel = C.c_int64(0)

    def step_healpix():
        # the whole healpix_map body on the device: Pos .-= center, shell filter, far-to-near selection
        # (filter_sort_particles incl. its sorted[mask] semantics), particle loop
        _lib.check(L.s2g_healpix_map_dev(ctx.handle, P(d_pos), P(d_h), P(d_m), P(d_rho), P(q_t), P(w_t), n_loc,
                                         hp_center, hp_shell, npix, kid, 1, 0, P(image), P(image[ncell:]),
                                         C.byref(hp_nsel)))
        if world > 1:
            dist.all_reduce(image)

    def step_device():
        if healpix:
            return step_healpix()
        _lib.check(L.s2g_sphmap_dev(ctx.handle, dims, P(d_pos), P(d_h), P(d_m), P(d_rho), P(q_t), P(w_t), n_loc, 1, 1,
                                    shift, 0, -1.0, half, float(par2.len2pix), npix, kid, 1, 0, P(image)))
        if world > 1:
            dist.all_reduce(image)
        if dims == 2:
            _lib.check(L.s2g_reduce_image_2d_dev(ctx.handle, P(image), npix, npix, 1, 1, P(out)))
        else:
            _lib.check(L.s2g_reduce_image_3d_dev(ctx.handle, P(image), npix, 1, P(out)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
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
        return float(t[0]) / steps, float(t[1]) / steps

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step, wall_step = timed(step_device, args.steps, args.warmup, sampler)
    # stats of one more deposit (the last library call of a step is the reduce) for the phase breakdown
    if healpix:
        step_healpix()
    else:
        _lib.check(L.s2g_sphmap_dev(ctx.handle, dims, P(d_pos), P(d_h), P(d_m), P(d_rho), P(q_t), P(w_t), n_loc, 1, 1,
                                    shift, 0, -1.0, half, float(par2.len2pix), npix, kid, 1, 0, P(image)))
    st = ctx.stats()
    cnt = torch.tensor([st["n_mapped"], st["footprint_pixels"], st["touched_pixels"], st["n_pairs"],
                        st["n_launches"] + 1], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    n_mapped, fpx_all, touched_all, pairs, launches = [int(x) for x in cnt.tolist()]
    # the roofline describes ONE kernel launch on ONE GPU: use this rank's own counters there
    fpx, touched = int(st["footprint_pixels"]), int(st["touched_pixels"])
    value = n_mapped / (ms_step * 1e-3) / 1e6

    # ---- end to end: pinned host buffers -> H2D -> step -> D2H of the reduced map, all inside the timed region
    e2e = None
    if not args.no_e2e and not healpix:
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
        h_pos, h_h, h_m, h_rho, h_q, h_w = pin(d_pos), pin(d_h), pin(d_m), pin(d_rho), pin(q_t), pin(w_t)
        h_out = torch.empty(ncell, dtype=f64, pin_memory=True)
        torch.cuda.synchronize(dev)
        if world == 1:
            hp = lambda t: _lib.ptr(t.data_ptr())

            def step_e2e():
                _lib.check(L.s2g_sphmap(ctx.handle, dims, hp(h_pos), hp(h_h), hp(h_m), hp(h_rho), hp(h_q), hp(h_w),
                                        n_loc, 1, 1, shift, 0, -1.0, half, float(par2.len2pix), npix, kid, 1, 1, 0,
                                        None, hp(h_out), None))
        else:
            def step_e2e():
                for dt_, ht in ((d_pos, h_pos), (d_h, h_h), (d_m, h_m), (d_rho, h_rho), (q_t, h_q), (w_t, h_w)):
                    dt_.copy_(ht, non_blocking=True)
                step_device()
                if rank == 0:
                    h_out.copy_(out, non_blocking=True)
                stream.synchronize()
        ms_e2e, wall_e2e = timed(step_e2e, max(1, args.steps), 1)
        e2e = {"value": n_mapped / (max(ms_e2e, wall_e2e) * 1e-3) / 1e6, "unit": "Mparticles/s",
               "h2d_bytes_per_step": int(n_total * 8 * 8), "d2h_bytes_per_step": int(ncell * 8),
               "ms_per_step": max(ms_e2e, wall_e2e)}

    if rank == 0:
        pk, pk_kind = peaks()
        hbm_peak = float(pk.get("hbm_gbs", 6650.0))
        # dominant kernel: the deposit phase (k_gather2d / k_scatter*); algorithmic bytes per map (SURVEY §8d roof 1)
        dep_ms = st["ms_deposit"] if st["ms_deposit"] > 0 else ms_step
        # launches of the dominant kernel per step on this rank: the 2D gather walks the shard in slices of 8 Mi
        # particles (S2G_BATCH_PARTICLES), one k_gather2d launch each; 3D / HEALPix: one launch per step
        n_dom = max(1, -(-n_loc // int(os.environ.get("S2G_BATCH_PARTICLES", 8 << 20)))) if dims == 2 else 1
        launch_ms = dep_ms / n_dom
        # algorithmic bytes per launch (SURVEY §8d roof 1): every particle field of the slice read once
        # (8 scalars x 8 B) + every image plane written once
        alg_bytes = (n_loc // n_dom) * 64 + ncell * planes * 8
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if args.workload == "c2" and world == 1 and os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]  # per launch, one ncu --set full capture
        r = C.c_double(0)
        _lib.check(L.s2g_microbench(ctx.handle, 0, 0, 20000, C.byref(r)))
        fp64_peak = r.value  # GFLOP/s, DFMA microbenchmark, same process, same clocks
        _lib.check(L.s2g_microbench(ctx.handle, 1, 256 << 20, 2000, C.byref(r)))
        red_peak = r.value   # Gred/s, 32 consecutive doubles per warp
        flop_per_px = {2: 40.0, 3: 45.0, 0: 80.0}[dims]   # SURVEY §8d roof 3: pass A + pass B per footprint pixel
        fp64_ach = fpx * flop_per_px / (dep_ms * 1e-3 + st["ms_norm"] * 1e-3) / 1e9
        atom_time_ms = planes * touched / (red_peak * 1e9) * 1e3
        roofline = {"bound": "hbm", "achieved": alg_bytes / (launch_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / (launch_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic, "peak_kind": pk_kind,
                    "kernel": {2: "k_gather2d", 3: "k_scatter3d", 0: "k_healpix"}[dims],
                    "launches_per_step": n_dom, "launch_ms": launch_ms, "alg_bytes_per_launch": alg_bytes,
                    "note": "this path is FP64-issue bound (gather) / L2-red bound (scatter), not HBM bound: the HBM "
                            "fraction is reported because the contract asks for it; see the fp64 and atomic roofs "
                            "(SURVEY.md §8d, DESIGN.md §4)",
                    "fp64": {"achieved_gflops": fp64_ach, "peak_gflops": fp64_peak, "frac": fp64_ach / fp64_peak,
                             "flop_per_footprint_pixel": flop_per_px, "footprint_pixels": fpx,
                             "phase_ms": {"norm": st["ms_norm"], "deposit": dep_ms, "sort": st["ms_sort"],
                                          "prep": st["ms_prep"]}},
                    "atomic": {"reds": planes * touched, "peak_gred_s": red_peak, "t_atomic_ms": atom_time_ms,
                               "frac_of_step": atom_time_ms / ms_step}}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline and not healpix:
            from oracle import oracle as orc
            cores = os.cpu_count() or 1
            sample = args.cpu_sample or (131072 if dims == 2 else 1 << 20)
            hp_, hh_, hm_, hr_, ht_ = host_particles(wl, sample)
            opar = orc.mapping_parameters(center=[0.5, 0.5, 0.5], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=npix)
            t0 = time.perf_counter()
            if dims == 2:
                orc.sph_mapping(hp_.copy(), hh_, hm_, hr_, ht_, hr_, param=opar, kernel=wl["kernel"], parallel=True,
                                n_workers=cores, calc_mean=True, reduce_image=True)
            else:
                orc.sph_mapping(hp_.copy(), hh_, hm_, hr_, hr_, np.ones_like(hr_), param=opar, kernel=wl["kernel"],
                                parallel=True, n_workers=cores, dimensions=3, reduce_image=True)
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": sample / dt / 1e6, "unit": "Mparticles/s", "cores": cores, "kind": "port",
                            "sample": f"first {sample} particles of the same stream, full-size image, {dt:.1f} s"}
        line = {"metric": "Mparticles/s mapped", "value": value, "unit": "Mparticles/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": "f64" if args.accum == "f64" else "f32 partial sums in k_gather2d, f64 elsewhere (1e-5 mode)",
                "data": "synthetic",
                "config": {"workload": wl["desc"], "particles": n_total, "npix": npix, "kernel": wl["kernel"],
                           "strategy": args.strategy, "accumulate": args.accum, "l2": "inputs (%.2f GB/rank) larger than L2" % (n_loc * 64 / 1e9),
                           "mapped_particles": n_mapped, "pairs": pairs,
                           "footprint_pixels_all_ranks": fpx_all, "touched_pixels_all_ranks": touched_all},
                "clocks": sampler.summary() if sampler else None, "e2e": e2e, "gpu_launches": launches * args.steps,
                "roofline": roofline, "cpu_baseline": cpu_baseline, "wall_ms_per_step": wall_step}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
