"""Multi-GPU mapping: one process per GPU (torchrun), particles sharded by `domain_decomposition`, partial flat images
summed with an NCCL all-reduce over NVLink before the `reduce_image` division.

Replaces the `parallel=true` branch of sphMapping (src/cic_interpolation/cic_interpolation.jl:171-215, :236-271:
`futures[i] = @spawnat id cic_mapping_2D(x[:,batch[i]], ...)`; `image = sum(fetch.(futures))`) and the master-side
accumulation of distributed_cic_map / distributed_allsky_map (src/distributed_mapping/cic.jl:58-70,
healpix.jl:40-52).  torch.distributed is plumbing only (rendezvous + the collective); the deposit is the C ABI.

The only exchange step of this path is the sum of the partial images: (n_images+1)*N^2 doubles (2D), 2*N^3 (3D),
2*12*Nside^2 (HEALPix).  At 8192^2 that is 1.07 GB per GPU, ~3 ms at the measured 725 GB/s all-reduce bus bandwidth
against seconds of deposit, so it is NOT fused into the deposit kernel.
"""
from __future__ import annotations

import numpy as np

from ._lib import check, dbl3, lib, ptr
from .mapping import domain_decomposition


def _dist():
    import torch.distributed as dist
    return dist


def world():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_range(n: int, world_size: int, rank: int):
    """Contiguous particle range of `rank` (parallel/domain_decomp.jl:7-17)."""
    return domain_decomposition(n, world_size)[rank]


def allreduce_sum_(tensor):
    """In-place sum over ranks (NCCL on CUDA tensors, gloo on CPU tensors); no-op without a process group."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def _reduce_scatter_plane(out, plane):
    """out (len/ws elements) = this rank's slice of the sum over ranks of `plane`."""
    dist = _dist()
    try:
        dist.reduce_scatter_tensor(out, plane, op=dist.ReduceOp.SUM)
    except (RuntimeError, NotImplementedError):
        # backends without reduce-scatter (gloo on the CPU tests): same result through an all-reduce
        tmp = plane.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
        r, n = dist.get_rank(), out.numel()
        out.copy_(tmp[r * n:(r + 1) * n])


def _all_gather_plane(out, mine):
    dist = _dist()
    try:
        dist.all_gather_into_tensor(out, mine)
    except (RuntimeError, NotImplementedError):
        parts = [out[i * mine.numel():(i + 1) * mine.numel()] for i in range(dist.get_world_size())]
        tmp = [torch_empty_like(mine) for _ in parts]
        dist.all_gather(tmp, mine)
        for a, b in zip(parts, tmp):
            a.copy_(b)


def torch_empty_like(t):
    import torch
    return torch.empty_like(t)


def exchange_reduce(image, n_images, ncell, dims, reduce_image, divide, work=None, gather="all"):
    """The one exchange step of the path, `image = sum(fetch.(futures))` (cic_interpolation.jl:199, :256), fused with the
    reduce_image division: the flat partial image ((n_images+1) planes of `ncell`, weight plane last) is
    REDUCE-SCATTERED plane by plane, so rank r receives the summed planes of pixel slice r only (half the traffic of an
    all-reduce), divides its slice (`divide(q_slice, w_slice, n, stride, n_images, dims, reduce_image)` — on the device:
    s2g_divide_slice_dev), and the quantity slices are gathered ("all": on every rank, like the reference's master
    returning the map to every caller of the sharded API; "root": on rank 0 only).
    Returns the reduced flat quantity planes (n_images * ncell, NOT transposed) or None on non-root ranks.
    Falls back to all-reduce + whole-image division when ncell is not divisible by the world size."""
    import torch
    dist = _dist()
    ws, rank = world()
    planes = n_images + 1
    if ws == 1 or ncell % ws != 0:
        if ws > 1:
            dist.all_reduce(image, op=dist.ReduceOp.SUM)
        divide(image[:n_images * ncell], image[n_images * ncell:], ncell, ncell, n_images, dims, reduce_image)
        return image[:n_images * ncell]
    sl = ncell // ws
    work = work if work is not None else {}
    mine = work.get("mine")
    if mine is None or mine.numel() != planes * sl:
        mine = work["mine"] = torch.empty(planes * sl, dtype=image.dtype, device=image.device)
    for k in range(planes):
        _reduce_scatter_plane(mine[k * sl:(k + 1) * sl], image[k * ncell:(k + 1) * ncell])
    divide(mine[:n_images * sl], mine[n_images * sl:], sl, sl, n_images, dims, reduce_image)
    if gather == "root" and rank != 0:
        for k in range(n_images):
            dist.gather(mine[k * sl:(k + 1) * sl], None, dst=0)
        return None
    out = work.get("out")
    if out is None or out.numel() != n_images * ncell:
        out = work["out"] = torch.empty(n_images * ncell, dtype=image.dtype, device=image.device)
    for k in range(n_images):
        if gather == "root":
            dist.gather(mine[k * sl:(k + 1) * sl], [out[k * ncell + i * sl:k * ncell + (i + 1) * sl] for i in range(ws)],
                        dst=0)
        else:
            _all_gather_plane(out[k * ncell:(k + 1) * ncell], mine[k * sl:(k + 1) * sl])
    return out


def device_divide(ctx):
    """`divide` callback of exchange_reduce for CUDA tensors: s2g_divide_slice_dev on the context's stream."""
    def divide(q, w, n, stride, n_images, dims, reduce_image):
        check(lib().s2g_divide_slice_dev(ctx.handle, int(dims), ptr(q.data_ptr()), ptr(w.data_ptr()), int(n),
                                         int(stride), int(n_images), int(bool(reduce_image))))
    return divide


def footprint_balanced_decomposition(footprints, world_size):
    """Contiguous particle ranges with (nearly) equal SUMS of footprint pixels instead of equal particle counts — the
    work of a particle is its footprint (cic_shared.jl:46-52 gives the box), so clustered inputs shard evenly.
    `domain_decomposition` (parallel/domain_decomp.jl:7-17) stays the literal, count-balanced mode.
    Returns [(start, end)] * world_size covering [0, N)."""
    f = np.asarray(footprints, dtype=np.float64)
    n = f.shape[0]
    if n == 0 or world_size <= 1:
        return [(0, n)] + [(n, n)] * (world_size - 1)
    c = np.cumsum(np.maximum(f, 1.0))          # every particle costs at least its set-up
    cuts = np.searchsorted(c, c[-1] * np.arange(1, world_size) / world_size, side="left") + 1
    cuts = np.minimum(np.maximum.accumulate(cuts), n)
    edges = [0] + [int(x) for x in cuts] + [n]
    return [(edges[i], edges[i + 1]) for i in range(world_size)]


def combine_partial_images(partial: np.ndarray, finite_guard: bool = False) -> np.ndarray:
    """Host-array front end of the exchange step (used by the gloo tests and by streaming accumulation):
    sums `partial` over ranks.  finite_guard reproduces distributed_mapping/cic.jl:63-69 (NaN/Inf entries of a
    partial map are skipped)."""
    import torch
    partial = np.asarray(partial, dtype=np.float64)
    f_order = partial.flags.f_contiguous and not partial.flags.c_contiguous
    t = torch.from_numpy(partial.ravel(order="F" if f_order else "C").copy())  # memory order
    if finite_guard:
        t = torch.where(torch.isfinite(t), t, torch.zeros_like(t))
    allreduce_sum_(t)
    return t.numpy().reshape(partial.shape, order="F" if f_order else "C")


def sph_mapping_sharded(ctx, pos, hs, mm, rr, bq, ww, nim, code, param, par, kid, dimensions, calc_mean, reduce_image,
                        return_both_maps, balance="footprint"):
    """Body of sphMapping(parallel=true) on this rank's GPU.  Every rank holds the full input (like the reference's
    master) and deposits its own contiguous slice; every rank returns the full result.
    balance = "footprint" (default): slices of equal summed footprint (s2g_footprints + prefix sum);
              "count": the reference's domain_decomposition (parallel/domain_decomp.jl:7-17)."""
    import torch
    from .mapping import center_particles
    ws, rank = world()
    n = pos.shape[0]
    s, e = shard_range(n, ws, rank)
    want = np.float32 if code == 0 else np.float64
    for name, a in (("HSML", hs), ("M", mm), ("Rho", rr), ("Bin_Quant", bq), ("Weights", ww)):
        if a.dtype != want:
            raise TypeError(f"sph_mapping_sharded: {name} is {a.dtype}, the call was declared {np.dtype(want)} "
                            "(sphMapping converts the fields before sharding)")
    shift, periodic, boxsize = param.center, int(param.periodic), float(param.boxsize)
    recentre_after = True
    if pos.dtype != want:
        # Float32 positions with Float64 fields (the common Gadget case): recentre in Float32 like the reference (Q2),
        # then deposit the widened copy without a further shift — s2g_sphmap_dev reads ONE dtype for all arrays
        center_particles(pos, param, ctx=ctx)
        pos_up = np.ascontiguousarray(pos[s:e], dtype=want)
        shift, periodic, boxsize, recentre_after = [0.0, 0.0, 0.0], 0, -1.0, False
    else:
        pos_up = pos[s:e]
    npix = int(par.Npixels[0])
    ncell = npix * npix if dimensions == 2 else npix ** 3
    planes = nim + 1 if dimensions == 2 else 2
    if balance == "footprint" and ws > 1 and n > 0:
        # every rank computes the same split from the same (full) arrays: pix_index_min_max boxes of the particles as
        # they will be deposited (positions recentred like the deposit recentres them)
        from .mapping import center_particles as _cp
        pc = np.array(pos, copy=True)
        if recentre_after:
            _cp(pc, param, ctx=ctx)
        b = np.zeros((n, 2 * dimensions), dtype=np.int64)
        check(lib().s2g_footprints(ctx.handle, ptr(np.ascontiguousarray(pc, dtype=want)), ptr(hs), n, code,
                                   float(par.len2pix), npix, dimensions, ptr(b)))
        fp = np.ones(n)
        for d_ in range(dimensions):
            fp *= np.maximum(b[:, 2 * d_ + 1] - b[:, 2 * d_] + 1, 0)
        s, e = footprint_balanced_decomposition(fp, ws)[rank]
        pos_up = np.ascontiguousarray(pos[s:e], dtype=want) if not recentre_after else pos[s:e]
    dev = torch.device("cuda", ctx.device)
    with torch.cuda.device(dev):
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a[s:e])).to(dev, non_blocking=False)
        d_pos = torch.from_numpy(np.ascontiguousarray(pos_up)).to(dev, non_blocking=False)
        d_hs, d_m, d_r, d_q, d_w = up(hs), up(mm), up(rr), up(bq), up(ww)
        image = torch.zeros(ncell * planes, dtype=torch.float64, device=dev)
        check(lib().s2g_sphmap_dev(ctx.handle, dimensions, ptr(d_pos.data_ptr()), ptr(d_hs.data_ptr()),
                                   ptr(d_m.data_ptr()), ptr(d_r.data_ptr()), ptr(d_q.data_ptr()), ptr(d_w.data_ptr()),
                                   e - s, nim, code, dbl3(shift), periodic, boxsize,
                                   dbl3(par.halfsize), float(par.len2pix), npix, kid, int(calc_mean), 0,
                                   ptr(image.data_ptr())))
        if dimensions == 2 and return_both_maps:
            allreduce_sum_(image)  # image = sum(fetch.(futures)); both maps go back undivided
            out = image.cpu().numpy().reshape((ncell, planes), order="F")
        else:
            # reduce-scatter + per-rank division of its pixel slice + gather (exchange_reduce), then the transposition
            nq = nim if dimensions == 2 else 1
            flat = exchange_reduce(image, nq, ncell, dimensions, reduce_image, device_divide(ctx))
            if dimensions == 2:
                red = torch.empty(ncell * nim, dtype=torch.float64, device=dev)
                check(lib().s2g_reduce_image_2d_dev(ctx.handle, ptr(flat.data_ptr()), npix, npix, nim, 0,
                                                    ptr(red.data_ptr())))   # reduce_image = 0: transposition only
                out = red.cpu().numpy().reshape((npix, npix, nim), order="F")
            else:
                out = flat.cpu().numpy().reshape((npix, npix, npix), order="F")
        torch.cuda.current_stream(dev).synchronize()
    # Q1: the reference recentres the caller's Pos before slicing
    if recentre_after:
        center_particles(pos, param, ctx=ctx)
    return out


class StreamingAccumulator:
    """Resident image + weight accumulators fed batch by batch — the device-side equivalent of the sub-snapshot loop
    of distributed_cic_map (src/distributed_mapping/cic.jl:24-110): per-file partial maps are summed with the
    reference's finite guard, then reduced once at the end."""

    def __init__(self, n_elements: int):
        self.sum = np.zeros(n_elements)

    def add(self, local: np.ndarray):
        loc = np.asarray(local, dtype=np.float64).reshape(-1, order="F")
        ok = np.isfinite(loc)
        self.sum[ok] += loc[ok]

    def result_over_ranks(self):
        return combine_partial_images(self.sum)


def _my_jobs(n_subfiles: int):
    """Sub-snapshot files of this rank: the reference's dynamic job queue (distributed_mapping/main.jl:10-26) hands
    job ids 0..n-1 to whichever worker is free; with one rank per GPU a round-robin split is the static equivalent."""
    ws, rank = world()
    return range(rank, n_subfiles, ws)


def _results(jobs, mapping_function, loader):
    """`mapping_function(subfile)` per job; with a `loader`, sub-file k+1 is read on a background thread while the
    device works on sub-file k and the function is called as `mapping_function(subfile, loader(subfile))`."""
    if loader is None:
        for job in jobs:
            yield mapping_function(job)
    else:
        from .gadget import SnapshotPrefetcher
        for job, data in SnapshotPrefetcher(jobs, loader):
            yield mapping_function(job, data)


def distributed_cic_map(cic_filename, Nsubfiles, mapping_function, param, Nimages=1, *, reduce_image=True, Ndim=2,
                        snap=0, units="", vtk=True, write=True, loader=None):
    """Mirror of distributed_cic_map (src/distributed_mapping/cic.jl:24-110): one map per sub-snapshot file
    (`mapping_function(subfile) -> (map(s), weight_map)`, e.g. the two parts of
    `sphMapping(...; return_both_maps=true)`), summed with the reference's NaN/Inf guard, reduced once, saved.
    Ranks split the files; the partial sums are combined with one all-reduce.  Returns the reduced image.
    `loader(subfile)` (optional, no reference counterpart) separates the file read from the mapping so that the next
    sub-file is read while the current one is deposited."""
    from .mapping import reduce_image_2D, reduce_image_3D
    n = int(param.Npixels[0])
    n_distr = n ** Ndim
    acc_map = StreamingAccumulator(n_distr * Nimages)
    acc_w = StreamingAccumulator(n_distr)
    for res in _results(_my_jobs(Nsubfiles), mapping_function, loader):
        if res is None or res[0] is None:
            continue
        local_map, local_weight = res
        acc_map.add(local_map)
        acc_w.add(local_weight)
    sum_map = acc_map.result_over_ranks().reshape((n_distr, Nimages), order="F")
    sum_w = acc_w.result_over_ranks()
    flat = np.asfortranarray(np.concatenate([sum_map, sum_w[:, None]], axis=1))
    if Ndim == 2:
        image = reduce_image_2D(flat, n, n, reduce_image)
    elif Ndim == 3:
        image = reduce_image_3D(flat, n, n, n, True)  # the reference always divides in 3D (cic.jl:87)
    else:
        raise ValueError("Only 2D and 3D are possible")
    if write and world()[1] == 0:
        if Ndim == 3 and vtk:
            from .io import write_vtk_image
            write_vtk_image(cic_filename, image, "map", param, snap=snap, units=units)  # cic.jl:96-98
        else:
            from .io import write_fits_image
            write_fits_image(cic_filename, image, param, snap=snap, units=units)
    return image


def distributed_allsky_map(allsky_filename, Nside, Nsubfiles, mapping_function, *, reduce_image=True, write=False,
                           loader=None):
    """Mirror of distributed_allsky_map (src/distributed_mapping/healpix.jl:16-87).  `mapping_function(subfile)` returns
    `(allsky_map, weight_map)` of `healpix_map`.  Returns `(sum_allsky, sum_weights)` (the first divided by the second
    where the weight is finite and non-zero when reduce_image)."""
    npix = 12 * int(Nside) ** 2
    acc_a, acc_w = StreamingAccumulator(npix), StreamingAccumulator(npix)
    for a, w in _results(_my_jobs(Nsubfiles), mapping_function, loader):
        acc_a.add(a)
        acc_w.add(w)
    sum_a, sum_w = acc_a.result_over_ranks(), acc_w.result_over_ranks()
    if reduce_image:
        ok = np.isfinite(sum_w) & (sum_w != 0)
        sum_a[ok] /= sum_w[ok]
    if write and world()[1] == 0:
        from .io import save_healpix_fits
        save_healpix_fits(allsky_filename, sum_a)  # saveToFITS(sum_allsky, allsky_filename), healpix.jl:73-77
    return sum_a, sum_w
