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
                        return_both_maps):
    """Body of sphMapping(parallel=true) on this rank's GPU.  Every rank holds the full input (like the reference's
    master) and deposits its own contiguous slice; every rank returns the full result."""
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
        allreduce_sum_(image)  # image = sum(fetch.(futures))
        if dimensions == 2 and return_both_maps:
            out = image.cpu().numpy().reshape((ncell, planes), order="F")
        else:
            red = torch.empty(ncell * (nim if dimensions == 2 else 1), dtype=torch.float64, device=dev)
            if dimensions == 2:
                check(lib().s2g_reduce_image_2d_dev(ctx.handle, ptr(image.data_ptr()), npix, npix, nim,
                                                    int(bool(reduce_image)), ptr(red.data_ptr())))
                out = red.cpu().numpy().reshape((npix, npix, nim), order="F")
            else:
                check(lib().s2g_reduce_image_3d_dev(ctx.handle, ptr(image.data_ptr()), npix, int(bool(reduce_image)),
                                                    ptr(red.data_ptr())))
                out = red.cpu().numpy().reshape((npix, npix, npix), order="F")
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
