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
    ws, rank = world()
    n = pos.shape[0]
    s, e = shard_range(n, ws, rank)
    npix = int(par.Npixels[0])
    ncell = npix * npix if dimensions == 2 else npix ** 3
    planes = nim + 1 if dimensions == 2 else 2
    dev = torch.device("cuda", ctx.device)
    with torch.cuda.device(dev):
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a[s:e])).to(dev, non_blocking=False)
        d_pos, d_hs, d_m, d_r, d_q, d_w = up(pos), up(hs), up(mm), up(rr), up(bq), up(ww)
        image = torch.zeros(ncell * planes, dtype=torch.float64, device=dev)
        check(lib().s2g_sphmap_dev(ctx.handle, dimensions, ptr(d_pos.data_ptr()), ptr(d_hs.data_ptr()),
                                   ptr(d_m.data_ptr()), ptr(d_r.data_ptr()), ptr(d_q.data_ptr()), ptr(d_w.data_ptr()),
                                   e - s, nim, code, dbl3(param.center), int(param.periodic), float(param.boxsize),
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
    from .mapping import center_particles
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
