"""sphMapping and friends — host mirror of src/cic_interpolation/cic_interpolation.jl:35-273 whose hot bodies
(cic_mapping_2D/3D, reduce_image_2D/3D, center_particles, filter_particles_in_image) are calls into
libsphtogrid_cuda.so.  Same names, argument meaning and error behaviour as the Julia API.

Array conventions (Python <-> Julia): `Pos` is Julia's Matrix(3,N) *memory*, i.e. a C-contiguous numpy (N,3) array
(a Fortran-ordered (3,N) array is accepted as the same memory).  Returned maps are Fortran-ordered numpy arrays that
index exactly like the Julia results: 2D `image[ix, iy, n]`, 3D `image[iz, iy, ix]`, flat both-maps
`image[idx, plane]` with idx = ix*N + iy (0-based calculate_index) and the weight plane last.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import F32, F64, check, dbl3, default_context, lib, ptr
from .kernels import AbstractSPHKernel
from .parameters import mappingParameters, recentred_parameters


# ------------------------------------------------------------------ helpers
def _as_pos(pos):
    pos = np.asarray(pos) if not isinstance(pos, np.ndarray) else pos
    if pos.ndim != 2:
        raise ValueError("Pos must be a (3 x Npart) matrix")
    if pos.shape == (3, 3):
        view = pos.T if (pos.flags.f_contiguous and not pos.flags.c_contiguous) else pos
    elif pos.shape[0] == 3:
        view = pos.T
    elif pos.shape[1] == 3:
        view = pos
    else:
        raise ValueError("Pos must be a (3 x Npart) matrix")
    if view.dtype not in (np.float32, np.float64) or not view.flags.c_contiguous:
        raise TypeError("Pos must be float32/float64 with Julia Matrix(3,N) memory layout "
                        "(C-contiguous (N,3) or Fortran-ordered (3,N)); it is recentred in place")
    return view


def _common_dtype(pos, *arrs):
    """The ABI takes one dtype for all six arrays (Julia's promotion to Float64 is exact, so widening is lossless)."""
    if pos.dtype == np.float32 and all(np.asarray(a).dtype == np.float32 for a in arrs):
        return np.float32, F32
    return np.float64, F64


def _prep(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _binq(Bin_Quant, n, dt):
    """Julia: Vector (N) or Matrix (N_images, N) -> memory n_images x N; numpy: (N,) or (N, N_images)."""
    bq = np.asarray(Bin_Quant)
    if bq.ndim == 1:
        if bq.shape[0] != n:
            raise ValueError("Bin_Quant has the wrong length")
        return _prep(bq, dt), 1
    if bq.shape[0] == n:
        return _prep(bq, dt), bq.shape[1]
    if bq.shape[1] == n and bq.flags.f_contiguous:
        return _prep(bq.T, dt), bq.shape[0]
    raise ValueError("Bin_Quant must be (N,) or (N, N_images)")


def _kernel_id(kernel):
    if not isinstance(kernel, AbstractSPHKernel):
        raise TypeError("kernel must be an SPHKernels kernel (Cubic(), WendlandC6(2), ...)")
    return kernel.kernel_id


# ------------------------------------------------------------------ L2 replacements
def cic_mapping_2D(Pos, HSML, M, Rho, Bin_Q, Weights, RM=None, *, param: mappingParameters,
                   kernel: AbstractSPHKernel, show_progress=False, calc_mean=True, stokes=False, ctx=None,
                   return_stats=False):
    """Underlying function to map SPH data to a 2D grid (cic_2D.jl:103-244).  Returns the flat image
    (Nx*Ny, N_images+1), weight plane last.

    With `RM` (one Float64 rotation measure per particle) and `stokes=True` the particles are composited in the order
    given and the Stokes Q/U planes (images 1 and 2) of every pixel that already holds emission are Faraday-rotated
    by `mod(RM[p]*pix_weight, pi)` first (cic_2D.jl:201-217, cic_shared.jl:129-159).  `RM` without `stokes` is inert
    in the reference (faraday_rotate_pixel! then only computes the angle), and so it is here."""
    ctx = ctx or default_context()
    pos = _as_pos(Pos)
    n = pos.shape[0]
    dt, code = _common_dtype(pos, HSML, M, Rho, Bin_Q, Weights)
    bq, nim = _binq(Bin_Q, n, dt)
    npix = int(param.Npixels[0])
    image = np.zeros((npix * npix, nim + 1), order="F")
    st = _lib.Stats()
    if RM is not None:
        rm = np.asarray(RM)
        if rm.dtype != np.float64:  # faraday_rotate_pixel!(…, pRM::Float64, …) has no other method
            raise TypeError("MethodError: no method matching faraday_rotate_pixel!(::Matrix{Float64}, ::Int64, "
                            f"::{rm.dtype}, ::Float64, ::Bool): RM must be Float64")
        rm = np.ascontiguousarray(rm)
        if rm.shape[0] < n:
            raise IndexError(f"BoundsError: attempt to access {rm.shape[0]}-element Vector{{Float64}} at index [{n}]")
        check(lib().s2g_deposit_2d_rm(ctx.handle, ptr(_prep(pos, dt)), ptr(_prep(HSML, dt)), ptr(_prep(M, dt)),
                                      ptr(_prep(Rho, dt)), ptr(bq), ptr(_prep(Weights, dt)), ptr(rm), n, nim, code,
                                      float(param.len2pix), npix, int(param.Npixels[1]), _kernel_id(kernel),
                                      int(calc_mean), int(bool(stokes)), ptr(image), C.byref(st)))
        return (image, st.asdict()) if return_stats else image
    check(lib().s2g_deposit_2d(ctx.handle, ptr(_prep(pos, dt)), ptr(_prep(HSML, dt)), ptr(_prep(M, dt)),
                               ptr(_prep(Rho, dt)), ptr(bq), ptr(_prep(Weights, dt)), n, nim, code,
                               float(param.len2pix), npix, int(param.Npixels[1]), _kernel_id(kernel), int(calc_mean),
                               ptr(image), C.byref(st)))
    return (image, st.asdict()) if return_stats else image


def cic_mapping_3D(Pos, HSML, M, Rho, Bin_Q, Weights, *, param: mappingParameters, kernel: AbstractSPHKernel,
                   show_progress=False, calc_mean=False, ctx=None, return_stats=False):
    """Underlying function to map SPH data to a 3D grid (cic_3D.jl:110-209).  Returns the flat image (N^3, 2)."""
    ctx = ctx or default_context()
    pos = _as_pos(Pos)
    n = pos.shape[0]
    dt, code = _common_dtype(pos, HSML, M, Rho, Bin_Q, Weights)
    npix = int(param.Npixels[0])
    image = np.zeros((npix ** 3, 2), order="F")
    st = _lib.Stats()
    check(lib().s2g_deposit_3d(ctx.handle, ptr(_prep(pos, dt)), ptr(_prep(HSML, dt)), ptr(_prep(M, dt)),
                               ptr(_prep(Rho, dt)), ptr(_prep(Bin_Q, dt)), ptr(_prep(Weights, dt)), n, code,
                               float(param.len2pix), npix, _kernel_id(kernel), int(calc_mean), ptr(image),
                               C.byref(st)))
    return (image, st.asdict()) if return_stats else image


def reduce_image_2D(image, x_pixels, y_pixels, reduce_image, ctx=None):
    """Unflattens an image array to a 2D array of pixels (reduce_image.jl:8-31)."""
    ctx = ctx or default_context()
    image = np.asfortranarray(image, dtype=np.float64)
    nim = image.shape[1] - 1
    out = np.zeros((y_pixels, x_pixels, nim), order="F")
    check(lib().s2g_reduce_image_2d(ctx.handle, ptr(image), int(x_pixels), int(y_pixels), nim, int(bool(reduce_image)),
                                    ptr(out)))
    return out


def reduce_image_3D(image, x_pixels, y_pixels, z_pixels, reduce_image=True, ctx=None):
    """Unflattens an image array to a 3D array of pixels (reduce_image.jl:39-55).  `reduce_image=False` reproduces
    `image[:,2] .= 1.0` of cic_interpolation.jl:230-232."""
    ctx = ctx or default_context()
    image = np.asfortranarray(image, dtype=np.float64)
    out = np.zeros((z_pixels, y_pixels, x_pixels), order="F")
    check(lib().s2g_reduce_image_3d(ctx.handle, ptr(image), int(x_pixels), int(bool(reduce_image)), ptr(out)))
    return out


def center_particles(x, par: mappingParameters, ctx=None):
    """Shifts all particles so that the image is centered on [0, 0, 0] (filter_shift.jl:6-32).  In place, in the
    precision of `x`; returns (x, recentred parameters)."""
    ctx = ctx or default_context()
    pos = _as_pos(x)
    code = F32 if pos.dtype == np.float32 else F64
    par2 = recentred_parameters(par)
    out = np.empty_like(pos)
    check(lib().s2g_center_filter(ctx.handle, ptr(pos), pos.shape[0], code, dbl3(par.center), int(par.periodic),
                                  float(par.boxsize), dbl3([0, 0, 0]), dbl3(par2.halfsize), ptr(out), None))
    pos[...] = out
    return x, par2


def filter_particles_in_image(pos, par: mappingParameters, sort_z: bool = False, ctx=None):
    """Checks if a particle is contained in the image and returns an array of Bool (filter_shift.jl:40-67).
    With sort_z the reference returns `reverse(sortperm(z))[mask]` — reproduced literally (0-based indices)."""
    ctx = ctx or default_context()
    p = _as_pos(pos)
    code = F32 if p.dtype == np.float32 else F64
    if not np.all(par.center == 0.0):
        # the reference compares against center -/+ halfsize; shift-free call with the box moved instead
        lo, hi = par.center - par.halfsize, par.center + par.halfsize
        mask = np.all((lo[None, :] <= p) & (p <= hi[None, :]), axis=1)
    else:
        m8 = np.zeros(p.shape[0], dtype=np.uint8)
        check(lib().s2g_center_filter(ctx.handle, ptr(p), p.shape[0], code, dbl3([0, 0, 0]), 0, -1.0, dbl3([0, 0, 0]),
                                      dbl3(par.halfsize), None, ptr(m8)))
        mask = m8.astype(bool)
    if sort_z:
        srt = np.argsort(p[:, 2], kind="stable")[::-1]
        return srt[mask]
    return mask


def domain_decomposition(N: int, N_workers: int):
    """Calculate relevant array slices for each worker (parallel/domain_decomp.jl:7-17); 0-based half-open ranges."""
    size = int(np.floor(N / N_workers))
    batch = [(i * size, (i + 1) * size) for i in range(N_workers - 1)]
    batch.append(((N_workers - 1) * size, N))
    return batch


# ------------------------------------------------------------------ weight presets (weight_functions.jl:10-51)
def part_weight_one(N):
    return np.ones(N)


def part_weight_physical(N, par=None, x_cgs=3.085678e21):
    if par is None or not isinstance(par, mappingParameters):
        if par is not None:
            x_cgs = par
        return np.ones(N) * x_cgs
    return np.ones(N) * par.pixelSideLength * x_cgs


def part_weight_emission(rho, T_K):
    return np.asarray(rho) ** 2 * np.sqrt(np.asarray(T_K))


def part_weight_spectroscopic(rho, T_K):
    return np.asarray(rho) ** 2 / np.sqrt(np.sqrt(np.asarray(T_K))) ** 3


# ------------------------------------------------------------------ public entry
def sphMapping(Pos, HSML, M, Rho, Bin_Quant, Weights=None, RM=None, *, param: mappingParameters,
               kernel: AbstractSPHKernel, show_progress: bool = True, parallel: bool = False,
               reduce_image: bool = True, return_both_maps: bool = False, dimensions: int = 2,
               calc_mean: bool = False, stokes: bool = False, sort_z: bool = False, ctx=None, return_stats=False,
               group=None, _projection=None):
    """Maps the data in `Bin_Quant` to a grid (cic_interpolation.jl:35-273).

    `Pos` is recentred IN PLACE like the reference does.  `parallel=True` shards the particles with
    `domain_decomposition` and sums the partial flat images before the division (`image = sum(fetch.(futures))`,
    :199/:256): with `group=DeviceGroup(...)` over the GPUs of that group inside this one process (host thread per
    device, peer-memory sum fused with reduce_image: s2g_group_sphmap), otherwise over the ranks of an initialised
    torch.distributed process group (one GPU per rank, NCCL all-reduce)."""
    if stokes:
        # cic_interpolation.jl:74-81: back-to-front order, serial.  NB the reference does NOT forward `stokes` to
        # cic_mapping_2D (:152-155), so through sphMapping the RM never rotates anything; reproduced: the map is
        # the z-sorted deposit.  Call cic_mapping_2D(..., RM, stokes=True) directly for the Faraday compositing.
        sort_z = True
        parallel = False
    if dimensions not in (2, 3):
        return None  # the reference falls through both branches and returns nothing
    ctx = ctx or default_context()
    if Weights is None:
        Weights = Rho
    pos = _as_pos(Pos)
    n = pos.shape[0]
    dt, code = _common_dtype(pos, HSML, M, Rho, Bin_Quant, Weights)
    bq, nim = _binq(Bin_Quant, n, dt)
    if dimensions == 3 and nim != 1:
        raise ValueError("3D mapping takes a single quantity")
    par = recentred_parameters(param)
    npix = int(par.Npixels[0])
    kid = _kernel_id(kernel)
    hs, mm, rr, ww = _prep(HSML, dt), _prep(M, dt), _prep(Rho, dt), _prep(Weights, dt)

    if _projection is not None and (parallel or sort_z or pos.dtype != dt):
        # map_it's projection on a path that is not the fused single-call one: rotate the (copied) positions on the
        # host like the reference does, then continue unprojected
        from .rotate import _permute, apply_matrix
        perm, rot = _projection
        _projection = None
        if perm is not None:
            _permute(perm, pos)
        else:
            pos = apply_matrix(rot, pos)
            dt, code = _common_dtype(pos, HSML, M, Rho, Bin_Quant, Weights)
            bq, nim = _binq(Bin_Quant, n, dt)
            hs, mm, rr, ww = _prep(HSML, dt), _prep(M, dt), _prep(Rho, dt), _prep(Weights, dt)

    if parallel and group is not None:
        both = bool(return_both_maps) and dimensions == 2
        if dimensions == 2:
            out = np.zeros((npix * npix, nim + 1), order="F") if both else np.zeros((npix, npix, nim), order="F")
        else:
            out = np.zeros((npix, npix, npix), order="F")
        if pos.dtype != dt:
            # Float32 positions with Float64 fields: recentre first in Float32 (Q2), then map without a further shift
            center_particles(pos, param, ctx=group.context(0))
            shift, periodic, boxsize, p_in, p_out = [0, 0, 0], 0, -1.0, _prep(pos, dt), None
        else:
            shift, periodic, boxsize, p_in = param.center, int(param.periodic), float(param.boxsize), pos
            p_out = np.empty_like(pos)
        gst = group.new_stats()
        check(lib().s2g_group_sphmap(group.handle, dimensions, ptr(p_in), ptr(hs), ptr(mm), ptr(rr), ptr(bq), ptr(ww), n,
                                     nim, code, dbl3(shift), periodic, boxsize, dbl3(par.halfsize), float(par.len2pix),
                                     npix, kid, int(calc_mean), int(bool(reduce_image)), int(both), ptr(p_out), ptr(out),
                                     gst))
        if p_out is not None:
            pos[...] = p_out
        return (out, [g.asdict() for g in gst]) if return_stats else out

    if parallel:
        from .distributed import sph_mapping_sharded
        return sph_mapping_sharded(ctx, pos, hs, mm, rr, bq, ww, nim, code, param, par, kid, dimensions, calc_mean,
                                   reduce_image, return_both_maps)

    if sort_z:
        # Q5: the reference indexes the z-sorted permutation with the UNSORTED mask; reproduce on the host
        center_particles(pos, param, ctx=ctx)
        sel = filter_particles_in_image(pos, par, True, ctx=ctx)
        x = np.ascontiguousarray(pos[sel])
        bqs = np.ascontiguousarray(bq[sel])
        if dimensions == 2:
            rms = None if RM is None else np.ascontiguousarray(np.asarray(RM)[sel])
            image = cic_mapping_2D(x, hs[sel], mm[sel], rr[sel], bqs, ww[sel], rms, param=par, kernel=kernel,
                                   calc_mean=calc_mean, ctx=ctx)
            if return_both_maps:
                return image
            return reduce_image_2D(image, int(param.Npixels[0]), int(param.Npixels[1]), reduce_image, ctx=ctx)
        image = cic_mapping_3D(x, hs[sel], mm[sel], rr[sel], bqs, ww[sel], param=par, kernel=kernel, ctx=ctx)
        return reduce_image_3D(image, npix, npix, npix, reduce_image, ctx=ctx)

    both = bool(return_both_maps) and dimensions == 2
    if dimensions == 2:
        out = np.zeros((npix * npix, nim + 1), order="F") if both else np.zeros((npix, npix, nim), order="F")
    else:
        out = np.zeros((npix, npix, npix), order="F")
    st = _lib.Stats()
    if _projection is not None:
        # map_it(projection=...): permutation / rotation applied inside the position load of the same fused call;
        # the positions are a private copy of map_it, nothing is written back
        perm, rot = _projection
        cperm = (C.c_int32 * 3)(*perm) if perm is not None else None
        crot = (C.c_double * 9)(*np.asarray(rot, dtype=np.float64).ravel()) if rot is not None else None
        check(lib().s2g_sphmap_projected(ctx.handle, dimensions, ptr(pos), ptr(hs), ptr(mm), ptr(rr), ptr(bq), ptr(ww),
                                         n, nim, code, cperm, crot, dbl3(param.center), int(param.periodic),
                                         float(param.boxsize), dbl3(par.halfsize), float(par.len2pix), npix, kid,
                                         int(calc_mean), int(bool(reduce_image)), int(both), None, ptr(out),
                                         C.byref(st)))
    elif pos.dtype == dt:
        # fused: centre (in the precision of Pos) + filter + deposit + reduce on the device; the recentred
        # positions come back so that the caller's Pos is mutated like in the reference (Q1)
        pos_out = np.empty_like(pos)
        check(lib().s2g_sphmap(ctx.handle, dimensions, ptr(pos), ptr(hs), ptr(mm), ptr(rr), ptr(bq), ptr(ww), n, nim,
                               code, dbl3(param.center), int(param.periodic), float(param.boxsize),
                               dbl3(par.halfsize), float(par.len2pix), npix, kid, int(calc_mean),
                               int(bool(reduce_image)), int(both), ptr(pos_out), ptr(out), C.byref(st)))
        pos[...] = pos_out
    else:
        # Float32 positions with Float64 fields: recentre first in Float32 (Q2), then map without a further shift
        center_particles(pos, param, ctx=ctx)
        check(lib().s2g_sphmap(ctx.handle, dimensions, ptr(_prep(pos, dt)), ptr(hs), ptr(mm), ptr(rr), ptr(bq),
                               ptr(ww), n, nim, code, dbl3([0, 0, 0]), 0, -1.0, dbl3(par.halfsize),
                               float(par.len2pix), npix, kid, int(calc_mean), int(bool(reduce_image)), int(both),
                               None, ptr(out), C.byref(st)))
    return (out, st.asdict()) if return_stats else out


def map_it(pos_in, hsml, mass, rho, bin_q, weights, RM=None, *, param: mappingParameters, kernel=None, snap=0,
           units="", image_prefix="dummy", reduce_image=True, parallel=True, calc_mean=True, show_progress=True,
           sort_z=False, stokes=False, renorm=False, projection="xy", write_fits=True):
    """Small helper function to copy positions, map particles and save the fits file
    (cic_interpolation.jl:312-384).  Writes `image_prefix.<projection>.fits` like the reference and also returns the
    map.  `parallel=True` shards over the ranks of an initialised process group (serial otherwise)."""
    from .kernels import WendlandC6
    kernel = kernel or WendlandC6(2)
    pos = np.array(_as_pos(pos_in), copy=True)   # pos = copy(pos_in): the caller's positions stay untouched
    from .rotate import projection_of
    perm, rot, par, projection = projection_of(projection, param)   # cic_interpolation.jl:331-345
    proj = None if (perm is None and rot is None) else (perm, rot)
    par_ok, dist = False, None
    if parallel:
        # torch is plumbing for the multi-process path only: serial use must work with numpy + ctypes alone
        try:
            import torch.distributed as dist
            par_ok = dist.is_available() and dist.is_initialized()
        except ImportError:
            par_ok = False
    m = sphMapping(pos, hsml, mass, rho, bin_q, weights, RM, param=par, kernel=kernel, show_progress=show_progress,
                   parallel=par_ok, reduce_image=reduce_image, calc_mean=calc_mean, sort_z=sort_z, stokes=stokes,
                   _projection=proj)
    if renorm:
        m /= np.max(m)
    if write_fits and (not par_ok or dist.get_rank() == 0):
        from .io import write_fits_image
        write_fits_image(f"{image_prefix}.{projection}.fits", m, param, snap=snap, units=units)
    return m
