"""healpix_map — host mirror of src/healpix_interpolation/main.jl:92-227; the particle loop is
libsphtogrid_cuda's s2g_healpix_deposit.  filter_sort_particles (filter_particles.jl:17-54) is O(N log N) host
logic and is reproduced literally, including its quirks (in-place recentre, sorted[mask], BoundsError)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import F64, check, default_context, lib, ptr
from .mapping import _as_pos, _kernel_id


def find_in_shell(dx, radius_limits):
    """filter_particles.jl:6-8"""
    return (radius_limits[0] <= dx) & (dx <= radius_limits[1])


def filter_sort_particles(Pos, Hsml, M, Rho, Bin_q, Weights, center, radius_limits, calc_mean):
    """filter_particles.jl:17-54"""
    pos = _as_pos(Pos)
    pos -= np.asarray(center, dtype=pos.dtype)[None, :]                      # Pos .-= center (in place)
    dx = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2 + pos[:, 2] ** 2)
    sel = find_in_shell(dx, radius_limits)
    if not calc_mean:
        sel = sel[np.asarray(Bin_q)[sel] > 0.0]
    srt = np.argsort(dx, kind="stable")[::-1]                                # reverse(sortperm(Δx))
    if sel.shape[0] != srt.shape[0]:
        raise IndexError(f"BoundsError: attempt to access {srt.shape[0]}-element Vector{{Int64}} at index "
                         f"[{sel.shape[0]}-element BitVector]")
    idx = srt[sel]
    g = lambda a: np.ascontiguousarray(np.asarray(a)[idx])
    return np.ascontiguousarray(pos[idx]), g(Hsml), g(M), g(Rho), g(Bin_q), g(Weights)


def healpix_deposit(pos, hsml, m, rho, bin_q, weights, Nside, kernel, calc_mean=True, ctx=None, return_stats=False):
    """The particle loop of healpix_map (main.jl:143-213) on already filtered, observer-centred particles."""
    ctx = ctx or default_context()
    p = np.ascontiguousarray(_as_pos(pos), dtype=np.float64)
    n = p.shape[0]
    npix = 12 * int(Nside) ** 2
    amap = np.zeros(npix); wmap = np.zeros(npix)
    st = _lib.Stats()
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    check(lib().s2g_healpix_deposit(ctx.handle, ptr(p), ptr(f(hsml)), ptr(f(m)), ptr(f(rho)), ptr(f(bin_q)),
                                    ptr(f(weights)), n, F64, int(Nside), _kernel_id(kernel), int(calc_mean),
                                    ptr(amap), ptr(wmap), C.byref(st)))
    return (amap, wmap, st.asdict()) if return_stats else (amap, wmap)


def healpix_map(Pos, Hsml, M, Rho, Bin_q, Weights, *, center=(0.0, 0.0, 0.0), radius_limits=(0.0, np.inf),
                Nside=1024, kernel, show_progress=True, output_from_all_workers=False, calc_mean=True, ctx=None,
                group=None, strict_reference=True):
    """Calculate an allsky map from SPH particles.  Returns `(image, weight_image)` in RING order (0-based storage =
    Julia pixels[i+1]); divide to reduce the image (main.jl:76-77).  `group=DeviceGroup(...)` spreads the particle loop
    over the GPUs of the group (same result: the shell selection is made over all particles, s2g_group_healpix_map).

    `strict_reference=True` (default) is bug-compatible with filter_sort_particles (filter_particles.jl:28-41): the
    shell mask in ORIGINAL order indexes the far-to-near permutation (`sorted[sel]`), so as soon as one particle lies
    outside the shell the `count(sel)` deposited particles are picked by radial rank, not by membership (a
    `RuntimeWarning` says so), and `calc_mean=False` raises the reference's BoundsError unless every particle is in
    the shell with `Bin_q > 0`.  `strict_reference=False` (no reference counterpart) maps exactly the particles
    inside the shell (with `Bin_q > 0` when `calc_mean=False`), far to near."""
    npix = 12 * int(Nside) ** 2
    if not strict_reference:
        pos = _as_pos(Pos)
        if pos.dtype != np.float64:
            raise TypeError("healpix_map requires Float64 inputs (method signatures `where T`, pixel_weights.jl:87-91)")
        pos -= np.asarray(center, dtype=np.float64)[None, :]
        dx = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2 + pos[:, 2] ** 2)
        keep = find_in_shell(dx, radius_limits)
        if not calc_mean:
            keep &= np.asarray(Bin_q) > 0.0
        idx = np.flatnonzero(keep)
        idx = idx[np.argsort(dx[idx], kind="stable")[::-1]]
        g = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64)[idx])
        return healpix_deposit(np.ascontiguousarray(pos[idx]), g(Hsml), g(M), g(Rho), g(Bin_q), g(Weights), Nside,
                               kernel, calc_mean, ctx=ctx or (group.context(0) if group is not None else None))
    if (not calc_mean) and np.sum(Bin_q) == 0:
        return np.zeros(npix), np.zeros(npix)
    pos = _as_pos(Pos)
    if pos.dtype != np.float64:
        raise TypeError("healpix_map requires Float64 inputs (method signatures `where T`, pixel_weights.jl:87-91)")
    if group is None:
        ctx = ctx or default_context()
    n = pos.shape[0]
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    hs, mm, rr, bq, ww = f(Hsml), f(M), f(Rho), f(Bin_q), f(Weights)
    if not calc_mean:
        # filter_particles.jl:28-30: `sel = sel[Bin_q[sel] .> 0.0]` shortens the mask; indexing the permutation with
        # it is a BoundsError unless every particle is in the shell AND has Bin_q > 0 (then nothing changes)
        dx = np.sqrt((pos[:, 0] - center[0]) ** 2 + (pos[:, 1] - center[1]) ** 2 + (pos[:, 2] - center[2]) ** 2)
        sel = find_in_shell(dx, radius_limits)
        n_short = int(np.count_nonzero(bq[sel] > 0.0))
        if n_short != n:
            pos -= np.asarray(center, dtype=np.float64)[None, :]  # the reference has already recentred Pos by then
            raise IndexError(f"BoundsError: attempt to access {n}-element Vector{{Int64}} at index "
                             f"[{n_short}-element BitVector]")
    amap = np.zeros(npix); wmap = np.zeros(npix)
    pos_out = np.empty_like(pos)
    if np.isfinite(radius_limits[1]) or radius_limits[0] > 0.0:
        dx_ = np.sqrt((pos[:, 0] - center[0]) ** 2 + (pos[:, 1] - center[1]) ** 2 + (pos[:, 2] - center[2]) ** 2)
        n_out = n - int(np.count_nonzero(find_in_shell(dx_, radius_limits)))
        if n_out:
            import warnings
            warnings.warn(f"healpix_map: {n_out} of {n} particles lie outside radius_limits; the reference indexes the "
                          "far-to-near permutation with the unsorted shell mask (filter_particles.jl:33-41), so the "
                          "deposited particles are chosen by radial rank, not by shell membership — reproduced here; "
                          "pass strict_reference=False to map the particles inside the shell", RuntimeWarning,
                          stacklevel=2)
    cen = (C.c_double * 3)(*[float(c) for c in center])
    rl = (C.c_double * 2)(float(radius_limits[0]), float(radius_limits[1]))
    if group is not None:
        check(lib().s2g_group_healpix_map(group.handle, ptr(pos), ptr(hs), ptr(mm), ptr(rr), ptr(bq), ptr(ww), n, cen,
                                          rl, int(Nside), _kernel_id(kernel), int(calc_mean), ptr(pos_out), ptr(amap),
                                          ptr(wmap), group.new_stats()))
        pos[...] = pos_out
        return amap, wmap
    st = _lib.Stats()
    check(lib().s2g_healpix_map(ctx.handle, ptr(pos), ptr(hs), ptr(mm), ptr(rr), ptr(bq), ptr(ww), n, cen, rl,
                                int(Nside), _kernel_id(kernel), int(calc_mean), ptr(pos_out), ptr(amap), ptr(wmap),
                                C.byref(st)))
    pos[...] = pos_out  # Pos .-= center, in place (filter_particles.jl:20)
    return amap, wmap
