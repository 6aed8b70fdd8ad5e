"""CIC / TSC grid assignment.  The reference holds only commented-out code for this
(src/tsc_interpolation/tsc_interpolation.jl:1-183, wrapping an external TSCInterpolation(...; average=true)); the
semantics implemented here are defined in DESIGN.md ("parity unpinned")."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import F32, F64, check, default_context, lib, ptr
from .mapping import _as_pos
from .parameters import mappingParameters


def _stencil(order, Pos, Bin_Quant, param: mappingParameters, dimensions, average, periodic, ctx):
    ctx = ctx or default_context()
    pos = _as_pos(Pos)
    dt, code = (np.float32, F32) if (pos.dtype == np.float32 and np.asarray(Bin_Quant).dtype == np.float32) \
        else (np.float64, F64)
    npix = int(param.Npixels[0])
    ncell = npix ** dimensions
    image = np.zeros((ncell, 2), order="F")
    st = _lib.Stats()
    check(lib().s2g_stencil_deposit(ctx.handle, order, dimensions, ptr(np.ascontiguousarray(pos, dtype=dt)),
                                    ptr(np.ascontiguousarray(Bin_Quant, dtype=dt)), pos.shape[0], code,
                                    float(param.len2pix), npix, int(bool(periodic)), ptr(image), C.byref(st)))
    if not average:
        return image
    wv = image[:, 1]
    out = np.where(wv > 0.0, image[:, 0] / np.where(wv > 0.0, wv, 1.0), image[:, 0])
    return out.reshape((npix,) * dimensions)


def cic_deposit(Pos, Bin_Quant, *, param, dimensions=3, average=True, periodic=False, ctx=None):
    """Cloud-in-cell assignment; positions relative to the image centre (as after center_particles)."""
    return _stencil(2, Pos, Bin_Quant, param, dimensions, average, periodic, ctx)


def tsc_deposit(Pos, Bin_Quant, *, param, dimensions=3, average=True, periodic=False, ctx=None):
    """Triangular-shaped-cloud assignment."""
    return _stencil(3, Pos, Bin_Quant, param, dimensions, average, periodic, ctx)
