"""mappingParameters — mirror of src/shared/parameters.jl:14-125 (pure host scalar logic, stays on the CPU)."""
from __future__ import annotations

import math

import numpy as np


class mappingParameters:
    """Parameter object for sph to grid mapping.  Define either `*_lim`, or `center` and `*_size`.
    Resolution is defined by `pixelSideLength` or `Npixels` (parameters.jl:44-54)."""

    __slots__ = ("x_lim", "y_lim", "z_lim", "center", "halfsize", "len2pix", "pixelSideLength", "Npixels", "boxsize",
                 "periodic")

    def __init__(self, T=float, *, x_lim=(-1.0, -1.0), y_lim=(-1.0, -1.0), z_lim=(-1.0, -1.0),
                 center=(-1.0, -1.0, -1.0), x_size=-1.0, y_size=-1.0, z_size=-1.0, pixelSideLength=-1.0, Npixels=0,
                 boxsize=-1.0):
        x_lim = [float(v) for v in x_lim]; y_lim = [float(v) for v in y_lim]; z_lim = [float(v) for v in z_lim]
        center = [float(v) for v in center]
        x_size, y_size, z_size = float(x_size), float(y_size), float(z_size)
        pixelSideLength = float(pixelSideLength)
        dflt2, dflt3 = [-1.0, -1.0], [-1.0, -1.0, -1.0]
        # parameters.jl:58-69
        if x_lim == dflt2 and (y_lim == dflt2 and z_lim == dflt2):
            if center != dflt3 and (x_size != -1.0 and (y_size != -1.0 and z_size != -1.0)):
                x_lim = [center[0] - 0.5 * x_size, center[0] + 0.5 * x_size]
                y_lim = [center[1] - 0.5 * y_size, center[1] + 0.5 * y_size]
                z_lim = [center[2] - 0.5 * z_size, center[2] + 0.5 * z_size]
            else:
                raise ValueError("Giving a center position requires extent in x, y and z direction.")
        # :72-80
        if x_size == -1.0:
            x_size = x_lim[1] - x_lim[0]
        if y_size == -1.0:
            y_size = y_lim[1] - y_lim[0]
        if z_size == -1.0:
            z_size = z_lim[1] - z_lim[0]
        # :82-86
        if center == dflt3:
            center = [x_lim[0] + 0.5 * x_size, y_lim[0] + 0.5 * y_size, z_lim[0] + 0.5 * z_size]
        max_size = max(x_size, y_size)  # :89
        # :91-99
        if (pixelSideLength == -1.0) and (Npixels != 0):
            pixelSideLength = max_size / Npixels
        elif (pixelSideLength != -1.0) and (Npixels == 0):
            Npixels = int(math.floor(max_size / pixelSideLength))
            pixelSideLength = max_size / Npixels
        else:
            raise ValueError("Please specify pixelSideLength or number of pixels!")
        self.x_lim = np.array(x_lim); self.y_lim = np.array(y_lim); self.z_lim = np.array(z_lim)
        self.center = np.array(center)
        self.halfsize = 0.5 * np.array([x_size, y_size, z_size])          # :113
        self.len2pix = 1.0 / pixelSideLength                              # :115
        self.pixelSideLength = pixelSideLength
        self.Npixels = np.array([int(Npixels)] * 3, dtype=np.int64)       # :105
        self.boxsize = float(boxsize)
        self.periodic = boxsize != -1.0                                   # :107-111

    def __repr__(self):
        return (f"mappingParameters(x_lim={self.x_lim.tolist()}, y_lim={self.y_lim.tolist()}, "
                f"z_lim={self.z_lim.tolist()}, center={self.center.tolist()}, Npixels={int(self.Npixels[0])}, "
                f"pixelSideLength={self.pixelSideLength}, boxsize={self.boxsize})")


def recentred_parameters(par: mappingParameters) -> mappingParameters:
    """The `par` that center_particles rebuilds (src/cic_interpolation/filter_shift.jl:24-31): limits shifted by the
    centre, centre [0,0,0], Npixels = maximum(par.Npixels), pixel size recomputed from the shifted limits."""
    cen = par.center
    return mappingParameters(center=[0.0, 0.0, 0.0], x_lim=par.x_lim - cen[0], y_lim=par.y_lim - cen[1],
                             z_lim=par.z_lim - cen[2], Npixels=int(par.Npixels.max()), boxsize=par.boxsize)
