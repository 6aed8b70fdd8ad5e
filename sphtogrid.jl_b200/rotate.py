"""Projection pre-step of `map_it`: axis swaps and Euler rotations of particle positions and of the map parameters
(src/shared/rotate_particles.jl:7-180, src/shared/rotate_parameters.jl:27-59).

Host mirrors for callers that want the rotated arrays themselves; `map_it(projection=...)` does NOT use them on the
particle data — it hands the permutation / matrix to `s2g_sphmap_projected`, which applies it inside the position
load of the deposit kernels (no extra pass over the 3xN array, no rotated copy).

Positions are (N,3) arrays here (= the memory of Julia's Matrix(3,N)).  Julia's `f!(x)` / `f(par)` method pairs are
one polymorphic function each: an ndarray argument is rotated IN PLACE and returned, a mappingParameters argument
gives the rotated parameters."""
from __future__ import annotations

import math

import numpy as np

from .parameters import mappingParameters

PERM_XY = (0, 1, 2)
PERM_XZ = (0, 2, 1)   # rotate_to_xz_plane!: (x, y, z) <- (x, z, y)       rotate_particles.jl:35-43
PERM_YZ = (1, 2, 0)   # rotate_to_yz_plane!: (x, y, z) <- (y, z, x)       rotate_particles.jl:65-74


def euler_matrix(alpha, beta, gamma):
    """RotXYZ(deg2rad(alpha), deg2rad(beta), deg2rad(gamma)) of Rotations.jl = Rx(alpha)*Ry(beta)*Rz(gamma), as the
    row-major 3x3 Float64 matrix `rotate_3D` multiplies with (rotate_particles.jl:7-13).  Rotations.jl is a third-party
    dependency that is not part of the reference tree: the element formulas below are the textbook product, their last
    ulp against Rotations.jl is unpinned."""
    t1, t2, t3 = math.radians(alpha), math.radians(beta), math.radians(gamma)
    s1, c1 = math.sin(t1), math.cos(t1)
    s2, c2 = math.sin(t2), math.cos(t2)
    s3, c3 = math.sin(t3), math.cos(t3)
    return np.array([[c2 * c3, -c2 * s3, s2],
                     [s1 * s2 * c3 + c1 * s3, c1 * c3 - s1 * s2 * s3, -s1 * c2],
                     [s1 * s3 - c1 * s2 * c3, c1 * s2 * s3 + s1 * c3, c1 * c2]])


def apply_matrix(rot, x):
    """rot * x for (N,3) positions, Float64 result; each component is (r0*x0 + r1*x1) + r2*x2 with individually
    rounded operations — the same expression the device evaluates in `ld_pos`."""
    x = np.asarray(x)
    x0, x1, x2 = (x[:, k].astype(np.float64) for k in range(3))
    out = np.empty((x.shape[0], 3))
    for d in range(3):
        out[:, d] = (rot[d, 0] * x0 + rot[d, 1] * x1) + rot[d, 2] * x2
    return out


def rotate_3D(x, alpha, beta, gamma):
    """Rotates an array of 3D positions around the Euler angles alpha, beta, gamma (degrees) — rotations around the
    x, y and z axis (rotate_particles.jl:7-13).  Returns a new Float64 array."""
    return apply_matrix(euler_matrix(alpha, beta, gamma), x)


def rotate_3D_(x, alpha, beta, gamma):
    """`rotate_3D!` (rotate_particles.jl:22-30): despite the name the reference rebinds the local and returns a new
    array, the argument stays untouched."""
    return rotate_3D(x, alpha, beta, gamma)


def _rotated_parameters(par, perm):
    lims = (par.x_lim, par.y_lim, par.z_lim)
    return mappingParameters(center=[par.center[perm[0]], par.center[perm[1]], par.center[perm[2]]],
                             x_lim=lims[perm[0]].copy(), y_lim=lims[perm[1]].copy(), z_lim=lims[perm[2]].copy(),
                             Npixels=int(max(par.Npixels)), boxsize=par.boxsize)


def _permute(perm, x, x_in=None):
    if isinstance(x, mappingParameters):
        return _rotated_parameters(x, perm)  # rotate_parameters.jl:27-59
    src = x.copy() if x_in is None else np.asarray(x_in)
    for d in range(3):
        x[:, d] = src[:, perm[d]]
    return x


def rotate_to_xz_plane(x, x_in=None):
    """ndarray: rotates the positions into the xz-plane in place (rotate_particles.jl:35-58);
    mappingParameters: the parameters of that projection (rotate_parameters.jl:27-40)."""
    return _permute(PERM_XZ, x, x_in)


def rotate_to_yz_plane(x, x_in=None):
    """ndarray: rotates the positions into the yz-plane in place (rotate_particles.jl:65-89; the two-argument method
    of the reference only loops over the first 3 particles — `size(x,1)` — which is not reproduced);
    mappingParameters: the parameters of that projection (rotate_parameters.jl:48-59)."""
    return _permute(PERM_YZ, x, x_in)


def project_along_axis(x, projection_axis=3, x_in=None):
    """Projects positions along one of the principal axes: 3 -> xy (nothing to do), 2 -> xz, 1 -> yz
    (rotate_particles.jl:97-180)."""
    if projection_axis == 3:
        return x if x_in is None else x_in
    if projection_axis == 2:
        return rotate_to_xz_plane(x, x_in)
    if projection_axis == 1:
        return rotate_to_yz_plane(x, x_in)
    return None  # the reference falls through


def projection_of(projection, param):
    """(perm | None, rot | None, rotated parameters, file-name tag) for map_it's `projection` argument
    (cic_interpolation.jl:331-345)."""
    if isinstance(projection, str):
        if projection == "xy":
            return None, None, param, "xy"
        if projection == "xz":
            return PERM_XZ, None, rotate_to_xz_plane(param), "xz"
        if projection == "yz":
            return PERM_YZ, None, rotate_to_yz_plane(param), "yz"
    elif isinstance(projection, (list, tuple, np.ndarray)) and len(projection) == 3:
        a, b, g = (float(v) for v in projection)
        # NB the reference leaves `par` unassigned in this branch (UndefVarError at the sphMapping call, "not used
        # yet!" in its docstring); the evident intent — rotate the particles, keep the parameters — is implemented.
        return None, euler_matrix(a, b, g), param, "alpha=%0.2fbeta=%0.2fgamma=%0.2f" % (a, b, g)
    raise ValueError("projection must be either along in 'xy', 'xz', or 'yz' plane of defined by a vector of Euler "
                     "angles!")
