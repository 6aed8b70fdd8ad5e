"""
oracle.py — ctypes front-end of the CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT.

Loads oracle/libs2g_oracle.so (built from oracle/s2g_oracle.c, see oracle/Makefile)
and restates the reference's host orchestration on top of it:

  mapping_parameters  <- src/shared/parameters.jl:44-125
  sph_mapping         <- src/cic_interpolation/cic_interpolation.jl:35-273
  healpix_map         <- src/healpix_interpolation/main.jl:92-227 (+ filter_particles.jl:17-54)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  Parity pinning status: see the header of s2g_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libs2g_oracle.so")

KERNEL_IDS = {"Cubic": 0, "Quintic": 1, "WendlandC2": 2, "WendlandC4": 3, "WendlandC6": 4, "WendlandC8": 5}


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("s2g_oracle.c", "s2g_oracle_exact.c")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_libs = {}
_which = "checker"
_fast_path = None
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


def select_library(which="checker", native=True):
    """"checker" (default): libs2g_oracle.so, -O2 -ffp-contract=off — the ONLY build parity is judged with.
    "fast": libs2g_oracle_fast.so, -O3 with FMA contraction — the CPU-baseline timing leg of bench.py; with
    `native` it is first rebuilt with -march=native for the host it runs on (falls back to the shipped x86-64-v3
    build when there is no compiler)."""
    global _which
    if which not in ("checker", "fast"):
        raise ValueError(which)
    global _fast_path
    if which == "fast" and "fast" not in _libs and native:
        # built OUTSIDE the tree: a -march=native binary must not travel to another host with the repo snapshot
        import tempfile
        out = os.path.join(tempfile.gettempdir(), "libs2g_oracle_fast_native_%d.so" % os.getuid())
        try:
            subprocess.run(["/usr/bin/gcc", "-O3", "-march=native", "-std=gnu11", "-fPIC", "-fno-fast-math",
                            "-fvisibility=hidden", "-fopenmp", "-shared", "-o", out,
                            os.path.join(_HERE, "s2g_oracle.c"), os.path.join(_HERE, "s2g_oracle_exact.c"),
                            "-lquadmath", "-lm"], check=True, capture_output=True)
            _fast_path = out
        except Exception:
            _fast_path = None
    _which = which
    return lib()


def lib():
    if _which not in _libs:
        if _which == "checker":
            build()
            path = _LIB_PATH
        else:
            path = _fast_path or os.path.join(_HERE, "libs2g_oracle_fast.so")
            if not os.path.exists(path):
                subprocess.run(["make", "-C", _HERE, "fast"], check=True, capture_output=True)
        L = C.CDLL(path)
        L.s2go_kernel_shape.restype = C.c_double
        L.s2go_kernel_shape.argtypes = [C.c_int, C.c_double]
        L.s2go_kernel_value.restype = C.c_double
        L.s2go_kernel_value.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.s2go_calculate_index_2d.restype = C.c_int64
        L.s2go_calculate_index_2d.argtypes = [C.c_int64] * 3
        L.s2go_calculate_index_3d.restype = C.c_int64
        L.s2go_calculate_index_3d.argtypes = [C.c_int64] * 5
        L.s2go_mapping_parameters.restype = C.c_int
        L.s2go_mapping_parameters.argtypes = [_dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                              C.c_int64, _dp, _ip]
        L.s2go_center_particles_f64.argtypes = [_dp, C.c_int64, _dp, C.c_int, C.c_double]
        L.s2go_center_particles_f32.argtypes = [_fp, C.c_int64, _dp, C.c_int, C.c_double]
        L.s2go_filter_particles_f64.argtypes = [_dp, C.c_int64, _dp, _dp, _bp]
        L.s2go_filter_particles_f32.argtypes = [_fp, C.c_int64, _dp, _dp, _bp]
        L.s2go_domain_decomposition.argtypes = [C.c_int64, C.c_int64, _ip, _ip]
        L.s2go_cic_mapping_2d_rm.restype = C.c_int
        L.s2go_cic_mapping_2d_rm.argtypes = [_dp] * 7 + [C.c_int64, C.c_int, C.c_double, C.c_int64, C.c_int, C.c_int,
                                                         C.c_int, C.c_int, _dp, _ip]
        L.s2go_cic_mapping_2d.restype = C.c_int
        L.s2go_cic_mapping_2d.argtypes = [_dp] * 6 + [C.c_int64, C.c_int, C.c_double, C.c_int64, C.c_int, C.c_int,
                                                      C.c_int, _dp, _ip, _ip]
        L.s2go_cic_mapping_3d.restype = C.c_int
        L.s2go_cic_mapping_3d.argtypes = [_dp] * 6 + [C.c_int64, C.c_double, C.c_int64, C.c_int, C.c_int, C.c_int,
                                                      _dp, _ip, _ip, _dp]
        L.s2go_cic_mapping_parallel.restype = C.c_int
        L.s2go_cic_mapping_parallel.argtypes = [C.c_int] + [_dp] * 6 + [C.c_int64, C.c_int, C.c_double, C.c_int64,
                                                                       C.c_int, C.c_int, C.c_int, C.c_int, _dp, _ip]
        L.s2go_max_threads.restype = C.c_int
        L.s2go_reduce_image_2d.argtypes = [_dp, C.c_int64, C.c_int64, C.c_int, C.c_int, _dp]
        L.s2go_reduce_image_3d.argtypes = [_dp, C.c_int64, C.c_int, _dp]
        L.s2go_accumulate_finite.argtypes = [_dp, _dp, C.c_int64]
        L.s2go_hp_ang2pix_ring.restype = C.c_int64
        L.s2go_hp_ang2pix_ring.argtypes = [C.c_int64, C.c_double, C.c_double]
        L.s2go_hp_pix2ang_ring.argtypes = [C.c_int64, C.c_int64, _dp, _dp]
        L.s2go_hp_pix2vec_ring.argtypes = [C.c_int64, C.c_int64, _dp]
        L.s2go_hp_vec2ang.argtypes = [C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.s2go_hp_query_disc_ring.restype = C.c_int64
        L.s2go_hp_query_disc_ring.argtypes = [C.c_int64, C.c_double, C.c_double, C.c_double, _ip, C.c_int64]
        L.s2go_hp_contributing_pixels.restype = C.c_int64
        L.s2go_hp_contributing_pixels.argtypes = [C.c_int64, _dp, C.c_double, _ip, C.c_int64]
        L.s2go_healpix_deposit.restype = C.c_int
        L.s2go_healpix_deposit.argtypes = [_dp] * 6 + [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, _dp, _dp, _ip]
        L.s2go_healpix_deposit_parallel.restype = C.c_int
        L.s2go_healpix_deposit_parallel.argtypes = [_dp] * 6 + [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                                                C.c_int, _dp, _dp]
        L.s2go_stencil_deposit.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int64, C.c_double, C.c_int64, C.c_int, _dp]
        L.s2go_stencil_average.argtypes = [_dp, C.c_int64, _dp]
        L.s2go_healpix_deposit_exact.restype = C.c_int
        L.s2go_healpix_deposit_exact.argtypes = [_dp] * 6 + [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, _dp, _dp,
                                                             _ip, _dp]
        L.s2go_healpix_deposit_chord64.restype = C.c_int
        L.s2go_healpix_deposit_chord64.argtypes = [_dp] * 6 + [C.c_int64, C.c_int64, C.c_int, C.c_int, _dp, _dp]
        L.s2go_hp_angdist_exact.argtypes = [C.c_int64, C.c_int64, _dp, _dp, C.c_void_p]
        _libs[_which] = L
    return _libs[_which]


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pos3xn(pos):
    """Julia Matrix(3,N) column-major memory == numpy (N,3) C-contiguous.  Accepts (N,3) C-order, or a
    Fortran-ordered (3,N) array (whose transpose is a C-contiguous (N,3) VIEW, so in-place recentring works)."""
    pos = np.asarray(pos)
    if pos.ndim != 2:
        raise ValueError("pos must be 2-D")
    if pos.shape == (3, 3):
        return pos.T if (pos.flags.f_contiguous and not pos.flags.c_contiguous) else pos
    if pos.shape[0] == 3:
        return pos.T
    return pos


# --------------------------------------------------------------------------- parameters
@dataclass
class MappingParameters:
    x_lim: np.ndarray
    y_lim: np.ndarray
    z_lim: np.ndarray
    center: np.ndarray
    halfsize: np.ndarray
    len2pix: float
    pixelSideLength: float
    Npixels: np.ndarray
    boxsize: float
    periodic: bool


class ReferenceError_(Exception):
    pass


def mapping_parameters(x_lim=(-1.0, -1.0), y_lim=(-1.0, -1.0), z_lim=(-1.0, -1.0), center=(-1.0, -1.0, -1.0),
                       x_size=-1.0, y_size=-1.0, z_size=-1.0, pixelSideLength=-1.0, Npixels=0, boxsize=-1.0):
    par = np.zeros(14)
    npix = C.c_int64(0)
    xl, yl, zl, cc = _c64(x_lim), _c64(y_lim), _c64(z_lim), _c64(center)
    rc = lib().s2go_mapping_parameters(_d(xl), _d(yl), _d(zl), _d(cc), float(x_size), float(y_size), float(z_size),
                                       float(pixelSideLength), int(Npixels), _d(par), C.byref(npix))
    if rc == 1:
        raise ReferenceError_("Giving a center position requires extent in x, y and z direction.")
    if rc == 2:
        raise ReferenceError_("Please specify pixelSideLength or number of pixels!")
    return MappingParameters(par[0:2].copy(), par[2:4].copy(), par[4:6].copy(), par[6:9].copy(), par[9:12].copy(),
                             float(par[12]), float(par[13]), np.array([npix.value] * 3, dtype=np.int64),
                             float(boxsize), boxsize != -1.0)


def center_particles(pos, par: MappingParameters):
    """filter_shift.jl:6-32.  pos: (N,3) C-contiguous f32/f64, modified IN PLACE."""
    cen = _c64(par.center)
    n = pos.shape[0]
    if pos.dtype == np.float32:
        lib().s2go_center_particles_f32(pos.ctypes.data_as(_fp), n, _d(cen), int(par.periodic), float(par.boxsize))
    else:
        lib().s2go_center_particles_f64(_d(pos), n, _d(cen), int(par.periodic), float(par.boxsize))
    par2 = mapping_parameters(center=[0.0, 0.0, 0.0], x_lim=par.x_lim - cen[0], y_lim=par.y_lim - cen[1],
                              z_lim=par.z_lim - cen[2], Npixels=int(par.Npixels.max()), boxsize=par.boxsize)
    return pos, par2


def filter_particles_in_image(pos, par: MappingParameters, sort_z=False):
    """filter_shift.jl:40-67 (incl. the sorted[mask] quirk Q5 when sort_z)."""
    n = pos.shape[0]
    mask = np.zeros(n, dtype=np.uint8)
    cen, hs = _c64(par.center), _c64(par.halfsize)
    if pos.dtype == np.float32:
        lib().s2go_filter_particles_f32(pos.ctypes.data_as(_fp), n, _d(cen), _d(hs), mask.ctypes.data_as(_bp))
    else:
        lib().s2go_filter_particles_f64(_d(pos), n, _d(cen), _d(hs), mask.ctypes.data_as(_bp))
    mask = mask.astype(bool)
    if sort_z:
        srt = np.argsort(pos[:, 2], kind="stable")[::-1]
        return srt[mask]
    return mask


def domain_decomposition(n, n_workers):
    s = np.zeros(n_workers, dtype=np.int64)
    e = np.zeros(n_workers, dtype=np.int64)
    lib().s2go_domain_decomposition(n, n_workers, _i(s), _i(e))
    return list(zip(s.tolist(), e.tolist()))


# --------------------------------------------------------------------------- deposits
def cic_mapping_2d(pos, hsml, m, rho, binq, w, len2pix, npix, kernel="WendlandC6", kernel_dim=2, calc_mean=True,
                   want_footprints=False, n_workers=0):
    """cic_2D.jl:103-244.  pos (N,3); binq (N,) or (N,I) [= Julia (I,N)].  Returns flat image (npix*npix, I+1)
    as a Fortran-ordered array (plane-separated memory), weight plane last."""
    pos = _c64(pos); hsml = _c64(hsml); m = _c64(m); rho = _c64(rho); w = _c64(w)
    binq = _c64(binq)
    n = hsml.shape[0]
    n_images = 1 if binq.ndim == 1 else binq.shape[1]
    image = np.zeros((npix * npix, n_images + 1), order="F")
    fp = np.zeros((n, 4), dtype=np.int64) if want_footprints else None
    st = np.zeros(4, dtype=np.int64)
    kid = KERNEL_IDS[kernel]
    if n_workers and n_workers > 0:
        rc = lib().s2go_cic_mapping_parallel(2, _d(pos), _d(hsml), _d(m), _d(rho), _d(binq), _d(w), n, n_images,
                                             float(len2pix), int(npix), kid, kernel_dim, int(calc_mean),
                                             int(n_workers), _d(image), _i(st))
    else:
        rc = lib().s2go_cic_mapping_2d(_d(pos), _d(hsml), _d(m), _d(rho), _d(binq), _d(w), n, n_images,
                                       float(len2pix), int(npix), kid, kernel_dim, int(calc_mean), _d(image), _i(fp),
                                       _i(st))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    stats = dict(n_mapped=int(st[0]), footprint_pixels=int(st[1]), touched_pixels=int(st[2]), n_fallback=int(st[3]))
    return (image, fp, stats) if want_footprints else (image, stats)


def cic_mapping_2d_rm(pos, hsml, m, rho, binq, w, rm, len2pix, npix, kernel="WendlandC6", kernel_dim=2,
                      calc_mean=True, stokes=True):
    """cic_mapping_2D with the RM argument (cic_2D.jl:103-244 + faraday_rotate_pixel! cic_shared.jl:129-159):
    particles are composited in the given order; planes 1/2 are Stokes Q/U."""
    pos = _c64(pos); hsml = _c64(hsml); m = _c64(m); rho = _c64(rho); w = _c64(w); rm = _c64(rm)
    binq = _c64(binq)
    n = hsml.shape[0]
    n_images = 1 if binq.ndim == 1 else binq.shape[1]
    image = np.zeros((npix * npix, n_images + 1), order="F")
    st = np.zeros(4, dtype=np.int64)
    rc = lib().s2go_cic_mapping_2d_rm(_d(pos), _d(hsml), _d(m), _d(rho), _d(binq), _d(w), _d(rm), n, n_images,
                                      float(len2pix), int(npix), KERNEL_IDS[kernel], kernel_dim, int(calc_mean),
                                      int(stokes), _d(image), _i(st))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    stats = dict(n_mapped=int(st[0]), footprint_pixels=int(st[1]), touched_pixels=int(st[2]), n_fallback=int(st[3]))
    return image, stats


def cic_mapping_3d(pos, hsml, m, rho, binq, w, len2pix, npix, kernel="Cubic", kernel_dim=3, calc_mean=False,
                   want_footprints=False, n_workers=0):
    pos = _c64(pos); hsml = _c64(hsml); m = _c64(m); rho = _c64(rho); w = _c64(w); binq = _c64(binq)
    n = hsml.shape[0]
    image = np.zeros((npix ** 3, 2), order="F")
    fp = np.zeros((n, 6), dtype=np.int64) if want_footprints else None
    st = np.zeros(4, dtype=np.int64)
    mass2 = np.zeros(2)
    kid = KERNEL_IDS[kernel]
    if n_workers and n_workers > 0:
        rc = lib().s2go_cic_mapping_parallel(3, _d(pos), _d(hsml), _d(m), _d(rho), _d(binq), _d(w), n, 1,
                                             float(len2pix), int(npix), kid, kernel_dim, int(calc_mean),
                                             int(n_workers), _d(image), _i(st))
    else:
        rc = lib().s2go_cic_mapping_3d(_d(pos), _d(hsml), _d(m), _d(rho), _d(binq), _d(w), n, float(len2pix),
                                       int(npix), kid, kernel_dim, int(calc_mean), _d(image), _i(fp), _i(st),
                                       _d(mass2))
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    stats = dict(n_mapped=int(st[0]), footprint_pixels=int(st[1]), touched_pixels=int(st[2]), n_fallback=int(st[3]),
                 grid_mass=float(mass2[0]), particle_mass=float(mass2[1]))
    return (image, fp, stats) if want_footprints else (image, stats)


def reduce_image_2d(image, nx, ny, reduce_image=True):
    """reduce_image.jl:8-31.  Returns array indexed [ix, iy, n] with Julia's memory layout (Fortran order)."""
    image = np.asfortranarray(image, dtype=np.float64)
    n_images = image.shape[1] - 1
    out = np.zeros((ny, nx, n_images), order="F")
    lib().s2go_reduce_image_2d(_d(image), nx, ny, n_images, int(reduce_image), _d(out))
    return out


def reduce_image_3d(image, npix, reduce_image=True):
    """reduce_image.jl:39-55 after cic_interpolation.jl:230-232.  Returns array indexed [iz, iy, ix] (Fortran order)."""
    image = np.asfortranarray(image, dtype=np.float64)
    out = np.zeros((npix, npix, npix), order="F")
    lib().s2go_reduce_image_3d(_d(image), npix, int(reduce_image), _d(out))
    return out


def sph_mapping(pos, hsml, m, rho, binq, weights=None, *, param: MappingParameters, kernel="WendlandC6",
                kernel_dim=None, parallel=False, n_workers=0, reduce_image=True, return_both_maps=False,
                dimensions=2, calc_mean=False, sort_z=False, stokes=False, rm=None):
    """cic_interpolation.jl:35-273.  pos is (N,3) [Julia (3,N)] and IS MUTATED (Q1)."""
    if weights is None:
        weights = rho
    if stokes:  # cic_interpolation.jl:74-81: back-to-front order, serial
        sort_z = True
        parallel = False
    pos = _pos3xn(pos)
    if not pos.flags.c_contiguous or pos.dtype not in (np.float32, np.float64):
        raise ValueError("pos must be C-contiguous (N,3) float32/float64 (it is recentred in place)")
    pos, par = center_particles(pos, param)
    sel = filter_particles_in_image(pos, par, sort_z)
    x = pos[sel]
    hs = np.asarray(hsml)[sel]; mm = np.asarray(m)[sel]; rr = np.asarray(rho)[sel]
    bq = np.asarray(binq)
    bq = bq[sel] if bq.ndim == 1 else bq[sel, :]
    ww = np.asarray(weights)[sel]
    npix = int(par.Npixels[0])
    nw = n_workers if parallel else 0
    if dimensions == 2:
        kd = 2 if kernel_dim is None else kernel_dim
        if rm is not None:
            if parallel:
                raise ValueError("the reference passes RM to the serial deposit only (cic_interpolation.jl:152)")
            # NB: sphMapping does NOT forward `stokes` to cic_mapping_2D (cic_interpolation.jl:152-155), so the
            # rotation branch of faraday_rotate_pixel! is dead through the public API; only a direct
            # cic_mapping_2D(...; stokes=true) call rotates.  Mirrored literally.
            image, _ = cic_mapping_2d_rm(x, hs, mm, rr, bq, ww, np.asarray(rm)[sel], par.len2pix, npix, kernel, kd,
                                         calc_mean, False)
        else:
            image, _ = cic_mapping_2d(x, hs, mm, rr, bq, ww, par.len2pix, npix, kernel, kd, calc_mean, n_workers=nw)
        if return_both_maps:
            return image
        return reduce_image_2d(image, int(param.Npixels[0]), int(param.Npixels[1]), reduce_image)
    elif dimensions == 3:
        kd = 3 if kernel_dim is None else kernel_dim
        # NB: the reference does not forward calc_mean to cic_mapping_3D (cic_interpolation.jl:219-221)
        image, _ = cic_mapping_3d(x, hs, mm, rr, bq, ww, par.len2pix, npix, kernel, kd, False, n_workers=nw)
        return reduce_image_3d(image, npix, reduce_image)
    raise ValueError("dimensions must be 2 or 3")


# --------------------------------------------------------------------------- projections (map_it pre-step)
def rotate_to_xz_plane(pos):
    """rotate_to_xz_plane! (rotate_particles.jl:35-43): swaps y and z of every particle, in place.  pos (N,3)."""
    for i in range(pos.shape[0]):
        pos3 = pos[i, 1].copy()
        pos[i, 1] = pos[i, 2]
        pos[i, 2] = pos3
    return pos


def rotate_to_yz_plane(pos):
    """rotate_to_yz_plane! (rotate_particles.jl:65-74): (x,y,z) <- (y,z,x), in place."""
    for i in range(pos.shape[0]):
        pos3 = pos[i, 0].copy()
        pos[i, 0] = pos[i, 1]
        pos[i, 1] = pos[i, 2]
        pos[i, 2] = pos3
    return pos


def rotate_parameters_xz(par: MappingParameters):
    """rotate_to_xz_plane(par) (rotate_parameters.jl:27-40)"""
    return mapping_parameters(center=[par.center[0], par.center[2], par.center[1]], x_lim=par.x_lim.copy(),
                              y_lim=par.z_lim.copy(), z_lim=par.y_lim.copy(), Npixels=int(par.Npixels.max()),
                              boxsize=par.boxsize)


def rotate_parameters_yz(par: MappingParameters):
    """rotate_to_yz_plane(par) (rotate_parameters.jl:48-59)"""
    return mapping_parameters(center=[par.center[1], par.center[2], par.center[0]], x_lim=par.y_lim.copy(),
                              y_lim=par.z_lim.copy(), z_lim=par.x_lim.copy(), Npixels=int(par.Npixels.max()),
                              boxsize=par.boxsize)


def rotate_3d(pos, alpha, beta, gamma):
    """rotate_3D (rotate_particles.jl:7-13): RotXYZ(deg2rad.(angles)) * x, with RotXYZ = Rx*Ry*Rz built here from the
    three elementary matrices (Rotations.jl itself is not in the reference tree: last-ulp parity unpinned)."""
    import math
    a, b, g = math.radians(alpha), math.radians(beta), math.radians(gamma)
    rx = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
    ry = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
    rz = np.array([[math.cos(g), -math.sin(g), 0], [math.sin(g), math.cos(g), 0], [0, 0, 1]])
    rot = rx @ ry @ rz
    return (rot @ np.asarray(pos, dtype=np.float64).T).T.copy()


def map_it(pos_in, hsml, m, rho, binq, weights, *, param: MappingParameters, kernel="WendlandC6", reduce_image=True,
           calc_mean=True, projection="xy", **kw):
    """map_it without the FITS output (cic_interpolation.jl:312-359)."""
    pos = np.array(pos_in, copy=True)
    if projection == "xy":
        par = param
    elif projection == "xz":
        pos = rotate_to_xz_plane(pos); par = rotate_parameters_xz(param)
    elif projection == "yz":
        pos = rotate_to_yz_plane(pos); par = rotate_parameters_yz(param)
    else:
        pos = rotate_3d(pos, *projection); par = param  # (`par` is unassigned in the reference: intent restated)
    return sph_mapping(np.ascontiguousarray(pos), hsml, m, rho, binq, weights, param=par, kernel=kernel,
                       reduce_image=reduce_image, calc_mean=calc_mean, **kw)


# --------------------------------------------------------------------------- HEALPix
def healpix_deposit(pos, hsml, m, rho, binq, w, nside, kernel="WendlandC4", kernel_dim=2, calc_mean=True, n_workers=0,
                    exact=False):
    """main.jl:143-213.  `exact=True`: the extended-precision arbiter (s2g_oracle_exact.c) — same pixel lists, weights
    in long double; `n_workers` threads share the maps through atomics there.  `exact="sens"` additionally returns
    stats["sens"] / stats["sens_q"], the per-pixel sensitivities Σ|∂pix_weight/∂dx| of the weight map and Σ|q ∂pix_weight/∂dx|
    of the quantity map (see s2go_healpix_deposit_exact)."""
    pos = _c64(pos); hsml = _c64(hsml); m = _c64(m); rho = _c64(rho); w = _c64(w); binq = _c64(binq)
    n = hsml.shape[0]
    npix = 12 * nside * nside
    amap = np.zeros(npix); wmap = np.zeros(npix)
    st = np.zeros(4, dtype=np.int64)
    kid = KERNEL_IDS[kernel]
    if exact:
        sens = np.zeros(2 * npix) if exact == "sens" else None
        lib().s2go_healpix_deposit_exact(_d(pos), _d(hsml), _d(m), _d(rho), _d(binq), _d(w), n, nside, kid,
                                         int(calc_mean), int(n_workers or 1), _d(amap), _d(wmap), _i(st),
                                         _d(sens) if sens is not None else None)
        if sens is not None:
            stats = dict(n_mapped=int(st[0]), footprint_pixels=int(st[1]), touched_pixels=int(st[2]),
                         n_fallback=int(st[3]), sens=sens[:npix], sens_q=sens[npix:])
            return amap, wmap, stats
    elif n_workers and n_workers > 0:
        lib().s2go_healpix_deposit_parallel(_d(pos), _d(hsml), _d(m), _d(rho), _d(binq), _d(w), n, nside, kid,
                                            kernel_dim, int(calc_mean), int(n_workers), _d(amap), _d(wmap))
    else:
        lib().s2go_healpix_deposit(_d(pos), _d(hsml), _d(m), _d(rho), _d(binq), _d(w), n, nside, kid, kernel_dim,
                                   int(calc_mean), _d(amap), _d(wmap), _i(st))
    stats = dict(n_mapped=int(st[0]), footprint_pixels=int(st[1]), touched_pixels=int(st[2]), n_fallback=int(st[3]))
    return amap, wmap, stats


def filter_sort_particles(pos, hsml, m, rho, binq, weights, center, radius_limits, calc_mean):
    """filter_particles.jl:17-54 incl. Q1 (in-place recentre), Q5 (sorted[mask]) and Q11 (BoundsError)."""
    pos -= np.asarray(center, dtype=pos.dtype)[None, :]
    dx = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2 + pos[:, 2] ** 2)
    sel = (radius_limits[0] <= dx) & (dx <= radius_limits[1])
    if not calc_mean:
        sel = sel[np.asarray(binq)[sel] > 0.0]
    srt = np.argsort(dx, kind="stable")[::-1]
    if sel.shape[0] != srt.shape[0]:
        raise IndexError("BoundsError: attempt to access %d-element Vector{Int64} at index [%d-element BitVector]"
                         % (srt.shape[0], sel.shape[0]))
    idx = srt[sel]
    return (pos[idx], np.asarray(hsml)[idx], np.asarray(m)[idx], np.asarray(rho)[idx], np.asarray(binq)[idx],
            np.asarray(weights)[idx])


def healpix_map(pos, hsml, m, rho, binq, weights, *, center=(0.0, 0.0, 0.0), radius_limits=(0.0, np.inf), nside=1024,
                kernel="WendlandC4", kernel_dim=2, calc_mean=True, n_workers=0, exact=False):
    """main.jl:92-227.  pos (N,3) f64, mutated in place (Q1).  Returns (map, weight_map), un-reduced."""
    pos = _pos3xn(pos)
    npix = 12 * nside * nside
    if (not calc_mean) and np.sum(binq) == 0:
        return np.zeros(npix), np.zeros(npix)
    p, h, mm, rr, bq, ww = filter_sort_particles(pos, hsml, m, rho, binq, weights, center, radius_limits, calc_mean)
    a, wm, st = healpix_deposit(p, h, mm, rr, bq, ww, nside, kernel, kernel_dim, calc_mean, n_workers, exact)
    return (a, wm, st) if exact == "sens" else (a, wm)


# --------------------------------------------------------------------------- stencils
def stencil_deposit(order, dims, pos, q, len2pix, npix, periodic=False):
    pos = _c64(pos); q = _c64(q)
    ncell = npix ** dims
    image = np.zeros((ncell, 2), order="F")
    lib().s2go_stencil_deposit(order, dims, _d(pos), _d(q), q.shape[0], float(len2pix), npix, int(periodic), _d(image))
    return image


def stencil_average(image):
    image = np.asfortranarray(image)
    out = np.zeros(image.shape[0])
    lib().s2go_stencil_average(_d(image), image.shape[0], _d(out))
    return out


def kernel_shape(kernel, u):
    return lib().s2go_kernel_shape(KERNEL_IDS[kernel], float(u))


def kernel_value(kernel, dim, u, h_inv):
    return lib().s2go_kernel_value(KERNEL_IDS[kernel], dim, float(u), float(h_inv))


def hp_angdist_exact(nside, pix, pos):
    """(long-double chord form, __float128 literal acos(min(d/r,1)), long-double acos form) of the angle between the
    particle at `pos` and the centre of RING pixel `pix`; plus the raw long doubles (np.longdouble[3])."""
    out = np.zeros(3)
    raw = np.zeros(3, dtype=np.longdouble)
    lib().s2go_hp_angdist_exact(int(nside), int(pix), _d(_c64(pos)), _d(out), raw.ctypes.data)
    return out, raw


def healpix_deposit_chord64(pos, hsml, m, rho, binq, w, nside, kernel="WendlandC4", calc_mean=True):
    """Float64 chord formulation (what the CUDA kernels evaluate), on the CPU — conditioning study only."""
    pos = _c64(pos); hsml = _c64(hsml); m = _c64(m); rho = _c64(rho); w = _c64(w); binq = _c64(binq)
    npix = 12 * nside * nside
    amap = np.zeros(npix); wmap = np.zeros(npix)
    lib().s2go_healpix_deposit_chord64(_d(pos), _d(hsml), _d(m), _d(rho), _d(binq), _d(w), hsml.shape[0], nside,
                                       KERNEL_IDS[kernel], int(calc_mean), _d(amap), _d(wmap))
    return amap, wmap
