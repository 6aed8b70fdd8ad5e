"""Faraday-rotation / Stokes compositing branch of cic_mapping_2D (cic_2D.jl:201-217, cic_shared.jl:129-159).

CPU part: the C oracle against a hand-derived known answer and against the independent numpy mirror.
GPU part: the ordered compositing kernels (s2g_stokes2d.cu, through s2g_deposit_2d_rm) against the oracle.

Tolerance.  Q and U of one pixel are the two components of ONE rotated vector; a rotation error of eps radians moves
each component by eps*Ipol, however small that component is.  The natural per-pixel bar is therefore
|d(Q,U)| <= 1e-10 * Ipol (+ the usual 1e-14-of-plane-maximum floor), the weight plane keeps the element-wise bar."""
import math
import os

import numpy as np
import pytest

from util import assert_parity, kern, random_particles


def polarised_parity(got, ref, nim, rtol=1e-10, what=""):
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref)), what + ": NaN pattern differs"
    g = np.nan_to_num(got); r = np.nan_to_num(ref)
    ipol = np.hypot(r[:, 0], r[:, 1])
    d = np.hypot(g[:, 0] - r[:, 0], g[:, 1] - r[:, 1])
    floor = 1e-14 * float(ipol.max()) if ipol.size else 0.0
    e = float(np.max(d / np.maximum(ipol, max(floor, 1e-300)))) if ipol.size else 0.0
    assert e <= rtol, f"{what}: (Q,U) differs by {e:.3e} of Ipol"
    for k in range(2 if nim >= 2 else 1, nim + 1):  # remaining planes incl. weights: element-wise
        assert_parity(got[:, k], ref[:, k], rtol, what + f" plane {k}")
    return e


def stokes_particles(seed, n, box=10.0, hmax=1.2, dtype=np.float64, nim=2):
    pos, hsml, m, rho, _, w = random_particles(seed, n, box=box, hmax=hmax, dtype=dtype)
    rng = np.random.default_rng(seed + 1000)
    q = rng.normal(size=(n, nim)).astype(dtype) if nim > 1 else rng.normal(size=n).astype(dtype)
    rm = rng.normal(size=n) * 0.5  # RM*pix_weight = O(1) rad: well conditioned under mod(., pi)
    return pos, hsml, m, rho, q, w, rm


# ----------------------------------------------------------------------------------------------- CPU (oracle)
def test_oracle_rotation_known_answer(oracle):
    """Two particles on the same single pixel (sub-pixel hsml -> 'no pixel centre covered' branch, weight = the whole
    particle): the second rotates what the first left by mod(RM2*pw2, pi), then adds its own (Q,U)."""
    npix, len2pix = 4, 1.0
    pos = np.array([[0.3, 0.3, 0.0], [0.4, 0.35, 0.0]])  # pixel (2,2) of a 4x4 map centred on 0
    hsml = np.array([0.05, 0.05]); m = np.array([2.0, 3.0]); rho = np.array([1.0, 1.0]); w = np.array([1.0, 1.0])
    q = np.array([[1.0, 2.0], [-0.5, 0.25]]); rm = np.array([123.0, 0.37])
    img, st = oracle.cic_mapping_2d_rm(pos, hsml, m, rho, q, w, rm, len2pix, npix, "Cubic", stokes=True)
    assert st["n_fallback"] == 2
    plain, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, len2pix, npix, "Cubic")
    idx = 2 * npix + 2
    pw = [None, None]
    for p in range(2):  # by hand: h = 0.05 px, footprint = one pixel, A = (2h)^2, wpp = 1/A, area_norm = area*wpp*dz
        h = hsml[p] * len2pix
        area = (2 * h) ** 2
        dz = m[p] / rho[p] / area
        pw[p] = 1.0 * area * (area / 1 * (1 / area) * w[p] * dz)
    assert plain[idx, 2] == pytest.approx(pw[0] + pw[1], rel=1e-14)
    Q1, U1 = q[0, 0] * pw[0], q[0, 1] * pw[0]
    theta = math.fmod(rm[1] * pw[1], math.pi)
    ipol = math.hypot(Q1, U1)
    psi = 0.5 * math.atan(U1 / Q1)
    Q2 = ipol * math.cos(2 * (psi + theta)) + q[1, 0] * pw[1]
    U2 = ipol * math.sin(2 * (psi + theta)) + q[1, 1] * pw[1]
    assert img[idx, 0] == pytest.approx(Q2, rel=1e-13)
    assert img[idx, 1] == pytest.approx(U2, rel=1e-13)
    assert img[idx, 2] == plain[idx, 2]
    assert np.count_nonzero(img) == 3
    # RM without stokes is inert (faraday_rotate_pixel! only computes the angle)
    inert, _ = oracle.cic_mapping_2d_rm(pos, hsml, m, rho, q, w, rm, len2pix, npix, "Cubic", stokes=False)
    assert np.array_equal(inert, plain)


def test_oracle_rotation_quirks(oracle):
    """atan(U/Q) drops the quadrant: a touched pixel with Q < 0 comes back sign-flipped even for RM = 0; a touched
    pixel with Q = U = 0 turns NaN."""
    npix, len2pix = 4, 1.0
    pos = np.array([[0.3, 0.3, 0.0], [0.4, 0.35, 0.0]])
    hsml = np.array([0.05, 0.05]); m = np.ones(2); rho = np.ones(2); w = np.ones(2)
    q = np.array([[-1.0, 2.0], [0.0, 0.0]]); rm = np.zeros(2)
    img, _ = oracle.cic_mapping_2d_rm(pos, hsml, m, rho, q, w, rm, len2pix, npix, "Cubic", calc_mean=True, stokes=True)
    idx = 2 * npix + 2
    assert img[idx, 0] == pytest.approx(+1.0, rel=1e-14) and img[idx, 1] == pytest.approx(-2.0, rel=1e-14)
    q0 = np.array([[0.0, 0.0], [1.0, 1.0]])
    img, _ = oracle.cic_mapping_2d_rm(pos, hsml, m, rho, q0, w, rm, len2pix, npix, "Cubic", calc_mean=True, stokes=True)
    assert np.isnan(img[idx, 0]) and np.isnan(img[idx, 1]) and img[idx, 2] == pytest.approx(2.0, rel=1e-14)


@pytest.mark.parametrize("kernel", ["Cubic", "WendlandC4"])
def test_oracle_rotation_vs_numpy_mirror(oracle, kernel):
    from oracle import numpy_mirror as M
    pos, hsml, m, rho, q, w, rm = stokes_particles(5, 250, box=10.0, hmax=1.5)
    npix = 24
    len2pix = npix / 10.0
    a, _ = oracle.cic_mapping_2d_rm(pos, hsml, m, rho, q, w, rm, len2pix, npix, kernel, stokes=True)
    b = M.cic_mapping_2d(pos, hsml, m, rho, q, w, len2pix, npix, kernel, rm=rm, stokes=True)
    polarised_parity(a, b, 2, 1e-12, "oracle vs mirror")


def test_oracle_sphmapping_does_not_forward_stokes(oracle):
    """cic_interpolation.jl:152-155 omits `stokes` in the cic_mapping_2D call: through sphMapping the RM is inert and
    stokes=true only enforces the far->near order."""
    pos, hsml, m, rho, q, w, rm = stokes_particles(6, 400, box=10.0)
    par = oracle.mapping_parameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=32)
    a = oracle.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=par, kernel="WendlandC4", reduce_image=False,
                           calc_mean=True, stokes=True, rm=rm)
    b = oracle.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=par, kernel="WendlandC4", reduce_image=False,
                           calc_mean=True, sort_z=True)
    assert np.array_equal(a, b)


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["Cubic", "Quintic", "WendlandC2", "WendlandC4", "WendlandC6", "WendlandC8"])
def test_gpu_stokes_parity(s2g, oracle, kernel):
    pos, hsml, m, rho, q, w, rm = stokes_particles(11, 5000, box=11.0, hmax=1.4)
    hsml[:200] *= 0.01  # "no pixel centre covered" branch
    q[50:60] = 0.0      # bin_q collapses to the scalar 0.0 (calc_mean=True keeps the particle)
    npix = 150
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    got, st = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, rm, param=par, kernel=kern(s2g, kernel), calc_mean=True,
                                 stokes=True, return_stats=True)
    ref, ost = oracle.cic_mapping_2d_rm(pos, hsml, m, rho, q, w, rm, par.len2pix, npix, kernel, stokes=True)
    polarised_parity(got, ref, 2, 1e-10, f"stokes {kernel}")
    assert st["n_mapped"] == ost["n_mapped"] and st["footprint_pixels"] == ost["footprint_pixels"]
    assert st["touched_pixels"] == ost["touched_pixels"] and st["n_fallback"] == ost["n_fallback"]
    # the rotation really happened (the plain deposit differs) and the weight plane is untouched by it
    plain, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, par.len2pix, npix, kernel)
    assert np.nanmax(np.abs(got[:, 0] - plain[:, 0])) > 1e-3 * np.abs(plain[:, 0]).max()
    assert_parity(got[:, 2], plain[:, 2], 1e-10, "weights")


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nim", [1, 2, 4])
def test_gpu_stokes_planes_and_dtypes(s2g, oracle, dtype, nim):
    """n_images = 1: plane 2 IS the weight plane (the reference does not check); > 2: further planes just accumulate."""
    pos, hsml, m, rho, q, w, rm = stokes_particles(12, 3000, box=11.0, hmax=1.2, dtype=dtype, nim=nim)
    npix = 96
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    for calc_mean in (True, False):
        got = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, rm, param=par, kernel=s2g.WendlandC4(2),
                                 calc_mean=calc_mean, stokes=True)
        ref, _ = oracle.cic_mapping_2d_rm(pos, hsml, m, rho, q, w, rm, par.len2pix, npix, "WendlandC4",
                                          calc_mean=calc_mean, stokes=True)
        polarised_parity(got, ref, nim, 1e-10, f"nim={nim} {dtype.__name__} calc_mean={calc_mean}")


@pytest.mark.gpu
def test_gpu_stokes_order_dependence_and_slicing(s2g, oracle):
    pos, hsml, m, rho, q, w, rm = stokes_particles(13, 4000, box=11.0, hmax=1.0)
    npix = 128
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    k = s2g.WendlandC6(2)
    a = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, rm, param=par, kernel=k, stokes=True)
    r = slice(None, None, -1)
    b = s2g.cic_mapping_2D(pos[r].copy(), hsml[r].copy(), m[r].copy(), rho[r].copy(), q[r].copy(), w[r].copy(),
                           rm[r].copy(), param=par, kernel=k, stokes=True)
    refb, _ = oracle.cic_mapping_2d_rm(pos[r], hsml[r], m[r], rho[r], q[r], w[r], rm[r], par.len2pix, npix,
                                       "WendlandC6", stokes=True)
    polarised_parity(b, refb, 2, 1e-10, "reversed order")
    assert np.nanmax(np.abs(a[:, 0] - b[:, 0])) > 1e-3 * np.nanmax(np.abs(a[:, 0]))  # compositing is order dependent
    # consecutive slices carry the pixel state (Q, U, weight, touched) through the image: bit-identical result
    old = os.environ.get("S2G_STOKES_BATCH")
    os.environ["S2G_STOKES_BATCH"] = "1024"
    try:
        c = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, rm, param=par, kernel=k, stokes=True)
    finally:
        if old is None:
            del os.environ["S2G_STOKES_BATCH"]
        else:
            os.environ["S2G_STOKES_BATCH"] = old
    assert np.array_equal(a, c)


@pytest.mark.gpu
def test_gpu_stokes_quirks(s2g, oracle):
    """sign flip for Q < 0 under RM = 0, NaN for a touched pixel with Q = U = 0 — the same as the reference."""
    npix = 4
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=4.0, y_size=4.0, z_size=4.0, Npixels=npix)
    pos = np.array([[0.3, 0.3, 0.0], [0.4, 0.35, 0.0]])
    hsml = np.array([0.05, 0.05]); m = np.ones(2); rho = np.ones(2); w = np.ones(2); rm = np.zeros(2)
    for q in (np.array([[-1.0, 2.0], [0.0, 0.0]]), np.array([[0.0, 0.0], [1.0, 1.0]])):
        got = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, rm, param=par, kernel=s2g.Cubic(2), stokes=True)
        ref, _ = oracle.cic_mapping_2d_rm(pos, hsml, m, rho, q, w, rm, par.len2pix, npix, "Cubic", stokes=True)
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert np.allclose(np.nan_to_num(got), np.nan_to_num(ref), rtol=1e-13, atol=0)


@pytest.mark.gpu
def test_gpu_rm_without_stokes_is_inert_and_sphmapping_quirk(s2g, oracle):
    pos, hsml, m, rho, q, w, rm = stokes_particles(14, 3000, box=11.0)
    npix = 64
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    k = s2g.WendlandC4(2)
    a = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, rm, param=par, kernel=k, stokes=False)
    b = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=k)
    assert_parity(a, b, 1e-10, "RM without stokes")
    with pytest.raises(TypeError):
        s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, rm.astype(np.float32), param=par, kernel=k, stokes=True)
    # through sphMapping `stokes` only sorts (the reference never forwards it)
    opar = oracle.mapping_parameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    ref = oracle.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=opar, kernel="WendlandC4", reduce_image=False,
                             calc_mean=True, stokes=True, rm=rm)
    got = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, rm, param=par, kernel=k, reduce_image=False, calc_mean=True,
                         stokes=True, show_progress=False)
    assert_parity(got, ref, 1e-10, "sphMapping(stokes=true)")
