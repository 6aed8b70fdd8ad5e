"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for footprints / index work, 1e-10 relative for FP64 maps (tolerance of BASELINE.json's north_star)."""
import math

import numpy as np
import pytest

from util import KERNELS, assert_parity, kern, random_particles

pytestmark = pytest.mark.gpu


def _oracle_2d(oracle, pos, hsml, m, rho, q, w, len2pix, npix, kernel, calc_mean):
    return oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, len2pix, npix, kernel, 2, calc_mean, want_footprints=True)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dims", [2, 3])
def test_footprints_bit_exact(s2g, oracle, dtype, dims):
    import ctypes as C
    from sphtogrid_b200 import _lib
    pos, hsml, m, rho, q, w = random_particles(1, 50000, box=12.0, hmax=3.0, dtype=dtype)
    hsml[:100] = 40.0          # much larger than the image
    hsml[100:200] = 1e-6       # sub-pixel
    pos[200:210] = 0.0         # exactly on pixel edges / centre
    pos[210:220, 0] = 5.0      # exactly on the image border
    npix = 96 if dims == 2 else 40
    len2pix = npix / 10.0
    ctx = s2g.default_context()
    out = np.zeros((pos.shape[0], 2 * dims), dtype=np.int64)
    _lib.check(s2g.lib().s2g_footprints(ctx.handle, _lib.ptr(pos), _lib.ptr(hsml), pos.shape[0],
                                        0 if dtype == np.float32 else 1, len2pix, npix, dims, _lib.ptr(out)))
    if dims == 2:
        _, fp, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, np.ones_like(q), w, len2pix, npix, "Cubic", 2, True,
                                         want_footprints=True)
    else:
        # footprints only: run the oracle on tiny hsml copies is not possible -> use 2D-equivalent formula per axis
        fp = np.zeros_like(out)
        p64 = pos.astype(np.float64); h = hsml.astype(np.float64) * len2pix
        for d in range(3):
            x = p64[:, d] * len2pix
            x = x + 0.5 * npix
            fp[:, 2 * d] = np.maximum(np.floor(x - h), 0)
            fp[:, 2 * d + 1] = np.minimum(np.floor(x + h), npix - 1)
    assert np.array_equal(out, fp)


@pytest.mark.parametrize("strategy", ["scatter", "gather", "auto"])
@pytest.mark.parametrize("kernel", KERNELS)
def test_deposit_2d_parity(s2g, oracle, strategy, kernel):
    pos, hsml, m, rho, q, w = random_particles(7, 6000, box=11.0, hmax=1.6)
    hsml[:300] *= 0.01        # "no pixel centre covered" branch
    q[5:50] = 0.0
    npix = 200
    len2pix = npix / 10.0
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    ctx = s2g.Context(0, strategy=strategy)
    for calc_mean in (True, False):
        got, st = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel), calc_mean=calc_mean,
                                     ctx=ctx, return_stats=True)
        ref, fp, ost = _oracle_2d(oracle, pos, hsml, m, rho, q, w, len2pix, npix, kernel, calc_mean)
        assert_parity(got, ref, what=f"{kernel}/{strategy}/calc_mean={calc_mean}")
        assert st["n_mapped"] == ost["n_mapped"]
        assert st["footprint_pixels"] == ost["footprint_pixels"]
        assert st["n_fallback"] == ost["n_fallback"]
        assert st["touched_pixels"] == ost["touched_pixels"]
    ctx.close()


@pytest.mark.parametrize("strategy", ["scatter", "gather"])
def test_deposit_2d_large_footprints_and_clipping(s2g, oracle, strategy):
    # footprints of several hundred pixels, many clipped by the image border, some entirely outside
    pos, hsml, m, rho, q, w = random_particles(9, 400, box=16.0, hmin=0.5, hmax=6.0)
    npix = 320
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    ctx = s2g.Context(0, strategy=strategy)
    got = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC6(2), calc_mean=True, ctx=ctx)
    ref, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, par.len2pix, npix, "WendlandC6", 2, True)
    assert_parity(got, ref, what="large footprints " + strategy)
    ctx.close()


@pytest.mark.parametrize("strategy", ["scatter", "gather"])
def test_deposit_2d_multi_image_and_f32(s2g, oracle, strategy):
    pos, hsml, m, rho, q, w = random_particles(11, 3000, dtype=np.float32, hmax=1.2)
    Q = np.stack([q, (q * 0 + 1).astype(np.float32), np.zeros_like(q)], axis=1)
    Q[7, :] = 0.0
    npix = 128
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    ctx = s2g.Context(0, strategy=strategy)
    got = s2g.cic_mapping_2D(pos, hsml, m, rho, Q, w, param=par, kernel=s2g.WendlandC4(2), calc_mean=True, ctx=ctx)
    ref, _ = oracle.cic_mapping_2d(pos.astype(np.float64), hsml.astype(np.float64), m.astype(np.float64),
                                   rho.astype(np.float64), Q.astype(np.float64), w.astype(np.float64), par.len2pix,
                                   npix, "WendlandC4", 2, True)
    assert got.shape == (npix * npix, 4)
    assert_parity(got, ref, what="multi-image f32 " + strategy)
    ctx.close()


def test_empty_and_degenerate_inputs(s2g):
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=32)
    z = np.zeros((0, 3)); e = np.zeros(0)
    img = s2g.cic_mapping_2D(z, e, e, e, e, e, param=par, kernel=s2g.Cubic(), calc_mean=True)
    assert img.shape == (1024, 2) and not img.any()
    out = s2g.sphMapping(z, e, e, e, e, e, param=par, kernel=s2g.Cubic(), show_progress=False)
    assert out.shape == (32, 32, 1) and not out.any()
    # all particles outside the image
    pos = np.full((10, 3), 100.0)
    one = np.ones(10)
    out = s2g.sphMapping(pos, one, one, one, one, one, param=par, kernel=s2g.Cubic(), show_progress=False)
    assert not out.any()
    with pytest.raises(s2g.S2GError):
        s2g.cic_mapping_2D(pos, one, one, one, one, one, param=par, kernel=s2g.AbstractSPHKernel(2, 99, "bogus"))
    # stokes=true through sphMapping only enforces the far->near order (cic_interpolation.jl:74-83, :152-155)
    out = s2g.sphMapping(pos, one, one, one, one, one, param=par, kernel=s2g.Cubic(), stokes=True, show_progress=False)
    assert out.shape == (32, 32, 1) and not out.any()


# ---- the reference's own known-answer tests, through the GPU path
def test_kat_2d_particle_not_overlapping_centers(s2g):
    """test/runtests.jl:718-738"""
    npix, r = 4, 10
    param = s2g.mappingParameters(x_lim=[-r, r], y_lim=[-r, r], z_lim=[-r, r], Npixels=npix)
    pos = np.array([[0.5, 0.0, 0.0]])
    hsms = np.array([1.0]); mass = np.array([1.0]); rho = np.ones(1)
    w = s2g.part_weight_physical(1, param, 1)
    mp = s2g.sphMapping(pos, hsms, mass, rho, rho, w, param=param, kernel=s2g.Cubic(), reduce_image=False,
                        show_progress=False)
    Apix = (param.x_lim[1] - param.x_lim[0]) ** 2 / npix ** 2
    assert math.isclose(Apix * mp.sum(), 1.0, rel_tol=1e-10)
    assert np.count_nonzero(mp) == 4


def test_kat_3d_mass_conservation(s2g):
    """test/runtests.jl:324-342"""
    npix, r = 200, 64
    param = s2g.mappingParameters(x_lim=[-r, r], y_lim=[-r, r], z_lim=[-r, r], Npixels=npix)
    pos = np.array([[0.0101, -0.001, 0.001]])
    w = s2g.part_weight_physical(1, param, 1)
    mp = s2g.sphMapping(pos, np.array([5.0]), np.array([3.0]), np.ones(1), np.ones(1), w, param=param, dimensions=3,
                        kernel=s2g.Cubic(), reduce_image=False, show_progress=False)
    Vpix = (param.x_lim[1] - param.x_lim[0]) ** 3 / npix ** 3
    assert math.isclose(Vpix * mp.sum(), 3.0, rel_tol=1e-8)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("reduce_image", [True, False])
def test_sphmapping_end_to_end(s2g, oracle, dtype, reduce_image):
    """centre (input precision, periodic quirk) -> filter -> deposit -> reduce, vs the oracle's sph_mapping"""
    pos, hsml, m, rho, q, w = random_particles(21, 8000, box=6.5, hmax=0.4, dtype=dtype, center=3.0)
    kw = dict(center=[3.0, 3.0, 3.0], x_size=5.4, y_size=5.4, z_size=5.4, Npixels=256, boxsize=6.0)
    par = s2g.mappingParameters(**kw)
    opar = oracle.mapping_parameters(**kw)
    p1, p2 = pos.copy(), pos.copy()
    got = s2g.sphMapping(p1, hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC4(2), calc_mean=True,
                         reduce_image=reduce_image, show_progress=False)
    ref = oracle.sph_mapping(p2, hsml, m, rho, q, w, param=opar, kernel="WendlandC4", calc_mean=True,
                             reduce_image=reduce_image)
    assert np.array_equal(p1, p2), "Pos must be recentred in place, bit-identically (Q1/Q2/Q3)"
    assert got.shape == (256, 256, 1)
    assert_parity(got, ref, what="sphMapping 2D")
    both = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC4(2), calc_mean=True,
                          return_both_maps=True, show_progress=False)
    oboth = oracle.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=opar, kernel="WendlandC4", calc_mean=True,
                               return_both_maps=True)
    assert_parity(both, oboth, what="return_both_maps")


def test_sphmapping_sort_z_quirk(s2g, oracle):
    pos, hsml, m, rho, q, w = random_particles(23, 3000, box=8.0, hmax=0.5, center=1.0)
    kw = dict(center=[1.0, 1.0, 1.0], x_size=6.0, y_size=6.0, z_size=6.0, Npixels=100)
    got = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=s2g.mappingParameters(**kw), kernel=s2g.Cubic(2),
                         calc_mean=True, sort_z=True, show_progress=False)
    ref = oracle.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=oracle.mapping_parameters(**kw), kernel="Cubic",
                             calc_mean=True, sort_z=True)
    assert_parity(got, ref, what="sort_z (Q5)")


def test_reduce_image_functions(s2g, oracle):
    rng = np.random.default_rng(3)
    n = 70
    flat = np.asfortranarray(rng.normal(size=(n * n, 3)))
    flat[::7, 2] = 0.0
    flat[::5, 2] *= -1
    for red in (True, False):
        assert np.array_equal(s2g.reduce_image_2D(flat, n, n, red), oracle.reduce_image_2d(flat, n, n, red))
    n3 = 21
    f3 = np.asfortranarray(rng.normal(size=(n3 ** 3, 2)))
    for red in (True, False):
        assert np.array_equal(s2g.reduce_image_3D(f3, n3, n3, n3, red), oracle.reduce_image_3d(f3, n3, red))


def test_size_independent_properties_large(s2g):
    """Properties that hold at any size (checked at a size the oracle could not finish quickly):
    (1) Σ weight plane = Σ_p w·m/ρ·len2pix³ for particles whose footprint is not clipped (normalisation identity of
    cic_2D.jl:187-199); (2) linearity in the mapped quantity; (3) scatter and gather strategies agree."""
    n = 300000
    pos, hsml, m, rho, q, w = random_particles(31, n, box=6.0, hmin=0.05, hmax=0.6)
    npix = 1024
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    ctx_g = s2g.Context(0, strategy="gather"); ctx_s = s2g.Context(0, strategy="scatter")
    a = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC6(2), ctx=ctx_g)
    b = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC6(2), ctx=ctx_s)
    assert_parity(a, b, rtol=1e-11, what="gather vs scatter")
    expect = np.sum(w * m / rho * par.len2pix ** 3)
    assert math.isclose(a[:, 1].sum(), expect, rel_tol=1e-11)
    c = s2g.cic_mapping_2D(pos, hsml, m, rho, 3.0 * q, w, param=par, kernel=s2g.WendlandC6(2), ctx=ctx_g)
    assert_parity(c[:, 0], 3.0 * a[:, 0], rtol=1e-12, what="linearity")
    ctx_g.close(); ctx_s.close()


@pytest.mark.parametrize("kernel", ["WendlandC6", "WendlandC8", "WendlandC4", "Quintic", "Cubic"])
def test_closed_form_normalisation_vs_numerical_sum(s2g, oracle, kernel):
    """Well-resolved, unclipped footprints use h^2*∫w instead of the pass-A sum (s2g_set_exact_norm).  Both modes must
    agree with each other to 1e-11 (design bound 5e-12) and with the oracle to the 1e-10 bar."""
    rng = np.random.default_rng(17)
    n = 300
    npix = 1024
    pos = (rng.random((n, 3)) - 0.5) * 4.0            # well inside the 10-unit image
    hsml = 0.3 + rng.random(n) * 1.2                   # 31 .. 154 pixels
    m = rng.random(n) + 0.1; rho = rng.random(n) + 0.1; q = rng.random(n) * 10; w = rng.random(n) + 0.5
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    c_fast = s2g.Context(0, strategy="gather")
    c_exact = s2g.Context(0, strategy="gather", exact_norm=True)
    a = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel), ctx=c_fast)
    b = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel), ctx=c_exact)
    assert_parity(a, b, rtol=1e-11, what="closed form vs numerical pass A")
    ref, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, par.len2pix, npix, kernel, 2, True)
    assert_parity(a, ref, what="closed form vs oracle")
    assert_parity(b, ref, what="numerical vs oracle")
    c_fast.close(); c_exact.close()


def test_distributed_cic_map_subfile_streaming(s2g, oracle, tmp_path):
    """distributed_cic_map (src/distributed_mapping/cic.jl): per-subfile maps with return_both_maps, finite-guarded
    sum, one reduce, FITS out — equals one sphMapping over the concatenated particles."""
    pos, hsml, m, rho, q, w = random_particles(61, 9000, box=9.0, hmax=0.7)
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=128)
    bounds = [0, 2000, 2000, 5500, 9000]  # 4 "sub-snapshot files", one of them empty

    def mapping_function(subfile):
        s, e = bounds[subfile], bounds[subfile + 1]
        img = s2g.sphMapping(pos[s:e].copy(), hsml[s:e], m[s:e], rho[s:e], q[s:e], w[s:e], param=par,
                             kernel=s2g.WendlandC6(2), calc_mean=True, return_both_maps=True, show_progress=False)
        if subfile == 0:
            img[17, 0] = np.nan  # poisoned entries are skipped by the master's accumulation (cic.jl:63-69)
        return img[:, :1], img[:, 1]

    fn = str(tmp_path / "map.fits")
    image = s2g.distributed_cic_map(fn, 4, mapping_function, par, 1, reduce_image=True, snap=7, units="K")
    whole = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC6(2), calc_mean=True,
                           show_progress=False)
    mask = np.ones_like(whole, dtype=bool)
    mask[17 // 128, 17 % 128, 0] = False
    assert_parity(image[mask], whole[mask], rtol=1e-12, what="distributed_cic_map vs single map")
    d, par2, snap, units = s2g.read_fits_image(fn)
    assert np.array_equal(d, image[:, :, 0]) and snap == 7 and units == "K"


@pytest.mark.parametrize("strategy", ["scatter", "gather"])
def test_non_finite_normalisation_marks_bounding_box(s2g, oracle, strategy):
    """rho = 0 (dz = Inf) or a NaN weight make pix_weight = wk*A*area_norm Inf inside the kernel and 0*Inf = NaN outside:
    the reference's `!iszero(pix_weight)` then updates the whole bounding box.  Zero weight deposits nothing."""
    pos, hsml, m, rho, q, w = random_particles(71, 400, box=9.0, hmin=0.2, hmax=1.5)
    rho[3] = 0.0
    w[10] = np.nan
    w[20] = 0.0
    m[30] = 0.0
    npix = 96
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    ctx = s2g.Context(0, strategy=strategy)
    got = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC4(2), calc_mean=True, ctx=ctx)
    ref, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, par.len2pix, npix, "WendlandC4", 2, True)
    assert np.isnan(ref).any() and np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    assert_parity(np.where(fin, got, 0.0), np.where(fin, ref, 0.0), what="finite part")
    ctx.close()


def test_map_it_like_the_precompile_workload(s2g, oracle, tmp_path):
    """src/precompile.jl:27-57: map_it on 100 random particles, 256^2, WendlandC4/C6, FITS out; positions untouched."""
    rng = np.random.default_rng(100)
    cic_pos = 15.0 * (rng.random((100, 3)) - 0.5)
    cic_hsml = 2.0 * rng.random(100); cic_mass = rng.random(100); cic_rho = rng.random(100) + 1e-3
    cic_T = 1.0e8 * rng.random(100)
    kw = dict(center=[0.0, 0.0, 0.0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=256)
    par = s2g.mappingParameters(**kw)
    for kname in ("WendlandC4", "WendlandC6"):
        p0 = cic_pos.copy()
        m = s2g.map_it(p0, cic_hsml, cic_mass, cic_rho, cic_T, cic_rho, units="T", param=par, reduce_image=True,
                       parallel=False, show_progress=False, snap=0, image_prefix=str(tmp_path / "dummy"),
                       kernel=getattr(s2g, kname)(2))
        assert np.array_equal(p0, cic_pos)
        ref = oracle.sph_mapping(cic_pos.copy(), cic_hsml, cic_mass, cic_rho, cic_T, cic_rho,
                                 param=oracle.mapping_parameters(**kw), kernel=kname, calc_mean=True)
        assert_parity(m, ref, what="map_it " + kname)
        d, _, _, units = s2g.read_fits_image(str(tmp_path / "dummy") + ".xy.fits")
        assert np.array_equal(d, m[:, :, 0]) and units == "T"


@pytest.mark.parametrize("bounce", ["1", "0"])
def test_overlapped_staging_path(s2g, oracle, bounce, monkeypatch):
    """The host-array entry points copy large inputs on a helper thread, chunk by chunk, while the sliced deposit runs
    (s2g_stage_wait per slice; csrc/s2g_api.cu).  The thresholds are lowered so that a 60 k-particle call goes through
    that machinery — many chunks, many slices, through own pinned bounce buffers (default) or the driver's pageable path —
    and must give the map of the plain path: 2D (gather + scatter bins), 3D and HEALPix."""
    monkeypatch.setenv("S2G_STAGE_MIN", "1000")
    monkeypatch.setenv("S2G_STAGE_CHUNK", "4096")
    monkeypatch.setenv("S2G_BATCH_PARTICLES", "10000")
    monkeypatch.setenv("S2G_STAGE_BOUNCE", bounce)
    ctx = s2g.Context(0)          # a fresh context: the staging thread and its buffers are created with these settings
    pos, hsml, m, rho, q, w = random_particles(61, 60000, box=6.5, hmin=0.01, hmax=0.5, center=3.0)
    hsml[:3000] *= 4.0
    kw = dict(center=[3.0, 3.0, 3.0], x_size=5.4, y_size=5.4, z_size=5.4, Npixels=192, boxsize=6.0)
    for it in range(2):           # twice: buffers and events are reused by the second call
        p1, p2 = pos.copy(), pos.copy()
        got, st = s2g.sphMapping(p1, hsml, m, rho, q, w, param=s2g.mappingParameters(**kw), kernel=s2g.WendlandC6(2),
                                 calc_mean=True, show_progress=False, ctx=ctx, return_stats=True)
        ref = oracle.sph_mapping(p2, hsml, m, rho, q, w, param=oracle.mapping_parameters(**kw), kernel="WendlandC6",
                                 calc_mean=True)
        assert np.array_equal(p1, p2)
        assert st["n_gather"] > 0 and st["n_scatter"] > 0
        assert_parity(got, ref, what=f"staged 2D, call {it}")
    kw3 = dict(kw, Npixels=40)
    got = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=s2g.mappingParameters(**kw3), kernel=s2g.Cubic(3),
                         dimensions=3, show_progress=False, ctx=ctx)
    ref = oracle.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=oracle.mapping_parameters(**kw3), kernel="Cubic",
                             dimensions=3)
    assert_parity(got, ref, what="staged 3D")
    hp = pos - 3.0
    a, wm = s2g.healpix_map(hp.copy(), hsml, m, rho, q, w, center=[0.1, -0.2, 0.05], Nside=64, kernel=s2g.WendlandC4(2),
                            show_progress=False, ctx=ctx)
    ea, ew, est = oracle.healpix_map(hp.copy(), hsml, m, rho, q, w, center=[0.1, -0.2, 0.05], nside=64, kernel="WendlandC4",
                                     exact="sens", n_workers=4)
    from util import assert_healpix_parity
    assert_healpix_parity(a, wm, ea, ew, est, what="staged HEALPix")
    ctx.close()


@pytest.mark.parametrize("kernel", ["WendlandC6", "Cubic"])
def test_scatter_classes_block_order_and_record_board(s2g, oracle, kernel, monkeypatch):
    """Round-2 scatter path at the sizes that switch its machinery on (csrc/s2g_cic2d.cu): an image larger than L2's share
    (2048^2 x 2 planes = 67 MB) and class lists above 65536 particles -> lists radix-sorted by 64x64-pixel block,
    32 particle records per warp visit, flattened footprints with pass-A weights cached in shared memory, in both the
    sub-warp (<= 64 pixels) and the warp (<= 1023 pixels) kernel; footprints above the 384-pixel cache, "no pixel centre
    covered" particles and zero quantities take the general branches.  Against the oracle, counters equal; and the same
    call with the order switched off gives the same map."""
    n = 300000
    pos, hsml, m, rho, q, w = random_particles(71, n, box=10.4, hmin=0.002, hmax=0.05)
    hsml[:4000] = 0.06 + 0.02 * np.arange(4000) / 4000.0          # 12..16 px radius: 600-1000 pixel footprints
    hsml[4000:9000] *= 0.02                                         # no pixel centre covered
    q[9000:9500] = 0.0
    npix = 2048
    len2pix = npix / 10.0
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    ctx = s2g.Context(0)
    got, st = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel), calc_mean=True,
                                 ctx=ctx, return_stats=True)
    ref, fp, ost = _oracle_2d(oracle, pos, hsml, m, rho, q, w, len2pix, npix, kernel, True)
    assert_parity(got, ref, what=f"ordered scatter classes, {kernel}")
    for k in ("n_mapped", "footprint_pixels", "n_fallback", "touched_pixels"):
        assert st[k] == ost[k], k
    assert st["n_scatter"] > 2 * 65536
    monkeypatch.setenv("S2G_2D_ORDER", "0")
    plain = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel), calc_mean=True, ctx=ctx)
    assert_parity(got, plain, rtol=1e-12, what="block order on vs off")
    ctx.close()


@pytest.mark.parametrize("both", [False, True])
def test_result_map_copied_back_by_the_staging_threads(s2g, oracle, both, monkeypatch):
    """A large result map goes back to the caller's pageable array through the staging threads' pinned bounce buffers
    (unstage_output, csrc/s2g_api.cu): 4-MB pieces, several threads, up to three copies in flight per thread.  The
    thresholds are lowered so that a 1536^2 map (19 MB = 5 pieces; 38 MB with return_both_maps) takes that path; the map
    must be the oracle's, and identical to the one the plain copy returns."""
    monkeypatch.setenv("S2G_STAGE_MIN", "1000")
    monkeypatch.setenv("S2G_STAGE_CHUNK", "4096")
    monkeypatch.setenv("S2G_UNSTAGE_MIN", "65536")
    ctx = s2g.Context(0)
    pos, hsml, m, rho, q, w = random_particles(81, 30000, box=6.5, hmin=0.005, hmax=0.05, center=3.0)
    kw = dict(center=[3.0, 3.0, 3.0], x_size=6.0, y_size=6.0, z_size=6.0, Npixels=1536)
    got = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=s2g.mappingParameters(**kw), kernel=s2g.WendlandC4(2),
                         calc_mean=True, show_progress=False, ctx=ctx, return_both_maps=both)
    ref = oracle.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=oracle.mapping_parameters(**kw), kernel="WendlandC4",
                             calc_mean=True, return_both_maps=both)
    assert_parity(got, ref, what=f"threaded result copy, return_both_maps={both}")
    monkeypatch.setenv("S2G_UNSTAGE_MIN", str(1 << 40))
    plain = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=s2g.mappingParameters(**kw), kernel=s2g.WendlandC4(2),
                           calc_mean=True, show_progress=False, ctx=ctx, return_both_maps=both)
    assert_parity(got, plain, rtol=1e-12, what="threaded vs plain result copy")
    ctx.close()
