import ctypes as C
import math

import numpy as np
import pytest

from util import KERNELS, assert_healpix_parity, assert_parity, kern, ncores, random_particles

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", ["Cubic", "WendlandC6", "Quintic"])
def test_deposit_3d_parity(s2g, oracle, kernel):
    pos, hsml, m, rho, q, w = random_particles(41, 3000, box=11.0, hmax=1.0)
    hsml[:200] *= 0.02
    q[3:30] = 0.0
    npix = 48
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    for calc_mean in (False, True):
        got, st = s2g.cic_mapping_3D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel, 3),
                                     calc_mean=calc_mean, return_stats=True)
        ref, ost = oracle.cic_mapping_3d(pos, hsml, m, rho, q, w, par.len2pix, npix, kernel, 3, calc_mean)
        assert_parity(got, ref, what=f"3D {kernel} calc_mean={calc_mean}")
        for k in ("n_mapped", "footprint_pixels", "n_fallback", "touched_pixels"):
            assert st[k] == ost[k], k


def test_sphmapping_3d_end_to_end(s2g, oracle):
    pos, hsml, m, rho, q, w = random_particles(43, 4000, box=7.0, hmax=0.5, dtype=np.float32, center=3.0)
    kw = dict(center=[3.0, 3.0, 3.0], x_size=6.0, y_size=6.0, z_size=6.0, Npixels=40, boxsize=6.0)
    for red in (True, False):
        p1, p2 = pos.copy(), pos.copy()
        got = s2g.sphMapping(p1, hsml, m, rho, q, w, param=s2g.mappingParameters(**kw), kernel=s2g.WendlandC6(3),
                             dimensions=3, reduce_image=red, show_progress=False)
        ref = oracle.sph_mapping(p2, hsml, m, rho, q, w, param=oracle.mapping_parameters(**kw), kernel="WendlandC6",
                                 dimensions=3, reduce_image=red)
        assert np.array_equal(p1, p2)
        assert got.shape == (40, 40, 40)
        assert_parity(got, ref, what=f"sphMapping 3D reduce={red}")


# ---------------------------------------------------------------- HEALPix
def _gpu_pixels(s2g, pos, radius, nside, cap=1 << 16):
    from sphtogrid_b200 import _lib
    out = np.zeros(cap, dtype=np.int64)
    cnt = C.c_int64(0)
    p = (C.c_double * 3)(*pos)
    _lib.check(s2g.lib().s2g_healpix_pixels(s2g.default_context().handle, p, float(radius), nside, _lib.ptr(out), cap,
                                            C.byref(cnt)))
    assert cnt.value <= cap
    return out[:cnt.value]


@pytest.mark.parametrize("nside", [1, 4, 64, 512, 2048])
def test_healpix_pixel_lists_identical(s2g, oracle, nside):
    """pixel-index work is bit-exact: same pixel SET per particle as the oracle's contributing_pixels"""
    rng = np.random.default_rng(nside)
    L = oracle.lib()
    buf = np.zeros(1 << 16, dtype=np.int64)
    ang = math.sqrt(4 * math.pi / (12 * nside * nside))
    cases = []
    for _ in range(150):
        v = rng.normal(size=3)
        cases.append((v, min(3.0, ang * rng.uniform(0.05, 40.0))))
    cases += [(np.array([0.0, 0.0, 1.0]), 5 * ang), (np.array([0.0, 0.0, -1.0]), 5 * ang),
              (np.array([1e-9, 0.0, 1.0]), 3 * ang), (np.array([1.0, -1e-12, 0.0]), 4 * ang),
              (np.array([1.0, 0.0, 0.0]), 4 * ang), (np.array([-1.0, 1e-13, 0.3]), 10 * ang)]
    for v, r in cases:
        if (math.pi * r * r) / (ang * ang) > 50000:
            r = ang * 100
        n = L.s2go_hp_contributing_pixels(nside, v.ctypes.data_as(C.POINTER(C.c_double)), r,
                                          buf.ctypes.data_as(C.POINTER(C.c_int64)), buf.size)
        assert n >= 0
        ref = buf[:n]
        got = _gpu_pixels(s2g, v, r, nside)
        assert np.array_equal(np.sort(got), np.sort(ref)), (nside, v, r)


@pytest.mark.parametrize("kernel", ["WendlandC4", "Cubic"])
@pytest.mark.parametrize("nside", [32, 128])
def test_healpix_deposit_parity_resolved(s2g, oracle, kernel, nside):
    """Discs resolved by >= 3 pixels across the radius: the 1e-10 bar of the north_star holds."""
    # dense coverage (every pixel is covered by tens of discs): a pixel fed ONLY by one disc's rim, where
    # w ~ (1-u)^6 -> 0, turns the ~1e-13 conditioning error of u = acos(.)/proj_hsml into an unbounded relative one
    rng = np.random.default_rng(5)
    n = 1500 if nside == 32 else 24000
    ang = math.sqrt(4 * math.pi / (12 * nside * nside))
    pos = rng.normal(size=(n, 3)) * 60.0
    dist = np.linalg.norm(pos, axis=1)
    hsml = dist * np.sin(ang * rng.uniform(3.0, 12.0, n))      # proj_hsml = 3..12 pixel diameters
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 1e4; w = rng.random(n) + 0.5
    q[200:220] = 0.0
    for calc_mean in (True, False):
        a, wm, st = s2g.healpix_deposit(pos, hsml, m, rho, q, w, nside, kern(s2g, kernel), calc_mean,
                                        return_stats=True)
        ra, rw, ost = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, kernel, 2, calc_mean)
        assert st["n_mapped"] == ost["n_mapped"] and st["n_fallback"] == ost["n_fallback"] == 0
        assert st["touched_pixels"] == ost["touched_pixels"]
        assert_parity(wm, rw, rtol=1e-10, what=f"healpix weight map nside={nside}")
        assert_parity(a, ra, rtol=1e-10, what=f"healpix map nside={nside}")
        assert math.isclose(wm.sum(), rw.sum(), rel_tol=1e-12)
        ea, ew, est = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, kernel, 2, calc_mean, n_workers=ncores(),
                                             exact="sens")
        assert_healpix_parity(a, wm, ea, ew, est, what=f"healpix vs extended precision, nside={nside}")


@pytest.mark.parametrize("nside", [32, 256, 2048])
def test_healpix_deposit_parity_all_regimes(s2g, oracle, nside):
    """Everything at once: sub-pixel discs (fallback branch), barely resolved discs, particles closer than hsml
    (skipped), zero quantities; sparse coverage, so many pixels hold nothing but one disc's rim.
    Integers (counters, pixel sets) against the literal Float64 oracle: exact.  Maps against the extended-precision
    arbiter at 1e-10 (+ the Float64 unit-vector resolution term, util.assert_healpix_parity) — no 5e-8 allowance: the
    literal Float64 acos form is what sits 1e-8 .. 1e-6 away from the exact value here, not the GPU."""
    rng = np.random.default_rng(6)
    n = 1500
    pos = rng.normal(size=(n, 3)) * 60.0
    hsml = rng.random(n) * 6.0 + 0.2
    hsml[:100] *= 0.01
    if nside == 2048:
        hsml *= 0.1
    pos[100:110] *= 0.01
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 1e4; w = rng.random(n) + 0.5
    q[200:220] = 0.0
    for calc_mean in (True, False):
        a, wm, st = s2g.healpix_deposit(pos, hsml, m, rho, q, w, nside, s2g.WendlandC4(2), calc_mean,
                                        return_stats=True)
        ra, rw, ost = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, "WendlandC4", 2, calc_mean)
        for k in ("n_mapped", "n_fallback", "touched_pixels"):
            assert st[k] == ost[k], k
        assert st["n_fallback"] > 0
        ea, ew, est = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, "WendlandC4", 2, calc_mean,
                                             n_workers=ncores(), exact="sens")
        assert_healpix_parity(a, wm, ea, ew, est, what=f"healpix all regimes nside={nside}")
        assert math.isclose(wm.sum(), ew.sum(), rel_tol=1e-12) and math.isclose(a.sum(), ea.sum(), rel_tol=1e-12)


def test_healpix_map_api(s2g, oracle):
    """src/precompile.jl:9-23 fixture (7 cluster positions, Nside 128) through healpix_map, incl. Q1"""
    center = np.array([247.980, 245.480, 255.290]) * 1.0e3
    hp_pos = np.array([[225160.875, 256677.5625, 242031.765625], [245040.453125, 327781.84375, 246168.6875],
                       [252454.0, 238890.296875, 233772.890625], [202408.359375, 245067.234375, 249614.09375],
                       [176696.22, 266631.8, 315976.38], [307184.78125, 247627.078125, 230736.484375],
                       [244450.578125, 255851.78125, 253668.1875]])
    hp_hsml = np.array([1771.7005615234375, 2177.5625, 714.1318359375, 798.90478515625, 1067.0657,
                        1813.0419921875, 1711.311767578125])
    one = np.ones(7)
    p1, p2 = hp_pos.copy(), hp_pos.copy()
    a, wm = s2g.healpix_map(p1, hp_hsml, one, one, one, one, center=center, kernel=s2g.WendlandC4(2), Nside=128,
                            show_progress=False)
    ra, rw = oracle.healpix_map(p2, hp_hsml, one, one, one, one, center=center, kernel="WendlandC4", nside=128)
    assert np.array_equal(p1, p2)
    assert_parity(a, ra, what="precompile fixture map")
    assert_parity(wm, rw, what="precompile fixture weights")
    ea, ew, est = oracle.healpix_map(hp_pos.copy(), hp_hsml, one, one, one, one, center=center, kernel="WendlandC4",
                                     nside=128, exact="sens")
    assert_healpix_parity(a, wm, ea, ew, est, what="precompile fixture vs extended precision")
    with pytest.raises(IndexError):
        s2g.healpix_map(hp_pos.copy(), hp_hsml, one, one, np.array([1, 1, 0, 1, 1, 1, 1.0]), one, center=center,
                        kernel=s2g.WendlandC4(2), Nside=128, calc_mean=False)


# ---------------------------------------------------------------- CIC / TSC stencils
@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("dims", [2, 3])
@pytest.mark.parametrize("periodic", [False, True])
def test_stencils_parity(s2g, oracle, order, dims, periodic):
    pos, hsml, m, rho, q, w = random_particles(51, 20000, box=10.4)
    npix = 64 if dims == 2 else 32
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    fn = s2g.cic_deposit if order == 2 else s2g.tsc_deposit
    got = fn(pos, q, param=par, dimensions=dims, average=False, periodic=periodic)
    ref = oracle.stencil_deposit(order, dims, pos, q, par.len2pix, npix, periodic)
    assert_parity(got, ref, rtol=1e-12, what=f"stencil order={order} dims={dims}")
    if not periodic:
        inside = np.all(np.abs(pos[:, :dims]) < 5.0 - 1.5 * par.pixelSideLength, axis=1)
        if dims == 3:
            assert got[:, 1].sum() <= pos.shape[0] + 1e-9
        assert got[:, 1].sum() >= inside.sum() - 1e-6
    else:
        if dims == 3:
            assert math.isclose(got[:, 1].sum(), pos.shape[0], rel_tol=1e-12)
    avg = fn(pos, q, param=par, dimensions=dims, average=True, periodic=periodic)
    assert_parity(avg.ravel(), oracle.stencil_average(ref), rtol=1e-12, what="average")


def test_healpix_map_device_filter_sort_quirk(s2g, oracle):
    """healpix_map with a shell that does NOT contain every particle: the reference deposits `sorted[mask]` (mask in
    original order applied to the far-to-near permutation, filter_particles.jl:33-41); the device path reproduces
    exactly that selection (stable radix sort of the radii)."""
    rng = np.random.default_rng(12)
    n = 4000
    pos = rng.normal(size=(n, 3)) * 40.0 + np.array([10.0, -5.0, 3.0])
    pos[::7] = pos[1::7][: len(pos[::7])]          # exact ties in the radii (stable-sort order matters)
    hsml = rng.random(n) * 3.0 + 0.5
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 10 + 1; w = rng.random(n) + 0.5
    center = np.array([10.0, -5.0, 3.0])
    for rl in ([20.0, 60.0], [0.0, 45.0], [0.0, np.inf]):
        p1, p2 = pos.copy(), pos.copy()
        a, wm = s2g.healpix_map(p1, hsml, m, rho, q, w, center=center, radius_limits=rl, Nside=64,
                                kernel=s2g.WendlandC4(2), show_progress=False)
        ra, rw = oracle.healpix_map(p2, hsml, m, rho, q, w, center=center, radius_limits=rl, nside=64,
                                    kernel="WendlandC4")
        assert np.array_equal(p1, p2)
        ea, ew, est = oracle.healpix_map(pos.copy(), hsml, m, rho, q, w, center=center, radius_limits=rl, nside=64,
                                         kernel="WendlandC4", exact="sens", n_workers=ncores())
        assert_healpix_parity(a, wm, ea, ew, est, what=f"healpix_map, shell {rl}")
        assert math.isclose(wm.sum(), rw.sum(), rel_tol=1e-12)


def test_size_independent_properties_3d_and_healpix_large(s2g):
    """Properties that hold at any size, checked well beyond what the oracle finishes quickly:
    3D: Σ weight plane = Σ_p w·(m/ρ)·len2pix⁴ for unclipped particles (volume_norm identity, cic_3D.jl:167-188) and
    linearity in the mapped quantity; HEALPix: Σ weight map = Σ_p w·(m/ρ)/(ang_pix·Δx)² (main.jl:32-38, 188-193)."""
    n = 400000
    pos, hsml, m, rho, q, w = random_particles(81, n, box=7.0, hmin=0.05, hmax=0.45)
    npix = 160
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    a = s2g.cic_mapping_3D(pos, hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC6(3))
    expect = np.sum(w * m / rho * par.len2pix ** 4)
    assert math.isclose(a[:, 1].sum(), expect, rel_tol=1e-11)
    b = s2g.cic_mapping_3D(pos, hsml, m, rho, -2.5 * q, w, param=par, kernel=s2g.WendlandC6(3))
    assert_parity(b[:, 0], -2.5 * a[:, 0], rtol=1e-12, what="3D linearity")
    assert_parity(b[:, 1], a[:, 1], rtol=1e-12, what="3D weight plane independent of the quantity")
    # HEALPix
    nside = 512
    ang = math.sqrt(4 * math.pi / (12 * nside * nside))
    rng = np.random.default_rng(82)
    hp = rng.normal(size=(200000, 3)) * 80.0
    dist = np.linalg.norm(hp, axis=1)
    hh = np.minimum(rng.random(200000) * 2.0 + 0.05, 0.9 * dist)
    mm = rng.random(200000) + 0.5; rr = rng.random(200000) + 0.5; qq = rng.random(200000); ww = rng.random(200000) + 0.5
    amap, wmap = s2g.healpix_deposit(hp, hh, mm, rr, qq, ww, nside, s2g.WendlandC4(2), True)
    expect = np.sum(ww * (mm / rr) / (ang * dist) ** 2)
    assert math.isclose(wmap.sum(), expect, rel_tol=1e-11)
    assert math.isclose(amap.sum(), np.sum(qq * ww * (mm / rr) / (ang * dist) ** 2), rel_tol=1e-11)


@pytest.mark.parametrize("strategy", ["scatter", "gather", "auto"])
@pytest.mark.parametrize("kernel", ["Cubic", "WendlandC4"])
def test_deposit_3d_strategies(s2g, oracle, strategy, kernel):
    """3D: warp-per-particle scatter, tile-owning gather and the per-particle AUTO split all match the oracle;
    footprints from sub-cell (no-centre branch) to larger than the grid, clipped by the border, zero quantities."""
    pos, hsml, m, rho, q, w = random_particles(47, 2500, box=12.0, hmin=0.01, hmax=2.2)
    hsml[:150] *= 0.02
    hsml[150:160] = 9.0
    q[5:40] = 0.0
    rho[77] = 0.0       # Inf normalisation marks the whole box of that particle
    npix = 56
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    ctx = s2g.Context(0, strategy=strategy)
    for calc_mean in (False, True):
        got, st = s2g.cic_mapping_3D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel, 3),
                                     calc_mean=calc_mean, ctx=ctx, return_stats=True)
        ref, ost = oracle.cic_mapping_3d(pos, hsml, m, rho, q, w, par.len2pix, npix, kernel, 3, calc_mean)
        assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(np.isinf(got), np.isinf(ref))
        fin = np.isfinite(ref)
        assert_parity(np.where(fin, got, 0.0), np.where(fin, ref, 0.0), what=f"3D {kernel}/{strategy}")
        for k in ("n_mapped", "footprint_pixels", "n_fallback"):
            assert st[k] == ost[k], k
    ctx.close()


@pytest.mark.parametrize("coop_rings", ["0", "8", "64"])
def test_healpix_cooperative_heavy_particles(s2g, oracle, coop_rings, monkeypatch):
    """Particles whose disc spans many rings are deposited by a whole CTA (ring batches dealt to its 8 warps, pass-A
    sums combined in shared memory); S2G_HP_COOP_RINGS sets the split (0 = off, default 512).  Same maps, same
    counters as the oracle whichever way the particles are split — incl. discs over the poles and the centre-pixel and
    'no pixel centre covered' branches."""
    monkeypatch.setenv("S2G_HP_COOP_RINGS", coop_rings)
    nside = 64
    rng = np.random.default_rng(31)
    n = 600
    ang = math.sqrt(4 * math.pi / (12 * nside * nside))
    pos = rng.normal(size=(n, 3)) * 50.0
    pos[:6] = [[0, 0, 40.0], [0, 0, -40.0], [1e-7, 0, 30.0], [30.0, 1e-9, 0], [-20.0, 0, 20.0], [0, -35.0, 1.0]]
    dist = np.linalg.norm(pos, axis=1)
    hsml = dist * np.sin(ang * rng.uniform(3.0, 12.0, n))
    hsml[:40] = dist[:40] * np.sin(ang * rng.uniform(20.0, 70.0, 40))   # 40..140 rings: "heavy" for 8 and 64
    hsml[40:60] = dist[40:60] * np.sin(ang * 0.05)                     # sub-pixel
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 1e4; w = rng.random(n) + 0.5
    q[100:110] = 0.0
    for calc_mean in (True, False):
        a, wm, st = s2g.healpix_deposit(pos, hsml, m, rho, q, w, nside, s2g.WendlandC4(2), calc_mean,
                                        return_stats=True)
        ra, rw, ost = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, "WendlandC4", 2, calc_mean)
        for k in ("n_mapped", "n_fallback", "touched_pixels"):
            assert st[k] == ost[k], (k, st[k], ost[k])
        ea, ew, est = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, "WendlandC4", 2, calc_mean,
                                             n_workers=ncores(), exact="sens")
        assert_healpix_parity(a, wm, ea, ew, est, what=f"coop={coop_rings}")
        assert math.isclose(wm.sum(), rw.sum(), rel_tol=1e-12) and math.isclose(a.sum(), ra.sum(), rel_tol=1e-12)


# ---------------------------------------------------------------- HEALPix tile-gather (csrc/s2g_hpgather.cu)
@pytest.mark.parametrize("nside,kernel", [(64, "WendlandC4"), (256, "WendlandC4"), (256, "Cubic"), (256, "Quintic"),
                                          (256, "WendlandC2"), (256, "WendlandC6"), (256, "WendlandC8"),
                                          (2048, "WendlandC4")])
def test_healpix_tile_gather(s2g, oracle, nside, kernel, monkeypatch):
    """Pass B as a tile-gather (band of 16 rings x sector of <= 64 pixels owned by a CTA, pixels in registers, no
    atomics in the inner loop): same counters as the oracle, maps at the bar of the extended-precision arbiter, for
    discs from 3 to ~100 pixels radius incl. discs next to the poles, across phi = 0 and in the polar caps; particles
    the gather cannot take (sub-pixel discs, discs over a pole, fallback branch) fall back to the scatter walk."""
    monkeypatch.setenv("S2G_HP_GATHER_MIN_PIXELS", "3")
    rng = np.random.default_rng(nside + len(kernel))
    n = 3000 if nside < 2048 else 1200
    ang = math.sqrt(4 * math.pi / (12 * nside * nside))
    pos = rng.normal(size=(n, 3)) * 50.0
    # special directions: next to the poles, across phi = 0 / 2 pi, inside the polar caps
    pos[:8] = [[0.3, 0.1, 40.0], [-0.2, 0.3, -40.0], [30.0, 1e-9, 1.0], [30.0, -1e-9, -2.0], [25.0, 0.01, 20.0],
               [5.0, -0.02, 30.0], [-3.0, 0.0, 35.0], [1.0, 1.0, -33.0]]
    dist = np.linalg.norm(pos, axis=1)
    rad = ang * rng.uniform(3.0, 40.0, n)
    rad[:200] = ang * rng.uniform(40.0, 110.0, 200)          # large discs, many bands and sectors
    rad[200:260] = ang * rng.uniform(0.05, 2.0, 60)          # sub-pixel / barely resolved: scatter walk
    rad = np.minimum(rad, 0.19)
    hsml = dist * np.sin(rad)
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 1e4; w = rng.random(n) + 0.5
    q[300:320] = 0.0
    for calc_mean in (True, False):
        a, wm, st = s2g.healpix_deposit(pos, hsml, m, rho, q, w, nside, kern(s2g, kernel), calc_mean, return_stats=True)
        ra, rw, ost = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, kernel, 2, calc_mean)
        assert st["n_pairs"] > 0, "the tile-gather did not run"
        for k in ("n_mapped", "n_fallback", "touched_pixels"):
            assert st[k] == ost[k], (k, st[k], ost[k])
        ea, ew, est = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, kernel, 2, calc_mean, n_workers=ncores(),
                                             exact="sens")
        assert_healpix_parity(a, wm, ea, ew, est, what=f"tile-gather nside={nside} {kernel}")
        assert math.isclose(wm.sum(), ew.sum(), rel_tol=1e-12) and math.isclose(a.sum(), ea.sum(), rel_tol=1e-12)
    # gather on == gather off (scatter walk only) to rounding
    monkeypatch.setenv("S2G_HP_GATHER", "0")
    a0, w0, st0 = s2g.healpix_deposit(pos, hsml, m, rho, q, w, nside, kern(s2g, kernel), False, return_stats=True)
    assert st0["n_pairs"] == 0
    assert_healpix_parity(a0, w0, ea, ew, est, what="scatter walk")


def test_healpix_tile_gather_slices(s2g, oracle, monkeypatch):
    """The gather list is walked in slices (S2G_HP_BATCH_PARTICLES) and a slice is halved when its pair count exceeds
    S2G_PAIR_CAP: same maps and counters whatever the slicing."""
    monkeypatch.setenv("S2G_HP_GATHER_MIN_PIXELS", "3")
    nside = 128
    rng = np.random.default_rng(77)
    n = 6000
    ang = math.sqrt(4 * math.pi / (12 * nside * nside))
    pos = rng.normal(size=(n, 3)) * 50.0
    dist = np.linalg.norm(pos, axis=1)
    hsml = dist * np.sin(ang * rng.uniform(3.0, 30.0, n))
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 1e4; w = rng.random(n) + 0.5
    ref = None
    for batch, cap in (("100000000", "1000000000"), ("1500", "1000000000"), ("4096", "3000")):
        monkeypatch.setenv("S2G_HP_BATCH_PARTICLES", batch)
        monkeypatch.setenv("S2G_PAIR_CAP", cap)
        a, wm, st = s2g.healpix_deposit(pos, hsml, m, rho, q, w, nside, s2g.WendlandC4(2), True, return_stats=True)
        if ref is None:
            ref = (a, wm, st)
            ea, ew, est = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, "WendlandC4", 2, True,
                                                 n_workers=ncores(), exact="sens")
            assert_healpix_parity(a, wm, ea, ew, est, what="one slice")
            assert st["touched_pixels"] == est["touched_pixels"] and st["n_mapped"] == est["n_mapped"]
        else:
            for k in ("n_mapped", "n_fallback", "touched_pixels", "n_pairs"):
                assert st[k] == ref[2][k], (k, batch, cap)
            assert_parity(wm, ref[1], rtol=1e-12, what=f"slices {batch}/{cap}")
            assert_parity(a, ref[0], rtol=1e-12, what=f"slices {batch}/{cap}")


def test_healpix_map_strict_reference_opt_out(s2g, oracle):
    """`strict_reference=False` maps exactly the particles inside the shell (with Bin_q > 0 when
    calc_mean=False) instead of reproducing the reference's sorted[mask] / BoundsError quirks."""
    rng = np.random.default_rng(21)
    n = 3000
    center = np.array([1.0, 2.0, -1.0])
    pos = rng.normal(size=(n, 3)) * 40.0 + center
    hsml = rng.random(n) * 3.0 + 0.5
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 10; w = rng.random(n) + 0.5
    q[::5] = 0.0
    rl = [20.0, 60.0]
    for calc_mean in (True, False):
        p1 = pos.copy()
        a, wm = s2g.healpix_map(p1, hsml, m, rho, q, w, center=center, radius_limits=rl, Nside=64,
                                kernel=s2g.WendlandC4(2), show_progress=False, calc_mean=calc_mean,
                                strict_reference=False)
        assert np.array_equal(p1, pos - center)
        d = pos - center
        r = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2)
        keep = (r >= rl[0]) & (r <= rl[1])
        if not calc_mean:
            keep &= q > 0
        ea, ew, est = oracle.healpix_deposit(d[keep], hsml[keep], m[keep], rho[keep], q[keep], w[keep], 64, "WendlandC4",
                                             2, calc_mean, n_workers=ncores(), exact="sens")
        assert_healpix_parity(a, wm, ea, ew, est, what=f"strict_reference=False calc_mean={calc_mean}")


@pytest.mark.parametrize("order,dims,npix", [(2, 3, 160), (3, 3, 160), (2, 2, 2048), (3, 2, 2048)])
def test_stencils_block_ordered_deposit(s2g, oracle, order, dims, npix, monkeypatch):
    """CIC / TSC on a grid larger than L2's share (160^3 / 2048^2 cells x 2 planes >= 65 MB) with more than 65536
    particles: the particles are deposited in the key order of their 16^3-cell (64^2-pixel) block (csrc/s2g_misc.cu,
    k_stencil_keys + radix sort).  Against the oracle, periodic and clipped, and against the unordered deposit."""
    pos, hsml, m, rho, q, w = random_particles(52, 150000, box=10.6)
    pos[:50] = np.nan_to_num(pos[:50]) * 40.0            # far outside the box: clamped keys, clipped / wrapped cells
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    fn = s2g.cic_deposit if order == 2 else s2g.tsc_deposit
    for periodic in (False, True):
        got = fn(pos, q, param=par, dimensions=dims, average=False, periodic=periodic)
        ref = oracle.stencil_deposit(order, dims, pos, q, par.len2pix, npix, periodic)
        assert_parity(got, ref, rtol=1e-12, what=f"ordered stencil order={order} dims={dims} periodic={periodic}")
    monkeypatch.setenv("S2G_STENCIL_ORDER", "0")
    plain = fn(pos, q, param=par, dimensions=dims, average=False, periodic=True)
    assert_parity(got, plain, rtol=1e-12, what="stencil order on vs off")
