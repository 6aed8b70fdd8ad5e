"""HEALPix RING pixelisation of the oracle (restated from the published HEALPix algorithms that Healpix.jl ports;
"parity unpinned" against Healpix.jl itself — see oracle/s2g_oracle.c header) checked for self-consistency:
brute-force disc membership, ang2pix/pix2ang round trips, pixel-centre geometry, equal-area coverage."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import numpy_mirror as nm


def _query(oracle, nside, th, ph, r):
    cap = 12 * nside * nside + 8
    buf = np.zeros(cap, dtype=np.int64)
    n = oracle.lib().s2go_hp_query_disc_ring(nside, th, ph, r, buf.ctypes.data_as(C.POINTER(C.c_int64)), cap)
    assert n >= 0
    return buf[:n]


@pytest.mark.parametrize("nside", [1, 2, 4, 8, 16])
def test_pix2ang_matches_ring_geometry_and_roundtrip(oracle, nside):
    L = oracle.lib()
    th = C.c_double(); ph = C.c_double()
    for pix in range(12 * nside * nside):
        L.s2go_hp_pix2ang_ring(nside, pix, C.byref(th), C.byref(ph))
        z, phi, _ = nm.hp_pix_center(nside, pix)
        assert math.cos(th.value) == pytest.approx(z, abs=1e-14)
        assert ph.value == pytest.approx(phi, abs=1e-13)
        assert L.s2go_hp_ang2pix_ring(nside, th.value, ph.value) == pix
        v = np.zeros(3)
        L.s2go_hp_pix2vec_ring(nside, pix, v.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.linalg.norm(v) == pytest.approx(1.0, abs=1e-15)


def test_ang2pix_equal_area(oracle):
    # random directions land in every pixel about equally often
    rng = np.random.default_rng(5)
    nside = 4
    n = 200000
    z = rng.uniform(-1, 1, n); phi = rng.uniform(0, 2 * math.pi, n)
    L = oracle.lib()
    cnt = np.zeros(12 * nside * nside, dtype=np.int64)
    for a, b in zip(np.arccos(z), phi):
        cnt[L.s2go_hp_ang2pix_ring(nside, a, b)] += 1
    exp = n / cnt.size
    assert cnt.min() > 0.85 * exp and cnt.max() < 1.15 * exp


@pytest.mark.parametrize("nside", [4, 8, 16])
def test_query_disc_vs_brute_force(oracle, nside):
    rng = np.random.default_rng(nside)
    cases = [(0.0, 0.3, 0.4), (math.pi, 1.0, 0.5), (1e-3, 2.0, 0.2), (math.pi - 1e-3, 4.0, 0.3),
             (math.pi / 2, 0.0, 0.25), (math.pi / 2, 6.28, 0.7), (0.8, 3.0, 3.0), (1.2, 5.0, 1e-4)]
    for _ in range(40):
        cases.append((math.acos(rng.uniform(-1, 1)), rng.uniform(0, 2 * math.pi), rng.uniform(0.01, 1.5)))
    for th, ph, r in cases:
        got = _query(oracle, nside, th, ph, r)
        assert len(np.unique(got)) == len(got)
        ref = nm.hp_brute_disc(nside, th, ph, r)
        # boundary pixels (centre within 1e-9 rad of the rim) may legitimately differ; compare away from the rim
        v = np.array([math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)])
        sure_in, sure_out = set(), set()
        for pix in range(12 * nside * nside):
            z, p2, _ = nm.hp_pix_center(nside, pix)
            s = math.sqrt(max(0.0, (1 - z) * (1 + z)))
            d = math.acos(max(-1.0, min(1.0, v[0] * s * math.cos(p2) + v[1] * s * math.sin(p2) + v[2] * z)))
            if d < r - 1e-9:
                sure_in.add(pix)
            elif d > r + 1e-9:
                sure_out.add(pix)
        g = set(got.tolist())
        assert sure_in <= g, (th, ph, r, sorted(sure_in - g))
        assert not (g & sure_out), (th, ph, r, sorted(g & sure_out))
        assert abs(len(g) - len(ref)) <= 2


def test_query_disc_full_sky(oracle):
    got = _query(oracle, 4, 1.0, 1.0, 3.2)
    assert np.array_equal(np.sort(got), np.arange(12 * 16))


def test_contributing_pixels_always_has_centre(oracle):
    nside = 64
    L = oracle.lib()
    rng = np.random.default_rng(3)
    buf = np.zeros(4096, dtype=np.int64)
    for _ in range(200):
        p = rng.normal(size=3) * 100
        n = L.s2go_hp_contributing_pixels(nside, p.ctypes.data_as(C.POINTER(C.c_double)), 1e-5,
                                          buf.ctypes.data_as(C.POINTER(C.c_int64)), 4096)
        th = C.c_double(); ph = C.c_double()
        L.s2go_hp_vec2ang(p[0], p[1], p[2], C.byref(th), C.byref(ph))
        assert n == 1 and buf[0] == L.s2go_hp_ang2pix_ring(nside, th.value, ph.value)


def test_healpix_deposit_mass_and_fallback(oracle):
    """Σ weight_map·Ω_pix·Δx²-type invariant: for one well-resolved particle the un-normalised weights sum to
    area_norm·Σ(wk·A) = area·dz·w (the normalisation identity of main.jl:32-38 / pixel_weights.jl:121-137)."""
    nside = 64
    npix = 12 * nside * nside
    ang_pix = math.sqrt(4 * math.pi / npix)
    pos = np.array([[30.0, -20.0, 50.0], [1000.0, 3.0, -5.0]])
    hsml = np.array([6.0, 0.5])  # second particle: sub-pixel -> fallback branch
    m = np.array([2.0, 3.0]); rho = np.array([0.5, 0.25]); q = np.array([7.0, 11.0]); w = np.array([1.5, 2.5])
    amap, wmap, st = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, "WendlandC4", 2, True)
    assert st["n_mapped"] == 2 and st["n_fallback"] == 1
    exp = 0.0
    for i in range(2):
        dist = np.linalg.norm(pos[i])
        dz = 2 * hsml[i]
        area = (m[i] / rho[i]) / dz
        exp += area * (dz / (ang_pix * dist) ** 2) * w[i]
    assert wmap.sum() == pytest.approx(exp, rel=1e-12)
    assert amap.sum() == pytest.approx(sum(q[i] * (m[i] / rho[i]) / (ang_pix * np.linalg.norm(pos[i])) ** 2 * w[i]
                                           for i in range(2)), rel=1e-12)
    # particle closer than its hsml is skipped (main.jl:172-174)
    a2, w2, st2 = oracle.healpix_deposit(np.array([[0.1, 0.0, 0.0]]), np.array([1.0]), m[:1], rho[:1], q[:1], w[:1],
                                         nside)
    assert st2["n_mapped"] == 0 and w2.sum() == 0


def test_healpix_map_filter_sort_quirks(oracle):
    rng = np.random.default_rng(8)
    n = 50
    pos = rng.normal(size=(n, 3)) * 50 + 100
    hsml = rng.random(n) * 3 + 0.5
    m = np.ones(n); rho = np.ones(n); q = rng.random(n); w = np.ones(n)
    center = np.array([100.0, 100.0, 100.0])
    p0 = pos.copy()
    a, wm = oracle.healpix_map(pos, hsml, m, rho, q, w, center=center, radius_limits=[0.0, np.inf], nside=16)
    assert np.allclose(pos, p0 - center)  # Q1: mutated in place
    # all particles selected -> equals a plain deposit of the recentred set (order-independent up to rounding)
    a2, w2, _ = oracle.healpix_deposit(p0 - center, hsml, m, rho, q, w, 16)
    np.testing.assert_allclose(a, a2, rtol=1e-11, atol=1e-300)
    # Q11: calc_mean=false with a shrinking mask raises (BoundsError in the reference)
    q[3] = 0.0
    with pytest.raises(IndexError):
        oracle.healpix_map(p0.copy(), hsml, m, rho, q, w, center=center, nside=16, calc_mean=False)


def test_published_healpy_known_answers(oracle):
    """Known answers of the reference HEALPix implementation (healpy / HEALPix C++), as printed in healpy's own
    docstrings for ang2pix, pix2ang, pix2vec and vec2pix at Nside 16 (RING scheme; quoted from the documentation — no
    HEALPix library exists in this image).  Healpix.jl is validated against the same library, so these pin the oracle's
    RING arithmetic to the standard at these vectors; everything else about Healpix.jl stays unpinned."""
    L = oracle.lib()
    th = [math.pi / 2, math.pi / 4, math.pi / 2, 0.0, math.pi]
    ph = [0.0, math.pi / 4, math.pi / 2, 0.0, 0.0]
    assert [L.s2go_hp_ang2pix_ring(16, a, b) for a, b in zip(th, ph)] == [1440, 427, 1520, 0, 3068]
    t = C.c_double(); p = C.c_double()
    L.s2go_hp_pix2ang_ring(16, 1440, C.byref(t), C.byref(p))
    assert t.value == pytest.approx(1.5291175943723188, abs=1e-15) and p.value == 0.0
    want_t = [1.52911759, 0.78550497, 1.57079633, 0.05103658, 3.09055608]
    want_p = [0.0, 0.78539816, 1.61988371, 0.78539816, 0.78539816]
    for pix, wt, wp in zip([1440, 427, 1520, 0, 3068], want_t, want_p):
        L.s2go_hp_pix2ang_ring(16, pix, C.byref(t), C.byref(p))
        assert t.value == pytest.approx(wt, abs=5e-9) and p.value == pytest.approx(wp, abs=5e-9)
    v = np.zeros(3)
    vp = v.ctypes.data_as(C.POINTER(C.c_double))
    L.s2go_hp_pix2vec_ring(16, 1504, vp)
    assert v == pytest.approx([0.99879545620517241, 0.049067674327418015, 0.0], abs=1e-15)
    L.s2go_hp_pix2vec_ring(16, 1440, vp)
    assert v == pytest.approx([0.99913157, 0.0, 0.04166667], abs=5e-9)
    L.s2go_hp_pix2vec_ring(16, 427, vp)
    assert v == pytest.approx([0.5000534, 0.5000534, 0.70703125], abs=5e-9)


@pytest.mark.parametrize("kernel", ["Cubic", "WendlandC4", "WendlandC6"])
@pytest.mark.parametrize("nside", [8, 16])
def test_c_vs_python_mirror_healpix_deposit(oracle, kernel, nside):
    """The C oracle's particle loop of healpix_map against the second, independent restatement in
    oracle/numpy_mirror.py (brute-force discs, pixel centres from ring geometry, the reference's acos form of the
    angular distance): resolved discs, sub-pixel particles (the `distr_weight == 0` branch) and calc_mean=false."""
    rng = np.random.default_rng(11 + nside)
    n = 40
    pos = rng.normal(size=(n, 3))
    pos /= np.linalg.norm(pos, axis=1)[:, None]
    pos *= (0.5 + rng.random(n))[:, None]
    hsml = 0.02 + rng.random(n) * 0.25
    hsml[:5] = 1e-4                      # no pixel centre covered -> wk := 1 branch (pixel_weights.jl:120-133)
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 10; w = rng.random(n) + 0.5
    q[5:9] = 0.0
    L = oracle.lib()
    cp = []
    for p in range(n):
        d = np.linalg.norm(pos[p])
        cp.append(L.s2go_hp_ang2pix_ring(nside, math.acos(pos[p, 2] / d), math.atan2(pos[p, 1], pos[p, 0]) % (2 * math.pi)))
    for calc_mean in (True, False):
        a, wa, st = oracle.healpix_deposit(pos, hsml, m, rho, q, w, nside, kernel, 2, calc_mean)
        b, wb = nm.healpix_deposit(pos, hsml, m, rho, q, w, nside, kernel, 2, calc_mean, centre_pixels=cp)
        assert st["n_fallback"] >= 5
        for x, y in ((a, b), (wa, wb)):
            den = np.maximum(np.maximum(np.abs(x), np.abs(y)), 1e-300)
            # u = acos(p.c/|p|)/proj_h amplifies a last-ulp difference between the two pix2vec formulations
            # (sin(acos z) vs sqrt((1-z)(1+z))) by eps/dx^2 (DESIGN.md "conditioning"): 1e-12 typical, 2e-10 observed
            # for a particle 1e-3 rad from a pixel centre at Nside 16
            assert np.max(np.abs(x - y) / den) < 2e-9
            assert np.array_equal(x == 0, y == 0)      # same pixel sets


# ------------------------------------------------------------------------------------------------------------------
# Which Float64 evaluation is right?  The extended-precision arbiter (oracle/s2g_oracle_exact.c) settles it on the CPU.
# ------------------------------------------------------------------------------------------------------------------
def test_arbiter_long_double_agrees_with_float128(oracle):
    """long double chord form == __float128 literal acos(min(d/r,1)) to the resolution of long double unit vectors
    (5e-20 rad absolute), i.e. both are 'the exact angle' at the 1e-10 bar; the literal expression in Float64 is not."""
    rng = np.random.default_rng(17)
    L = oracle.lib()
    for nside, bound64 in ((32, 1e-11), (256, 1e-10), (2048, 1e-9)):
        worst_ld, worst_64 = 0.0, 0.0
        for _ in range(400):
            v = rng.normal(size=3); v *= rng.uniform(0.1, 3.0) / np.linalg.norm(v)
            th, ph = np.arccos(v[2] / np.linalg.norm(v)), np.arctan2(v[1], v[0]) % (2 * np.pi)
            pix = L.s2go_hp_ang2pix_ring(nside, th, ph)          # the pixel holding the particle: the smallest dx
            _, raw = oracle.hp_angdist_exact(nside, pix, v)
            c = np.zeros(3); L.s2go_hp_pix2vec_ring(nside, pix, c.ctypes.data_as(oracle._dp))
            f64 = np.arccos(min(np.dot(v, c) / np.sqrt(np.sum(v * v)), 1.0))   # distance_to_pixel_center, literally
            worst_ld = max(worst_ld, abs(float((raw[0] - raw[1]) / raw[1])))
            worst_64 = max(worst_64, abs(float((np.longdouble(f64) - raw[1]) / raw[1])))
        assert worst_ld < 1e-12, (nside, worst_ld)
        assert worst_64 > bound64, (nside, worst_64)   # the acos form loses eps/dx^2


@pytest.mark.parametrize("nside", [32, 256, 2048])
def test_conditioning_study(oracle, nside):
    """Sparse particles from 0.1 to 400 pixels across, WendlandC4.  Against the extended-precision maps:
      * the Float64 CHORD formulation (what csrc/s2g_healpix*.cu evaluate; here on the CPU) meets
        |x - exact| <= 1e-10 max(|x|,|exact|) + 4 ulp * sens on EVERY pixel, with no absolute floor;
      * the literal Float64 acos form (pixel_weights.jl:16-22 as written, oracle/s2g_oracle.c) misses that bar on
        thousands of pixels at Nside >= 256, by up to 1e-7 .. 1e-5 — it is the 1e-8 side of the old 5e-8 allowance."""
    from util import hp_violations
    rng = np.random.default_rng(1)
    n = 2000
    pos = rng.normal(size=(n, 3)); pos *= (rng.uniform(0.5, 2.0, n) / np.linalg.norm(pos, axis=1))[:, None]
    hs = 10 ** rng.uniform(-4, -1, n)
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 1e4
    a, w, st = oracle.healpix_deposit(pos, hs, m, rho, q, rho, nside, "WendlandC4")
    ea, ew, est = oracle.healpix_deposit(pos, hs, m, rho, q, rho, nside, "WendlandC4", exact="sens", n_workers=4)
    ca, cw = oracle.healpix_deposit_chord64(pos, hs, m, rho, q, rho, nside, "WendlandC4")
    for k in ("n_mapped", "footprint_pixels", "n_fallback"):
        assert st[k] == est[k]
    assert hp_violations(cw, ew, est["sens"], ulps=4.0)[0] == 0
    assert hp_violations(ca, ea, est["sens_q"], ulps=4.0)[0] == 0
    nbad, worst = hp_violations(w, ew, est["sens"], ulps=8.0)
    if nside >= 256:
        assert nbad > 100 and worst > 1e-8, (nbad, worst)
    # and the literal form is still "the same map" at its own conditioning level
    assert worst < 1e-4
