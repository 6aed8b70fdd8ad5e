import numpy as np

KERNELS = ["Cubic", "Quintic", "WendlandC2", "WendlandC4", "WendlandC6", "WendlandC8"]


def kern(s2g, name, dim=2):
    return getattr(s2g, name)(dim)


def rel_err(a, b, floor=0.0):
    """max over elements of |a-b| / max(|a|,|b|,floor)"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(den > 0, np.abs(a - b) / den, 0.0)
    return float(np.max(r)) if r.size else 0.0


def assert_parity(got, ref, rtol=1e-10, what=""):
    """FP64-mode bar of BASELINE.json's north_star: per-element relative 1e-10.  The absolute floor
    (1e-14 of the largest magnitude of the plane) only forgives pixels that receive nothing but kernel-edge
    contributions ~(1-u)^8 -> 0, where a last-ulp difference in u is amplified without bound."""
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), what + ": NaN pattern differs"
    floor = 1e-4 * float(np.nanmax(np.abs(ref))) if ref.size else 0.0
    e = rel_err(np.nan_to_num(got), np.nan_to_num(ref), floor=floor * 1e-10)
    assert e <= rtol, f"{what}: max rel err {e:.3e} > {rtol:g}"
    return e


def random_particles(seed, n, box=10.0, hmin=0.01, hmax=1.0, dtype=np.float64, center=0.0):
    rng = np.random.default_rng(seed)
    pos = ((rng.random((n, 3)) - 0.5) * box + center).astype(dtype)
    hsml = (hmin + rng.random(n) ** 2 * (hmax - hmin)).astype(dtype)
    m = (rng.random(n) + 0.1).astype(dtype)
    rho = (rng.random(n) + 0.1).astype(dtype)
    q = (rng.random(n) * 1e3).astype(dtype)
    w = (rng.random(n) + 0.5).astype(dtype)
    return pos, hsml, m, rho, q, w


# ---------------------------------------------------------------------------------------------------------------
# HEALPix parity against the EXTENDED-PRECISION arbiter (oracle/s2g_oracle_exact.c).
# The reference's dx = acos(min(p·c/Δx, 1)) loses ε/dx² in Float64 (1e-8 .. 1e-6 at the pixel scale of Nside 256 .. 2048),
# so the literal Float64 oracle is NOT the yardstick for the maps (it still is for every integer: pixel sets,
# counters).  The yardstick is the same formulas in long double; the bar per pixel is
#     |got - exact| <= 1e-10 * max(|got|, |exact|)  +  HP_ULPS * eps * sens[pix]
# where sens = Σ |∂pix_weight/∂dx| over the contributions to the pixel (returned by the arbiter) and HP_ULPS * eps is
# the resolution of Float64 unit vectors: the second term is what NO Float64 evaluation can resolve (kernel-rim
# contributions (1-u)^k -> 0 have unbounded relative sensitivity).  No global absolute floor.
# tests/test_oracle_healpix.py::test_conditioning_study shows on the CPU that the Float64 chord formulation (what the
# CUDA kernels evaluate) meets this bar with 4 ulps at Nside 32/256/2048 while the literal acos form misses it by 1e-5.
# ---------------------------------------------------------------------------------------------------------------
HP_ULPS = 8.0
EPS = 2.220446049250313e-16


def hp_violations(got, exact, sens, rtol=1e-10, ulps=HP_ULPS):
    """(number of pixels over the bar, worst excess relative to the pixel value)"""
    got = np.asarray(got); exact = np.asarray(exact)
    d = np.abs(got - exact)
    den = np.maximum(np.abs(got), np.abs(exact))
    allow = rtol * den + ulps * EPS * sens
    bad = d > allow
    worst = float(np.max((d - ulps * EPS * sens) / np.where(den > 0, den, 1.0))) if d.size else 0.0
    return int(bad.sum()), worst


def assert_healpix_parity(got_a, got_w, exact_a, exact_w, stats, what="", rtol=1e-10, ulps=HP_ULPS):
    assert np.array_equal(np.isnan(got_a), np.isnan(exact_a)) and np.array_equal(np.isnan(got_w), np.isnan(exact_w)), \
        what + ": NaN pattern differs"
    nb, worst = hp_violations(np.nan_to_num(got_w), np.nan_to_num(exact_w), stats["sens"], rtol, ulps)
    assert nb == 0, f"{what}: weight map, {nb} pixels over the bar, worst excess {worst:.3e}"
    nb, worst_q = hp_violations(np.nan_to_num(got_a), np.nan_to_num(exact_a), stats["sens_q"], rtol, ulps)
    assert nb == 0, f"{what}: quantity map, {nb} pixels over the bar, worst excess {worst_q:.3e}"
    return max(worst, worst_q)


def ncores():
    import os
    return max(1, os.cpu_count() or 1)
