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
