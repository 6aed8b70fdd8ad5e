"""Optional FP32-accumulate mode (BASELINE.json north_star: "1e-5 in an optional FP32-accumulate mode").

`Context.set_accumulate_mode("f32")` switches the per-pixel chain of the 2D tile-gather kernel (kernel evaluation and
the partial sums of each 256-particle batch) to single precision; footprints, pass A, the scatter kernel and everything
outside the 2D Smac deposit stay FP64.  Bar, written here: per pixel |gpu - oracle| <= 1e-5 * max(|gpu|, |oracle|)
+ 1e-9 * max|plane| (the absolute term forgives pixels fed only by kernel-rim contributions (1-u)^k -> 0, whose
relative error in single precision is unbounded)."""
import numpy as np
import pytest

from util import KERNELS, kern, random_particles


def _f32_err(got, ref):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    worst = 0.0
    planes = got.reshape(-1, got.shape[-1]) if got.ndim > 1 else got.reshape(-1, 1)
    refs = ref.reshape(planes.shape)
    for k in range(planes.shape[1]):
        a, b = planes[:, k], refs[:, k]
        floor = 1e-9 * float(np.max(np.abs(b))) if b.size else 0.0
        den = 1e-5 * np.maximum(np.abs(a), np.abs(b)) + floor
        with np.errstate(invalid="ignore", divide="ignore"):
            r = np.where(den > 0, np.abs(a - b) / den, 0.0)
        worst = max(worst, float(np.max(r)) if r.size else 0.0)
    return worst   # <= 1 passes


def test_accumulate_mode_rejects_unknown_values(s2g):
    """no device needed: the setter validates before touching the context"""
    from sphtogrid_b200 import _lib
    assert s2g.lib().s2g_set_accumulate_mode(None, 1) == _lib.S2G_EINVAL
    assert set(_lib.ACCUMULATE) == {"f64", "f32"}


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", KERNELS)
def test_fp32_accumulate_flat_image_within_1e5(s2g, oracle, kernel):
    pos, hsml, m, rho, q, w = random_particles(21, 5000, box=11.0, hmin=0.3, hmax=2.5)
    npix = 256
    len2pix = npix / 10.0
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
    ref, _, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, len2pix, npix, kernel, 2, True, want_footprints=True)
    ctx = s2g.Context(0, strategy="gather")
    exact, st64 = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel), calc_mean=True,
                                     ctx=ctx, return_stats=True)
    ctx.set_accumulate_mode("f32")
    got, st32 = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel), calc_mean=True,
                                   ctx=ctx, return_stats=True)
    assert st32["n_gather"] > 0 and st32["n_gather"] == st64["n_gather"]
    assert st32["footprint_pixels"] == st64["footprint_pixels"], "footprints are FP64/integer work in both modes"
    e = _f32_err(got, ref)
    assert e <= 1.0, f"{kernel}: FP32-accumulate error {e:.3f} x the 1e-5 bar"
    # the mode really ran in single precision (differs from the FP64 result) ...
    assert not np.array_equal(got, exact)
    # ... and switching back restores the FP64 result to the FP64 bar
    ctx.set_accumulate_mode("f64")
    again = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kernel), calc_mean=True, ctx=ctx)
    from util import assert_parity
    assert_parity(again, ref, what="back in FP64 mode")
    # totals: the plane sum stays within half the per-pixel bar (rounding errors mostly cancel)
    assert abs(got[:, 1].sum() - ref[:, 1].sum()) <= 5e-6 * abs(ref[:, 1].sum())
    ctx.close()


@pytest.mark.gpu
def test_fp32_accumulate_through_sphmapping_multi_image_and_clipping(s2g, oracle):
    """Through the public entry, with clipped footprints, several quantities, the mean (q/w) epilogue and long
    per-tile particle lists (several 256-record batches and 4096-pair chunks per tile)."""
    pos, hsml, m, rho, q, w = random_particles(22, 40000, box=7.0, hmin=0.4, hmax=1.2, center=3.0)
    q3 = np.stack([q, 3.0 * q + 2.0, np.sqrt(q)], axis=1)
    kw = dict(center=[3.0, 3.0, 3.0], x_size=5.4, y_size=5.4, z_size=5.4, Npixels=192, boxsize=6.0)
    par = s2g.mappingParameters(**kw)
    opar = oracle.mapping_parameters(**kw)
    ctx = s2g.Context(0)
    ctx.set_accumulate_mode("f32")
    for extra in (dict(return_both_maps=True), dict(reduce_image=True), dict(reduce_image=False)):
        got = s2g.sphMapping(pos.copy(), hsml, m, rho, q3, w, param=par, kernel=s2g.WendlandC6(2), calc_mean=True,
                             show_progress=False, ctx=ctx, **extra)
        ref = oracle.sph_mapping(pos.copy(), hsml, m, rho, q3, w, param=opar, kernel="WendlandC6", calc_mean=True,
                                 **extra)
        e = _f32_err(got, ref)
        assert e <= 1.0, f"{extra}: FP32-accumulate error {e:.3f} x the 1e-5 bar"
    ctx.close()


@pytest.mark.gpu
def test_fp32_mode_leaves_scatter_3d_and_healpix_in_fp64(s2g, oracle):
    from util import assert_parity
    ctx = s2g.Context(0, strategy="scatter")
    ctx.set_accumulate_mode("f32")
    pos, hsml, m, rho, q, w = random_particles(23, 3000, box=11.0, hmax=1.0)
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=128)
    got = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC4(2), calc_mean=True, ctx=ctx)
    ref, _, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, 12.8, 128, "WendlandC4", 2, True, want_footprints=True)
    assert_parity(got, ref, what="scatter kernel is FP64 in either mode")
    par3 = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=40)
    g3 = s2g.cic_mapping_3D(pos, hsml, m, rho, q, w, param=par3, kernel=s2g.Cubic(3), ctx=ctx)
    r3 = oracle.cic_mapping_3d(pos, hsml, m, rho, q, w, 4.0, 40, "Cubic", 3)
    r3 = r3[0] if isinstance(r3, tuple) else r3
    assert_parity(g3, r3, what="3D deposit is FP64 in either mode")
    ctx.close()
