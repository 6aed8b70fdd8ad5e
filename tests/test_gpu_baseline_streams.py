"""GPU-vs-oracle parity ON THE BASELINE STREAMS (BASELINE.json configs[1..4], the synthetic sets bench.py times):
the first particles of each stream, FULL image size, same kernels.  Exact counters, per-pixel relative error at the
north-star bar (1e-10).  bench.py prints the same comparison as the `parity` object of its JSON line.

  C2  : first 2^17 particles of the 16 Mi stream -> 4096^2, WendlandC6, calc_mean T map
  C5s : first 2^17 particles of the 1 Gi stream  -> 8192^2, WendlandC6
  C3  : first 2^20 particles of the 64 Mi stream -> 512^3, Cubic
  C4  : the first 2^16 particles of the 128 Mi stream that lie in the shell [0.05, 0.5] L -> Nside 2048, WendlandC4,
        against the extended-precision arbiter (util.assert_healpix_parity)
"""
import os
import sys

import numpy as np
import pytest

from util import assert_healpix_parity, assert_parity, ncores

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _stream(name, count):
    import bench
    wl = bench.WORKLOADS[name]
    pos, hsml, m, rho, temp = bench.host_particles(wl, count)
    return wl, np.array(pos), hsml, m, rho, temp


@pytest.mark.parametrize("name,count", [("c2", 1 << 17), ("c5", 1 << 17)])
def test_2d_baseline_stream_subsample(s2g, oracle, name, count):
    wl, pos, hsml, m, rho, temp = _stream(name, count)
    kw = dict(center=[0.5, 0.5, 0.5], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=wl["npix"])
    p1, p2 = pos.copy(), pos.copy()
    got, st = s2g.sphMapping(p1, hsml, m, rho, temp, rho, param=s2g.mappingParameters(**kw),
                             kernel=getattr(s2g, wl["kernel"])(2), calc_mean=True, reduce_image=True,
                             show_progress=False, return_stats=True)
    opar = oracle.mapping_parameters(**kw)
    p2c, par_c = oracle.center_particles(p2, opar)
    # one private image + two scratch planes per worker (like the reference's workers): bound the host memory
    flat, ost = oracle.cic_mapping_2d(p2c, hsml, m, rho, temp, rho, par_c.len2pix, wl["npix"], wl["kernel"], 2, True,
                                      n_workers=min(ncores(), 8))
    ref = oracle.reduce_image_2d(flat, wl["npix"], wl["npix"], True)
    assert np.array_equal(p1, p2c)
    for k in ("n_mapped", "footprint_pixels", "touched_pixels", "n_fallback"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    assert st["n_mapped"] == count
    assert_parity(got, ref, what=f"{name} stream, first {count} particles, {wl['npix']}^2")


def test_3d_baseline_stream_subsample(s2g, oracle):
    count = 1 << 20
    wl, pos, hsml, m, rho, temp = _stream("c3", count)
    kw = dict(center=[0.5, 0.5, 0.5], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=wl["npix"])
    one = np.ones(count)
    p1, p2 = pos.copy(), pos.copy()
    got, st = s2g.sphMapping(p1, hsml, m, rho, rho, one, param=s2g.mappingParameters(**kw), kernel=s2g.Cubic(3),
                             dimensions=3, reduce_image=True, show_progress=False, return_stats=True)
    opar = oracle.mapping_parameters(**kw)
    p2c, par_c = oracle.center_particles(p2, opar)
    flat, ost = oracle.cic_mapping_3d(p2c, hsml, m, rho, rho, one, par_c.len2pix, wl["npix"], "Cubic", 3, False,
                                      n_workers=min(ncores(), 8))   # 4.3 GB of image + scratch per worker
    ref = oracle.reduce_image_3d(flat, wl["npix"], True)
    for k in ("n_mapped", "footprint_pixels", "touched_pixels", "n_fallback"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    assert_parity(got, ref, what="c3 stream, first 2^20 particles, 512^3")


def shell_particles(count, oversample=2):
    """The first `count` particles of the C4 stream inside the shell [0.05, 0.5] around the box centre, recentred."""
    wl, pos, hsml, m, rho, temp = _stream("c4", count * oversample)
    pos = pos - 0.5
    r = np.sqrt(pos[:, 0] ** 2 + pos[:, 1] ** 2 + pos[:, 2] ** 2)
    sel = np.flatnonzero((r >= 0.05) & (r <= 0.5))[:count]
    assert sel.size == count
    return wl, np.ascontiguousarray(pos[sel]), hsml[sel], m[sel], rho[sel], temp[sel]


def test_healpix_baseline_stream_nside_2048(s2g, oracle):
    count = 1 << 16
    wl, pos, hsml, m, rho, temp = shell_particles(count)
    nside = wl["npix"]
    a, wm, st = s2g.healpix_deposit(pos, hsml, m, rho, temp, rho, nside, s2g.WendlandC4(2), True, return_stats=True)
    ea, ew, est = oracle.healpix_deposit(pos, hsml, m, rho, temp, rho, nside, "WendlandC4", 2, True,
                                         n_workers=ncores(), exact="sens")
    # counters: the arbiter takes its pixel lists from the literal Float64 query_disc, so these are the reference's
    for k in ("n_mapped", "touched_pixels", "n_fallback"):
        assert st[k] == est[k], (k, st[k], est[k])
    assert st["n_mapped"] == count and st["touched_pixels"] > 500 * count
    assert np.array_equal(wm > 0, ew > 0)
    assert_healpix_parity(a, wm, ea, ew, est, what="c4 stream, 2^16 shell particles, Nside 2048")
    # the same through healpix_map (device filter: every particle is in the shell, so no far-to-near selection quirk)
    a2, w2 = s2g.healpix_map(pos.copy(), hsml, m, rho, temp, rho, center=[0.0, 0.0, 0.0], radius_limits=[0.05, 0.5],
                             Nside=nside, kernel=s2g.WendlandC4(2), show_progress=False)
    assert_healpix_parity(a2, w2, ea, ew, est, what="c4 stream through healpix_map")


def test_sharded_path_float32_positions_float64_fields(s2g, oracle):
    """sphMapping(parallel=True) over ranks with Float32 Pos and Float64 Bin_Quant (the common Gadget
    case) used to upload 12n bytes of positions and read them as 24n.  World size 1 goes through the same
    sph_mapping_sharded code."""
    rng = np.random.default_rng(5)
    n = 20000
    pos = (rng.random((n, 3)) * 6.0).astype(np.float32)
    hsml = (0.02 + rng.random(n) * 0.3).astype(np.float32)
    m = (rng.random(n) + 0.5).astype(np.float32)
    rho = (rng.random(n) + 0.5).astype(np.float32)
    q = rng.random(n) * 1e-30                       # Float64 on purpose (cgs-scale values that Float32 flushes)
    kw = dict(center=[3.0, 3.0, 3.0], x_size=5.0, y_size=5.0, z_size=5.0, Npixels=128, boxsize=6.0)
    p1, p2, p3 = pos.copy(), pos.copy(), pos.copy()
    got = s2g.sphMapping(p1, hsml, m, rho, q, rho, param=s2g.mappingParameters(**kw), kernel=s2g.WendlandC6(2),
                         calc_mean=True, parallel=True, show_progress=False)
    ser = s2g.sphMapping(p2, hsml, m, rho, q, rho, param=s2g.mappingParameters(**kw), kernel=s2g.WendlandC6(2),
                         calc_mean=True, parallel=False, show_progress=False)
    ref = oracle.sph_mapping(p3, hsml, m, rho, q, rho, param=oracle.mapping_parameters(**kw), kernel="WendlandC6",
                             calc_mean=True)
    assert np.array_equal(p1, p2) and np.array_equal(p1, p3) and p1.dtype == np.float32
    assert got.max() > 0
    assert_parity(got, ser, rtol=1e-12, what="sharded == serial")
    assert_parity(got, ref, what="sharded, Float32 Pos + Float64 Bin_Quant vs oracle")
    with pytest.raises(TypeError):
        from sphtogrid_b200 import distributed
        par = s2g.mappingParameters(**kw)
        distributed.sph_mapping_sharded(s2g.default_context(), pos.astype(np.float64), hsml, m.astype(np.float64),
                                        rho.astype(np.float64), q, rho.astype(np.float64), 1, 1, par,
                                        s2g.recentred_parameters(par), 4, 2, True, True, False)
