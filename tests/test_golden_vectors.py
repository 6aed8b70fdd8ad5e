"""Committed golden vectors (tests/golden/oracle_vectors.npz, made by make_oracle_vectors.py): the oracle must keep
reproducing them bit for bit (CPU), and the CUDA path must match them at the parity bar (GPU)."""
import os

import numpy as np
import pytest

from util import assert_healpix_parity, assert_parity
from test_stokes_rm import polarised_parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz"))
ARGS = [G[k] for k in ("pos", "hsml", "m", "rho", "q", "w")]


def test_oracle_reproduces_golden_vectors(oracle):
    pos, hsml, m, rho, q, w = ARGS
    Q = np.stack([q, np.sqrt(q + 1.0)], axis=1)
    assert np.array_equal(oracle.cic_mapping_2d(pos, hsml, m, rho, Q, w, 6.4, 64, "WendlandC6", 2, True)[0],
                          G["map2d_WendlandC6"])
    assert np.array_equal(oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, 6.4, 64, "Cubic", 2, False)[0],
                          G["map2d_Cubic_nomean"])
    assert np.array_equal(oracle.cic_mapping_3d(pos, hsml, m, rho, q, w, 2.0, 20, "WendlandC4", 3, False)[0],
                          G["map3d_WendlandC4"])
    a, wm, _ = oracle.healpix_deposit(G["hp_pos"], hsml * 12.0, m, rho, q, w, 16, "WendlandC4", 2, True)
    assert np.array_equal(a, G["hp_map"]) and np.array_equal(wm, G["hp_wmap"])
    ea, ew, est = oracle.healpix_deposit(G["hp_pos"], hsml * 12.0, m, rho, q, w, 16, "WendlandC4", 2, True,
                                         n_workers=1, exact="sens")
    # long double libm (sinl/cosl/asinl) and the azimuth recurrence of the ring walker: allow a drift at the level of
    # the extended type (1e-13 of a pixel value, 1e-17 of the map maximum for kernel-rim pixels)
    for x, y in ((ea, G["hp_map_exact"]), (ew, G["hp_wmap_exact"])):
        assert np.allclose(x, y, rtol=1e-13, atol=1e-17 * np.abs(y).max())
    assert np.array_equal(oracle.stencil_deposit(2, 3, pos, q, 2.0, 20, False), G["cic3d"])
    assert np.array_equal(oracle.stencil_deposit(3, 2, pos, q, 6.4, 64, True), G["tsc2d"])
    o = G["stokes_order"]
    st = oracle.cic_mapping_2d_rm(pos[o], hsml[o], m[o], rho[o], G["stokes_qu"][o], w[o], G["stokes_rm"][o], 6.4, 64,
                                  "WendlandC4", 2, True, True)[0]
    # atan / sin / cos come from libm: allow an ulp-level drift between libm builds instead of demanding identical bits
    polarised_parity(st, G["stokes_WendlandC4"], 2, 1e-13, "golden stokes (oracle)")


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", ["auto", "scatter", "gather"])
def test_gpu_matches_golden_vectors(s2g, strategy):
    pos, hsml, m, rho, q, w = ARGS
    ctx = s2g.Context(0, strategy=strategy)
    p2 = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=64)
    p3 = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=20)
    Q = np.stack([q, np.sqrt(q + 1.0)], axis=1)
    assert_parity(s2g.cic_mapping_2D(pos, hsml, m, rho, Q, w, param=p2, kernel=s2g.WendlandC6(2), calc_mean=True, ctx=ctx),
                  G["map2d_WendlandC6"], what="golden 2D WendlandC6 multi-image")
    assert_parity(s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=p2, kernel=s2g.Cubic(2), calc_mean=False, ctx=ctx),
                  G["map2d_Cubic_nomean"], what="golden 2D Cubic calc_mean=false")
    assert_parity(s2g.cic_mapping_3D(pos, hsml, m, rho, q, w, param=p3, kernel=s2g.WendlandC4(3), ctx=ctx),
                  G["map3d_WendlandC4"], what="golden 3D")
    a, wm = s2g.healpix_deposit(G["hp_pos"], hsml * 12.0, m, rho, q, w, 16, s2g.WendlandC4(2), True, ctx=ctx)
    # the yardstick is the committed extended-precision vector (util.assert_healpix_parity: 1e-10 + the Float64
    # unit-vector resolution term); the literal Float64 vector only has to agree to the acos conditioning of Nside 16
    assert_healpix_parity(a, wm, G["hp_map_exact"], G["hp_wmap_exact"], dict(sens=G["hp_sens"], sens_q=G["hp_sens_q"]),
                          what="golden healpix vs extended precision")
    assert_parity(a, G["hp_map"], rtol=1e-8, what="golden healpix map (literal Float64 acos form)")
    assert_parity(wm, G["hp_wmap"], rtol=1e-8, what="golden healpix weights (literal Float64 acos form)")
    if strategy == "auto":
        o = G["stokes_order"]
        st = s2g.cic_mapping_2D(pos[o], hsml[o], m[o], rho[o], G["stokes_qu"][o], w[o], G["stokes_rm"][o], param=p2,
                                kernel=s2g.WendlandC4(2), calc_mean=True, stokes=True, ctx=ctx)
        polarised_parity(st, G["stokes_WendlandC4"], 2, 1e-10, "golden stokes")
    assert_parity(s2g.cic_deposit(pos, q, param=p3, dimensions=3, average=False, ctx=ctx), G["cic3d"], rtol=1e-12,
                  what="golden CIC")
    assert_parity(s2g.tsc_deposit(pos, q, param=p2, dimensions=2, average=False, periodic=True, ctx=ctx), G["tsc2d"],
                  rtol=1e-12, what="golden TSC")
    ctx.close()
