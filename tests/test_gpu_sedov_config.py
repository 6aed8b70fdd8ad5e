"""BASELINE.json configs[0]: the reference's Sedov test (test/runtests.jl:182-275) at its exact mapping parameters —
WendlandC4(2), Npixels 256, center [3,3,3], size 5.4, boxsize 6, rho map (part_weight_physical, reduce_image=false)
and T map (weights=rho, reduce_image=true), both calc_mean=true (map_it).  The snapshot `snap_sedov` is downloaded
by the reference's test-suite and is NOT available offline, so the comparison against sedov_*_reference.fits cannot be
executed here; a synthetic Sedov-like particle set goes through the GPU path and the oracle instead.  If a Gadget
snapshot is ever placed at test_data/snap_sedov the literal comparison should be added (not executed = not claimed)."""
import numpy as np
import pytest

from util import assert_parity

pytestmark = pytest.mark.gpu


def sedov_like(n_side=40, box=6.0, seed=0):
    """Glass-like lattice in a periodic box with a Sedov-Taylor-like radial density/temperature profile."""
    rng = np.random.default_rng(seed)
    g = (np.arange(n_side) + 0.5) * box / n_side
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    pos += rng.normal(scale=0.15 * box / n_side, size=pos.shape)
    pos %= box
    r = np.linalg.norm(pos - box / 2, axis=1)
    rs = 1.6                                         # shock radius
    rho0 = 0.00247                                   # ambient density of the reference's test snapshot
    rho = rho0 * np.where(r < rs, 0.05 + 3.95 * (r / rs) ** 6, 1.0)
    m = np.full(len(r), rho0 * box ** 3 / len(r))
    hsml = np.cbrt(3 * 200 * m / (4 * np.pi * rho))  # 200 neighbours (WendlandC4)
    T = np.where(r < rs, 5.0 * (rs / np.maximum(r, 0.05)) ** 2 * 1e-2, 5e-9)
    return pos.astype(np.float32), hsml.astype(np.float32), m.astype(np.float32), rho.astype(np.float32), \
        T.astype(np.float32)


def test_sedov_config_rho_and_T_maps(s2g, oracle):
    pos, hsml, m, rho, T = sedov_like()
    kw = dict(center=[3.0, 3.0, 3.0], x_size=5.4, y_size=5.4, z_size=5.4, Npixels=256, boxsize=6.0)
    par, opar = s2g.mappingParameters(**kw), oracle.mapping_parameters(**kw)
    x_cgs = 3.085678e21
    rho_cgs = rho.astype(np.float64) * 6.77e-22
    # rho map: column density
    w = s2g.part_weight_physical(len(m), par, x_cgs)
    a = s2g.sphMapping(pos.copy(), hsml, m, rho, rho_cgs, w, param=par, kernel=s2g.WendlandC4(2), calc_mean=True,
                       reduce_image=False, show_progress=False)
    b = oracle.sph_mapping(pos.copy(), hsml, m, rho, rho_cgs, w, param=opar, kernel="WendlandC4", calc_mean=True,
                           reduce_image=False)
    assert a.shape == (256, 256, 1) and np.count_nonzero(a) == a.size
    assert_parity(a, b, what="Sedov-like rho map")
    # T map: density weighted mean
    a = s2g.sphMapping(pos.copy(), hsml, m, rho, T, rho, param=par, kernel=s2g.WendlandC4(2), calc_mean=True,
                       reduce_image=True, show_progress=False)
    b = oracle.sph_mapping(pos.copy(), hsml, m, rho, T, rho, param=opar, kernel="WendlandC4", calc_mean=True,
                           reduce_image=True)
    assert_parity(a, b, what="Sedov-like T map")
    assert a.min() > 0 and a.max() < 5.0 * (1.6 / 0.05) ** 2 * 1e-2 * 1.001


# ------------------------------------------------------------------------------------------------------------------
# The LITERAL configs[0] comparison (test/runtests.jl:175-275): runs whenever the snapshot is there.
# ------------------------------------------------------------------------------------------------------------------
def _find(name):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in (os.environ.get("S2G_TEST_DATA", ""), os.path.join(root, "test_data"), os.path.join(root, "tests", "golden"),
              "/root/reference/test"):
        if d and os.path.isfile(os.path.join(d, name)):
            return os.path.join(d, name)
    return None


@pytest.mark.parametrize("parallel", [False, True])
def test_literal_sedov_snapshot_against_reference_fits(s2g, parallel, tmp_path):
    """test/runtests.jl:175-275: read `snap_sedov` (Gadget format 2), map rho (column density, reduce_image=false) and
    T (rho-weighted mean) with WendlandC4(2) at 256^2 through map_it, compare with sedov_rho_reference.fits /
    sedov_T_reference.fits.  The snapshot is downloaded by the reference's test-suite (runtests.jl:5-6) and cannot be
    fetched offline: the test SKIPS unless `snap_sedov` is found in $S2G_TEST_DATA, test_data/ or tests/golden/ (the two
    reference images: the same places, or the reference checkout).  The unit factors come from GadgetUnits.jl
    (`GadgetPhysical(xH=0.752)`, third party, not vendored: unpinned), and every one of them scales a map uniformly, so
    the comparison is: image / reference is ONE constant over the map to the reference's own `≈` tolerance
    (rtol sqrt(eps) in the 2-norm, as `@test image ≈ rho_ref`), and that constant is 1 to the accuracy of the
    published unit constants (1e-3)."""
    snap = _find("snap_sedov")
    rho_fits, t_fits = _find("sedov_rho_reference.fits"), _find("sedov_T_reference.fits")
    if snap is None or rho_fits is None or t_fits is None:
        pytest.skip("snap_sedov / sedov_*_reference.fits not present (the reference downloads the snapshot at test time)")
    from sphtogrid_b200 import gadget, io
    rho_ref = io.read_fits_image(rho_fits)[0]
    t_ref = io.read_fits_image(t_fits)[0]
    h = gadget.read_header(snap)
    data = {b: gadget.read_block(snap, b, parttype=0) for b in ("POS", "MASS", "HSML", "RHO", "U")}
    # GadgetPhysical(xH=0.752) with hpar = 1, a_scale = 1: x/rho/m_physical = 1
    x_cgs, m_cgs = 3.085678e21, 1.989e43
    rho_cgs = m_cgs / x_cgs ** 3
    xH, gamma, mp, kB = 0.752, 5.0 / 3.0, 1.6726219e-24, 1.380649e-16
    T_K = (gamma - 1.0) * 1e10 * mp * (4.0 / (5.0 * xH + 3.0)) / kB
    pos, hsml, rho, mass = data["POS"], data["HSML"], data["RHO"], data["MASS"]
    rho_gcm3 = data["RHO"].astype(np.float64) * rho_cgs
    T = data["U"].astype(np.float64) * T_K
    center = np.ones(3) * 0.5 * h.boxsize
    size = 0.9 * h.boxsize
    par = s2g.mappingParameters(center=center, x_size=size, y_size=size, z_size=size, Npixels=256, boxsize=h.boxsize)
    k = s2g.WendlandC4(2)
    w = s2g.part_weight_physical(len(hsml), par, x_cgs)
    pre = str(tmp_path / "sedov")
    a = s2g.map_it(pos, hsml, mass, rho, rho_gcm3, w, kernel=k, units="g/cm^2", param=par, reduce_image=False,
                   parallel=parallel, snap=50, image_prefix=pre + "_rho", show_progress=False)
    b = s2g.map_it(pos, hsml, mass, rho, T, rho, kernel=k, units="K", param=par, reduce_image=True, parallel=parallel,
                   snap=50, image_prefix=pre + "_T", show_progress=False)
    for got, ref, what in ((io.read_fits_image(pre + "_rho.xy.fits")[0], rho_ref, "rho"),
                           (io.read_fits_image(pre + "_T.xy.fits")[0], t_ref, "T")):
        got = np.asarray(got, dtype=np.float64).reshape(ref.shape)
        scale = float(np.vdot(ref, got) / np.vdot(ref, ref))          # least-squares unit factor
        assert abs(scale - 1.0) < 1e-3, f"{what}: unit factor {scale}"
        assert np.linalg.norm(got - scale * ref) <= 1.5e-8 * max(np.linalg.norm(got), np.linalg.norm(ref)), what
    assert a.shape[:2] == (256, 256) and b.shape[:2] == (256, 256)
