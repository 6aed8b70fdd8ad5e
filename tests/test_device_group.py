"""Device group (s2g_group_*): `parallel=true` of sphMapping (cic_interpolation.jl:171-215, 236-271) and healpix_map
driven by ONE process over several GPUs — host thread per device, particle slices of `domain_decomposition`
(parallel/domain_decomp.jl:7-17), partial images summed over peer memory fused with `reduce_image`.

A device may be listed more than once, so the sharding, the peer-memory sum (direct loads and the staged-copy variant)
and the epilogues are all exercised on a single GPU with `DeviceGroup([0, 0, 0])`; with two or more GPUs visible the
same checks run across real peers."""
import math

import numpy as np
import pytest

from util import assert_healpix_parity, assert_parity, ncores, random_particles


# ------------------------------------------------------------------ CPU: host logic of the shard map
@pytest.mark.parametrize("n,parts", [(0, 1), (0, 3), (1, 4), (10, 3), (16, 4), (17, 8), (1000003, 7), (2 ** 31 + 5, 8)])
def test_domain_decomposition_c_abi_matches_reference_formula(s2g, oracle, n, parts):
    """s2g_domain_decomposition needs no device; same slices as the oracle's restatement of domain_decomp.jl:7-17
    and as the host mirror."""
    from sphtogrid_b200 import _lib
    starts, counts = _lib.domain_decomposition_c(n, parts)
    ref = s2g.domain_decomposition(n, parts)
    assert [(s, s + c) for s, c in zip(starts, counts)] == ref
    assert sum(counts) == n and starts[0] == 0
    size = n // parts
    assert all(c == size for c in counts[:-1]) and counts[-1] == n - size * (parts - 1)
    if n < 10 ** 7:
        o = oracle.domain_decomposition(n, parts)
        assert [(int(a), int(b)) for a, b in o] == ref


def test_domain_decomposition_rejects_bad_arguments(s2g):
    from sphtogrid_b200 import _lib
    with pytest.raises(s2g.S2GError):
        _lib.domain_decomposition_c(-1, 2)
    with pytest.raises(s2g.S2GError):
        _lib.domain_decomposition_c(5, 0)


def test_group_init_fails_loudly_without_a_device(s2g):
    if s2g.lib().s2g_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(s2g.S2GError) as e:
        s2g.DeviceGroup([0, 0])
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_group_init_rejects_bad_lists(s2g):
    from sphtogrid_b200 import _lib
    import ctypes as C
    h = C.c_void_p()
    assert s2g.lib().s2g_group_init(None, 2, C.byref(h)) == _lib.S2G_EINVAL
    arr = (C.c_int32 * 1)(0)
    assert s2g.lib().s2g_group_init(arr, 0, C.byref(h)) == _lib.S2G_EINVAL
    assert s2g.lib().s2g_group_init(arr, 17, C.byref(h)) == _lib.S2G_EINVAL
    assert s2g.lib().s2g_group_size(None) == 0


# ------------------------------------------------------------------ GPU
def _device_lists(s2g):
    lists = [[0], [0, 0], [0, 0, 0]]
    nd = s2g.lib().s2g_device_count()
    if nd >= 2:
        lists.append(list(range(nd)))
    return lists


PAR = dict(center=[3.0, 3.0, 3.0], x_size=5.4, y_size=5.4, z_size=5.4, boxsize=6.0)


@pytest.mark.gpu
@pytest.mark.parametrize("no_p2p", [False, True])
def test_group_sphmapping_2d_equals_serial_and_oracle(s2g, oracle, no_p2p, monkeypatch):
    if no_p2p:
        monkeypatch.setenv("S2G_GROUP_NO_P2P", "1")
    pos, hsml, m, rho, q, w = random_particles(99, 20001, box=6.5, hmax=0.5, center=3.0)
    par = s2g.mappingParameters(Npixels=200, **PAR)
    kern = s2g.WendlandC6(2)
    p0 = pos.copy()
    serial = s2g.sphMapping(p0, hsml, m, rho, q, w, param=par, kernel=kern, calc_mean=True, show_progress=False)
    ref = oracle.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=oracle.mapping_parameters(Npixels=200, **PAR),
                             kernel="WendlandC6", calc_mean=True)
    for devs in _device_lists(s2g):
        grp = s2g.DeviceGroup(devs)
        assert len(grp) == len(devs)
        assert grp.peer_access == (not no_p2p)
        p1 = pos.copy()
        got, stats = s2g.sphMapping(p1, hsml, m, rho, q, w, param=par, kernel=kern, calc_mean=True, parallel=True,
                                    group=grp, show_progress=False, return_stats=True)
        assert np.array_equal(p1, p0), "Pos must be recentred in place exactly like the serial call"
        assert len(stats) == len(devs)
        assert sum(s["n_in"] for s in stats) == pos.shape[0]
        assert [s["n_in"] for s in stats] == [b - a for a, b in s2g.domain_decomposition(pos.shape[0], len(devs))]
        # (not bit-equal even for one device: the scatter kernel's red.add order differs from run to run)
        assert_parity(got, serial, rtol=1e-12, what=f"group {devs} vs serial")
        assert_parity(got, ref, what=f"group {devs} vs oracle")
        grp.close()


@pytest.mark.gpu
def test_group_sphmapping_both_maps_multi_image_and_no_reduce(s2g, oracle):
    pos, hsml, m, rho, q, w = random_particles(5, 6007, box=6.5, hmax=0.7, center=3.0)
    q3 = np.stack([q, 2.0 * q + 1.0, np.sqrt(q)], axis=1)          # (N, N_images)
    par = s2g.mappingParameters(Npixels=97, **PAR)                  # odd size: ragged 32x32 tiles and pixel slices
    kern = s2g.WendlandC4(2)
    grp = s2g.DeviceGroup([0, 0, 0])
    for kw in (dict(return_both_maps=True), dict(reduce_image=False), dict(reduce_image=True)):
        serial = s2g.sphMapping(pos.copy(), hsml, m, rho, q3, w, param=par, kernel=kern, calc_mean=True,
                                show_progress=False, **kw)
        got = s2g.sphMapping(pos.copy(), hsml, m, rho, q3, w, param=par, kernel=kern, calc_mean=True, parallel=True,
                             group=grp, show_progress=False, **kw)
        assert got.shape == serial.shape
        assert_parity(got, serial, rtol=1e-12, what=f"group vs serial {kw}")
    ref = oracle.sph_mapping(pos.copy(), hsml, m, rho, q3, w, param=oracle.mapping_parameters(Npixels=97, **PAR),
                             kernel="WendlandC4", calc_mean=True)
    assert_parity(got, ref, what="group multi-image vs oracle")


@pytest.mark.gpu
def test_group_sphmapping_3d_and_float32(s2g, oracle):
    pos, hsml, m, rho, q, w = random_particles(17, 5003, box=6.5, hmax=0.6, center=3.0)
    par = s2g.mappingParameters(Npixels=45, **PAR)
    grp = s2g.DeviceGroup([0, 0, 0])
    for reduce_image in (True, False):
        serial = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=par, kernel=s2g.Cubic(3), dimensions=3,
                                reduce_image=reduce_image, show_progress=False)
        got = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=par, kernel=s2g.Cubic(3), dimensions=3,
                             reduce_image=reduce_image, parallel=True, group=grp, show_progress=False)
        assert_parity(got, serial, rtol=1e-12, what=f"3D group vs serial reduce_image={reduce_image}")
    ref = oracle.sph_mapping(pos.copy(), hsml, m, rho, q, w, param=oracle.mapping_parameters(Npixels=45, **PAR),
                             kernel="Cubic", dimensions=3, reduce_image=False)
    assert_parity(got, ref, what="3D group vs oracle")
    # Float32 inputs: recentred in Float32 on every device (Q2)
    f = lambda a: a.astype(np.float32)
    par2 = s2g.mappingParameters(Npixels=64, **PAR)
    p32a, p32b = f(pos), f(pos)
    serial = s2g.sphMapping(p32a, f(hsml), f(m), f(rho), f(q), f(w), param=par2, kernel=s2g.WendlandC6(2),
                            calc_mean=True, show_progress=False)
    got = s2g.sphMapping(p32b, f(hsml), f(m), f(rho), f(q), f(w), param=par2, kernel=s2g.WendlandC6(2), calc_mean=True,
                         parallel=True, group=grp, show_progress=False)
    assert p32b.dtype == np.float32 and np.array_equal(p32a, p32b)
    assert_parity(got, serial, rtol=1e-12, what="Float32 group vs serial")
    # Float32 positions with Float64 fields: recentred on the host side of the call first, then no further shift
    p32c, p32d = f(pos), f(pos)
    serial = s2g.sphMapping(p32c, hsml, m, rho, q, w, param=par2, kernel=s2g.WendlandC6(2), calc_mean=True,
                            show_progress=False)
    got = s2g.sphMapping(p32d, hsml, m, rho, q, w, param=par2, kernel=s2g.WendlandC6(2), calc_mean=True, parallel=True,
                         group=grp, show_progress=False)
    assert np.array_equal(p32c, p32d)
    assert_parity(got, serial, rtol=1e-12, what="mixed-precision group vs serial")


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 2, 5])
def test_group_fewer_particles_than_devices(s2g, n):
    """domain_decomposition with N < workers: all slices but the last are empty."""
    pos, hsml, m, rho, q, w = random_particles(3, max(n, 1), box=4.0, hmax=0.5, center=3.0)
    pos, hsml, m, rho, q, w = pos[:n], hsml[:n], m[:n], rho[:n], q[:n], w[:n]
    par = s2g.mappingParameters(Npixels=32, **PAR)
    grp = s2g.DeviceGroup([0, 0, 0, 0])
    for dims, kern in ((2, s2g.WendlandC6(2)), (3, s2g.Cubic(3))):
        serial = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=par, kernel=kern, dimensions=dims,
                                show_progress=False)
        got = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=par, kernel=kern, dimensions=dims, parallel=True,
                             group=grp, show_progress=False)
        assert_parity(got, serial, rtol=1e-12, what=f"n={n} dims={dims}")


@pytest.mark.gpu
def test_group_strategy_is_per_device_context(s2g):
    pos, hsml, m, rho, q, w = random_particles(8, 3000, box=6.0, hmax=0.9, center=3.0)
    par = s2g.mappingParameters(Npixels=128, **PAR)
    out = {}
    for strat in ("scatter", "gather"):
        grp = s2g.DeviceGroup([0, 0], strategy=strat, exact_norm=True)
        out[strat], st = s2g.sphMapping(pos.copy(), hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC6(2),
                                        calc_mean=True, parallel=True, group=grp, show_progress=False,
                                        return_stats=True)
        if strat == "scatter":
            assert all(s["n_gather"] == 0 for s in st)
        else:
            assert all(s["n_gather"] > 0 for s in st)
    assert_parity(out["scatter"], out["gather"], what="scatter vs gather through the group")


@pytest.mark.gpu
def test_group_healpix_map_selection_is_global(s2g, oracle):
    """healpix_map over a group: `sorted[sel]` (filter_particles.jl:33-41) depends on the radii of ALL particles, so
    the group makes the selection once over the whole input — same maps as the single call and the oracle, whether or
    not the shell contains every particle."""
    rng = np.random.default_rng(12)
    n = 4001
    center = np.array([10.0, -5.0, 3.0])
    pos = rng.normal(size=(n, 3)) * 40.0 + center
    pos[::7] = pos[1::7][: len(pos[::7])]          # exact ties in the radii (stable-sort order matters)
    hsml = rng.random(n) * 3.0 + 0.5
    m = rng.random(n) + 0.5; rho = rng.random(n) + 0.5; q = rng.random(n) * 10 + 1; w = rng.random(n) + 0.5
    for devs in _device_lists(s2g)[1:]:
        grp = s2g.DeviceGroup(devs)
        for rl in ([20.0, 60.0], [0.0, 45.0], [0.0, np.inf]):
            p1, p2, p3 = pos.copy(), pos.copy(), pos.copy()
            a, wm = s2g.healpix_map(p1, hsml, m, rho, q, w, center=center, radius_limits=rl, Nside=64,
                                    kernel=s2g.WendlandC4(2), show_progress=False, group=grp)
            sa, sw = s2g.healpix_map(p2, hsml, m, rho, q, w, center=center, radius_limits=rl, Nside=64,
                                     kernel=s2g.WendlandC4(2), show_progress=False)
            ra, rw = oracle.healpix_map(p3, hsml, m, rho, q, w, center=center, radius_limits=rl, nside=64,
                                        kernel="WendlandC4")
            assert np.array_equal(p1, p2) and np.array_equal(p1, p3)
            assert np.array_equal(wm > 0, sw > 0), "same set of touched pixels as the single call"
            assert_parity(wm, sw, rtol=1e-11, what=f"group healpix weights vs single, shell {rl}")
            assert_parity(a, sa, rtol=1e-11, what=f"group healpix map vs single, shell {rl}")
            ea, ew, est = oracle.healpix_map(pos.copy(), hsml, m, rho, q, w, center=center, radius_limits=rl, nside=64,
                                             kernel="WendlandC4", exact="sens", n_workers=ncores())
            assert_healpix_parity(a, wm, ea, ew, est, what=f"group healpix vs extended precision, shell {rl}")
            assert math.isclose(wm.sum(), rw.sum(), rel_tol=1e-12)
        grp.close()


@pytest.mark.gpu
def test_group_errors_name_the_device(s2g):
    grp = s2g.DeviceGroup([0, 0])
    pos, hsml, m, rho, q, w = random_particles(1, 100, box=4.0, center=3.0)
    par = s2g.mappingParameters(Npixels=32, **PAR)

    bad = s2g.AbstractSPHKernel(dim=2, kernel_id=99, name="bad")
    with pytest.raises(s2g.S2GError) as e:
        s2g.sphMapping(pos, hsml, m, rho, q, w, param=par, kernel=bad, parallel=True, group=grp, show_progress=False)
    assert "device 0 of the group" in str(e.value) and "unknown kernel id" in str(e.value)
    grp.close()
    with pytest.raises(s2g.S2GError):
        grp.handle
