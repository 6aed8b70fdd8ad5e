"""Pins the CPU oracle against the reference's own self-contained known-answer tests
(/root/reference/test/runtests.jl, line numbers cited per test) and against an independent
pure-Python restatement.  CPU only."""
import math

import numpy as np
import pytest

from oracle import numpy_mirror as nm


# ---- test/runtests.jl:38-58
def test_mapping_parameters_errors(oracle):
    with pytest.raises(Exception, match="Giving a center position requires extent in x, y and z direction."):
        oracle.mapping_parameters()
    with pytest.raises(Exception, match="Please specify pixelSideLength or number of pixels!"):
        oracle.mapping_parameters(center=[0.0, 0.0, 0.0], x_lim=[-1.0, 1.0], y_lim=[-1.0, 1.0], z_lim=[-1.0, 1.0])
    p = oracle.mapping_parameters(center=[0.0, 0.0, 0.0], x_lim=[-1.0, 1.0], y_lim=[-1.0, 1.0], z_lim=[-1.0, 1.0],
                                  Npixels=100)
    assert p.Npixels.tolist() == [100, 100, 100] and p.pixelSideLength == 2.0 / 100 and not p.periodic
    p = oracle.mapping_parameters(center=[0.0, 0.0, 0.0], x_lim=[-1.0, 1.0], y_lim=[-1.0, 1.0], z_lim=[-1.0, 1.0],
                                  pixelSideLength=0.2)
    assert p.Npixels.tolist() == [10, 10, 10] and p.len2pix == 1.0 / (2.0 / 10)


# ---- test/runtests.jl:60-77
def test_filter_particles(oracle):
    par = oracle.mapping_parameters(center=[3.0, 3.0, 3.0], x_size=6.0, y_size=6.0, z_size=6.0, Npixels=500)
    x = np.array([[1.0, 1.0, 1.0], [-5.0, -1.0, 1.0]])
    mask = oracle.filter_particles_in_image(x, par)
    assert mask.tolist() == [True, False]


# ---- test/runtests.jl:79-99
def test_shift_particles(oracle):
    par = oracle.mapping_parameters(center=[1.0, 1.0, 1.0], x_size=6.0, y_size=6.0, z_size=6.0, Npixels=500)
    x = np.array([[1.0, 1.0, 1.0], [-3.0, -1.0, 1.0]])
    x2, par2 = oracle.center_particles(x, par)
    assert np.allclose(x2[0], [0, 0, 0]) and np.allclose(x2[1], [-4.0, -2.0, 0.0])
    assert x2 is x  # in place (Q1)
    assert par.center.tolist() == [1.0, 1.0, 1.0] and par2.center.tolist() == [0.0, 0.0, 0.0]


def test_center_particles_f32_and_periodic_quirk(oracle):
    # Q2: arithmetic in Float32 storage; Q3: wrap by boxsize/2
    par = oracle.mapping_parameters(center=[3.0, 3.0, 3.0], x_size=6.0, y_size=6.0, z_size=6.0, Npixels=16, boxsize=6.0)
    x = np.array([[0.1, 5.9, 3.0], [6.5, -0.7, 2.9]], dtype=np.float32)
    ref = x.astype(np.float64) - 3.0
    ref = ref.astype(np.float32)
    big = np.abs(ref.astype(np.float64)) > 3.0
    wrapped = np.where(ref > 0, ref.astype(np.float64) - 3.0, ref.astype(np.float64) + 3.0).astype(np.float32)
    ref = np.where(big, wrapped, ref)
    oracle.center_particles(x, par)
    assert x.dtype == np.float32 and np.array_equal(x, ref)


# ---- test/runtests.jl:139-166
def test_index_bijection(oracle):
    L = oracle.lib()
    N = 128
    seen = np.zeros(N * N, dtype=np.int64)
    count = 1
    for i in range(N):
        for j in range(N):
            seen[L.s2go_calculate_index_2d(i, j, N)] = count
            count += 1
    assert np.array_equal(seen, np.arange(1, N * N + 1))
    N3 = 24
    seen = np.zeros(N3 ** 3, dtype=np.int64)
    count = 1
    for i in range(N3):
        for j in range(N3):
            for k in range(N3):
                seen[L.s2go_calculate_index_3d(i, j, k, N3, N3)] = count
                count += 1
    assert np.array_equal(seen, np.arange(1, N3 ** 3 + 1))
    # the full N=128 3D enumeration, vectorised with the same formula
    i, j, k = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    idx = (i * N * N + j * N + k).ravel()
    assert np.array_equal(idx, np.arange(N ** 3))
    assert L.s2go_calculate_index_3d(127, 126, 125, N, N) == 127 * N * N + 126 * N + 125


# ---- test/runtests.jl:718-738 ("Particle not overlapping with centers in 2D")
def test_kat_2d_fallback_mass_conservation(oracle):
    npix, r = 4, 10
    param = oracle.mapping_parameters(x_lim=[-r, r], y_lim=[-r, r], z_lim=[-r, r], Npixels=npix)
    pos = np.array([[0.5, 0.0, 0.0]])
    hsml = np.array([1.0]); mass = np.array([1.0]); rho = np.ones(1)
    w = np.ones(1) * param.pixelSideLength * 1  # part_weight_physical(N, param, 1)
    mp = oracle.sph_mapping(pos, hsml, mass, rho, rho, w, param=param, kernel="Cubic", kernel_dim=3,
                            reduce_image=False)
    assert mp.shape == (4, 4, 1)
    Apix = (param.x_lim[1] - param.x_lim[0]) ** 2 / npix ** 2
    assert math.isclose(Apix * mp.sum(), 1.0, rel_tol=1e-10)
    # worked example: len2pix 0.2, h 0.2 px, x=2.1, y=2.0 -> footprint i in {1,2}... wait x-h=1.9
    img, fp, st = oracle.cic_mapping_2d(np.array([[0.5, 0.0, 0.0]]), hsml, mass, rho, rho, w, param.len2pix, npix,
                                        "Cubic", 3, False, want_footprints=True)
    assert fp[0].tolist() == [1, 2, 1, 2] and st["n_fallback"] == 1 and st["touched_pixels"] == 4
    assert math.isclose(img[:, 0].sum(), 0.04, rel_tol=1e-12)


# ---- test/runtests.jl:324-342 ("Mass conservation", 3D)
def test_kat_3d_mass_conservation(oracle):
    npix, r = 200, 64
    param = oracle.mapping_parameters(x_lim=[-r, r], y_lim=[-r, r], z_lim=[-r, r], Npixels=npix)
    pos = np.array([[0.0101, -0.001, 0.001]])
    hsml = np.array([5.0]); mass = np.array([3.0]); rho = np.ones(1)
    w = np.ones(1) * param.pixelSideLength * 1
    mp = oracle.sph_mapping(pos, hsml, mass, rho, rho, w, param=param, kernel="Cubic", kernel_dim=3, dimensions=3,
                            reduce_image=False)
    Vpix = (param.x_lim[1] - param.x_lim[0]) ** 3 / npix ** 3
    assert math.isclose(Vpix * mp.sum(), 3.0, rel_tol=1e-8)  # Julia's `≈` is rtol sqrt(eps)


# ---- kernels: shapes vs independent mirror, norms by quadrature (∫ W dV = 1)
@pytest.mark.parametrize("kernel", ["Cubic", "Quintic", "WendlandC2", "WendlandC4", "WendlandC6", "WendlandC8"])
def test_kernel_values_and_norms(oracle, kernel):
    from scipy import integrate
    for dim in (2, 3):
        for u in np.linspace(0, 1.05, 43):
            a = oracle.kernel_value(kernel, dim, u, 0.7)
            b = nm.W(kernel, dim, u, 0.7)
            assert a == pytest.approx(b, rel=1e-13, abs=1e-300)
        f = (lambda u: oracle.kernel_value(kernel, 2, u, 1.0) * 2 * math.pi * u) if dim == 2 else \
            (lambda u: oracle.kernel_value(kernel, 3, u, 1.0) * 4 * math.pi * u * u)
        val, _ = integrate.quad(f, 0, 1, epsabs=1e-12, epsrel=1e-12, points=[1 / 3, 0.5, 2 / 3])
        assert val == pytest.approx(1.0, rel=1e-9)
    assert oracle.kernel_value(kernel, 2, 1.0, 1.0) == 0.0


# ---- C oracle vs the independent pure-Python restatement
def _random_set(rng, n, box=10.0, hmax=2.0):
    pos = (rng.random((n, 3)) - 0.5) * box
    hsml = rng.random(n) * hmax + 1e-3
    m = rng.random(n) + 0.1
    rho = rng.random(n) + 0.1
    q = rng.random(n) * 1e3
    w = rng.random(n) + 0.5
    return pos, hsml, m, rho, q, w


@pytest.mark.parametrize("kernel", ["Cubic", "WendlandC4", "WendlandC6", "Quintic", "WendlandC8", "WendlandC2"])
def test_c_vs_python_mirror_2d(oracle, kernel):
    rng = np.random.default_rng(11)
    pos, hsml, m, rho, q, w = _random_set(rng, 60)
    hsml[:10] *= 0.02  # exercise the fallback branch
    q[5] = 0.0
    npix, len2pix = 24, 24 / 10.0
    for calc_mean in (True, False):
        img, st = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, len2pix, npix, kernel, 2, calc_mean)
        ref = nm.cic_mapping_2d(pos, hsml, m, rho, q, w, len2pix, npix, kernel, 2, calc_mean)
        assert st["n_fallback"] > 0
        np.testing.assert_allclose(img, ref, rtol=1e-12, atol=0)


def test_c_vs_python_mirror_2d_multi_image(oracle):
    rng = np.random.default_rng(12)
    pos, hsml, m, rho, q, w = _random_set(rng, 40)
    Q = np.stack([q, rng.random(40), np.zeros(40)], axis=1)
    Q[3, :] = 0.0
    img, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, Q, w, 2.0, 20, "WendlandC4", 2, True)
    ref = nm.cic_mapping_2d(pos, hsml, m, rho, Q, w, 2.0, 20, "WendlandC4", 2, True)
    assert img.shape == (400, 4)
    np.testing.assert_allclose(img, ref, rtol=1e-12, atol=0)


@pytest.mark.parametrize("kernel", ["Cubic", "WendlandC6"])
def test_c_vs_python_mirror_3d(oracle, kernel):
    rng = np.random.default_rng(13)
    pos, hsml, m, rho, q, w = _random_set(rng, 25, hmax=1.5)
    hsml[:5] *= 0.05
    img, st = oracle.cic_mapping_3d(pos, hsml, m, rho, q, w, 1.2, 12, kernel, 3, False)
    ref = nm.cic_mapping_3d(pos, hsml, m, rho, q, w, 1.2, 12, kernel, 3, False)
    np.testing.assert_allclose(img, ref, rtol=1e-12, atol=0)
    # grid-mass diagnostic of cic_3D.jl:186: equals particle mass only for unclipped, non-fallback particles
    assert st["particle_mass"] == pytest.approx(m.sum())


def test_parallel_slices_equal_serial(oracle):
    rng = np.random.default_rng(14)
    pos, hsml, m, rho, q, w = _random_set(rng, 501)
    a, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, 3.2, 32, "WendlandC6", 2, True)
    b, _ = oracle.cic_mapping_2d(pos, hsml, m, rho, q, w, 3.2, 32, "WendlandC6", 2, True, n_workers=3)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-300)
    assert oracle.domain_decomposition(10, 3) == [(0, 3), (3, 6), (6, 10)]  # parallel/domain_decomp.jl:7-17


def test_reduce_image_layouts(oracle):
    n = 5
    flat = np.zeros((n * n, 3), order="F")
    flat[:, 0] = np.arange(n * n) + 1.0
    flat[:, 1] = -(np.arange(n * n) + 1.0)
    flat[:, 2] = 2.0
    flat[7, 2] = 0.0
    out = oracle.reduce_image_2d(flat, n, n, True)
    for ix in range(n):
        for iy in range(n):
            k = ix * n + iy
            wgt = flat[k, 2]
            assert out[ix, iy, 0] == (flat[k, 0] / wgt if wgt > 0 else flat[k, 0])
            assert out[ix, iy, 1] == (flat[k, 1] / wgt if wgt > 0 else flat[k, 1])
    out = oracle.reduce_image_2d(flat, n, n, False)
    assert out[2, 3, 0] == flat[2 * n + 3, 0]
    f3 = np.zeros((n ** 3, 2), order="F")
    f3[:, 0] = np.arange(n ** 3) - 10.0
    f3[:, 1] = 4.0
    o3 = oracle.reduce_image_3d(f3, n, True)
    for (ix, iy, iz) in [(0, 0, 0), (1, 2, 3), (4, 4, 4), (2, 0, 1)]:
        mm = ix * n * n + iy * n + iz
        v = f3[mm, 0]
        assert o3[iz, iy, ix] == (v / 4.0 if v > 0 else v)  # Q7: gate on the quantity plane
    o3 = oracle.reduce_image_3d(f3, n, False)
    assert o3[3, 2, 1] == f3[1 * n * n + 2 * n + 3, 0]


@pytest.mark.parametrize("dims", [2, 3])
def test_c_vs_python_mirror_edge_biased_fuzz(oracle, dims):
    """Differential test of the two restatements on edge-biased random cases: particles snapped to pixel edges, centres
    and image borders (± 1e-15), hsml from 1e-6 pixels to several images, zero quantities and weights, 1-pixel images."""
    from oracle import numpy_mirror as nm
    rng = np.random.default_rng(1000 + dims)
    kernels = ["Cubic", "Quintic", "WendlandC2", "WendlandC4", "WendlandC6", "WendlandC8"]
    for it in range(60 if dims == 2 else 36):
        n = int(rng.integers(1, 10))
        npix = int(rng.choice([1, 2, 3, 4, 7, 16] if dims == 2 else [1, 2, 3, 5]))
        box = 10.0
        len2pix = npix / box
        pos = (rng.random((n, 3)) - 0.5) * box * rng.choice([0.5, 1.0, 1.3])
        for p in range(n):
            if rng.random() < 0.4:
                pos[p, rng.integers(0, dims)] = (rng.integers(0, npix + 1) / len2pix - box / 2) + \
                    rng.choice([0.0, 1e-15, -1e-15, 0.5 / len2pix])
        hs = rng.choice([1e-6, 0.01, 0.3, 1.0, 3.0, 20.0], size=n) * rng.random(n)
        m = rng.random(n) + 0.1; rho = rng.random(n) + 0.1; q = rng.random(n) * 10; w = rng.random(n) + 0.5
        if rng.random() < 0.3:
            q[rng.integers(0, n)] = 0.0
        if rng.random() < 0.2:
            w[rng.integers(0, n)] = 0.0
        k = kernels[it % 6]
        for calc_mean in (True, False):
            if dims == 2:
                a = oracle.cic_mapping_2d(pos, hs, m, rho, q, w, len2pix, npix, k, 2, calc_mean)[0]
                b = nm.cic_mapping_2d(pos, hs, m, rho, q, w, len2pix, npix, k, 2, calc_mean)
            else:
                a = oracle.cic_mapping_3d(pos, hs, m, rho, q, w, len2pix, npix, k, 3, calc_mean)
                a = a[0] if isinstance(a, tuple) else a
                b = nm.cic_mapping_3d(pos, hs, m, rho, q, w, len2pix, npix, k, 3, calc_mean)
            a = np.asarray(a).reshape(b.shape)
            assert np.array_equal(np.isnan(a), np.isnan(b))
            den = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
            assert np.max(np.nan_to_num(np.abs(a - b) / den)) < 1e-11, (it, k, calc_mean, npix)
