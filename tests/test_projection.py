"""map_it's projection pre-step (cic_interpolation.jl:331-345; rotate_particles.jl, rotate_parameters.jl): host mirrors
against the oracle restatement, and the fused device path (s2g_sphmap_projected: the permutation / rotation applied
inside the position load) against the oracle run on explicitly rotated copies."""
import numpy as np
import pytest

from util import assert_parity, random_particles


def _same_par(a, b):
    for k in ("x_lim", "y_lim", "z_lim", "center", "halfsize"):
        assert np.array_equal(np.asarray(getattr(a, k)), np.asarray(getattr(b, k))), k
    assert a.len2pix == b.len2pix and int(a.Npixels[0]) == int(b.Npixels[0]) and a.boxsize == b.boxsize


def test_axis_rotations_host_vs_oracle(oracle):
    import __graft_entry__ as ge
    s2g = ge.load_package()
    rng = np.random.default_rng(0)
    for dt in (np.float64, np.float32):
        x = rng.normal(size=(50, 3)).astype(dt)
        a = s2g.rotate_to_xz_plane(x.copy()); b = oracle.rotate_to_xz_plane(x.copy())
        assert a.dtype == dt and np.array_equal(a, b) and np.array_equal(a, x[:, [0, 2, 1]])
        a = s2g.rotate_to_yz_plane(x.copy()); b = oracle.rotate_to_yz_plane(x.copy())
        assert a.dtype == dt and np.array_equal(a, b) and np.array_equal(a, x[:, [1, 2, 0]])
        out = np.empty_like(x)
        assert np.array_equal(s2g.rotate_to_xz_plane(out, x), x[:, [0, 2, 1]])
        assert s2g.project_along_axis(x, 3) is x
        assert np.array_equal(s2g.project_along_axis(x.copy(), 2), x[:, [0, 2, 1]])
        assert np.array_equal(s2g.project_along_axis(x.copy(), 1), x[:, [1, 2, 0]])
    p = s2g.mappingParameters(center=[1.0, 2.0, 3.0], x_size=2.0, y_size=4.0, z_size=6.0, Npixels=16, boxsize=20.0)
    o = oracle.mapping_parameters(center=[1.0, 2.0, 3.0], x_size=2.0, y_size=4.0, z_size=6.0, Npixels=16, boxsize=20.0)
    _same_par(s2g.rotate_to_xz_plane(p), oracle.rotate_parameters_xz(o))
    _same_par(s2g.rotate_to_yz_plane(p), oracle.rotate_parameters_yz(o))
    assert s2g.rotate_to_xz_plane(p).halfsize.tolist() == [1.0, 3.0, 2.0]


def test_euler_rotation_host_vs_oracle(oracle):
    import __graft_entry__ as ge
    s2g = ge.load_package()
    rng = np.random.default_rng(1)
    x = rng.normal(size=(200, 3))
    for ang in ((90.0, 0.0, 0.0), (0.0, 90.0, 0.0), (0.0, 0.0, 90.0), (12.5, -40.0, 77.0)):
        a = s2g.rotate_3D(x, *ang); b = oracle.rotate_3d(x, *ang)
        assert np.abs(a - b).max() < 4e-16 * np.abs(x).max() * 3
        assert np.allclose(np.linalg.norm(a, axis=1), np.linalg.norm(x, axis=1), rtol=1e-14)
    # known answers: Rx(90): (x,y,z) -> (x,-z,y); Ry(90): (z,y,-x); Rz(90): (-y,x,z)
    v = np.array([[1.0, 2.0, 3.0]])
    assert np.allclose(s2g.rotate_3D(v, 90, 0, 0), [[1, -3, 2]], atol=1e-15)
    assert np.allclose(s2g.rotate_3D(v, 0, 90, 0), [[3, 2, -1]], atol=1e-15)
    assert np.allclose(s2g.rotate_3D(v, 0, 0, 90), [[-2, 1, 3]], atol=1e-15)
    r = s2g.euler_matrix(12.5, -40.0, 77.0)
    assert np.allclose(r @ r.T, np.eye(3), atol=1e-15) and np.isclose(np.linalg.det(r), 1.0)
    assert s2g.rotate_3D(x.astype(np.float32), 10, 20, 30).dtype == np.float64  # rot * x promotes
    y = x.copy(); s2g.rotate_3D_(y, 10, 20, 30); assert np.array_equal(y, x)    # `rotate_3D!` does not mutate


def test_reference_rotate_particles_testset():
    """The reference's own "Rotate particles" testset (test/runtests.jl:101-137), literally: positions are Julia's
    Matrix(3, N) there, (N, 3) rows here."""
    import __graft_entry__ as ge
    s2g = ge.load_package()
    x_rand = np.random.default_rng(3).random((10, 3))
    x_out = s2g.rotate_3D(x_rand, 0.0, 0.0, 0.0)                   # matrix, no rotation (:104-106)
    assert np.allclose(x_out, x_rand, rtol=1.5e-8, atol=0)
    s2g.rotate_3D_(x_out, 0.0, 0.0, 0.0)                           # "inplace" (:109-110)
    assert np.allclose(x_out, x_rand, rtol=1.5e-8, atol=0)
    x_in = np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0]])            # two particles (1,1,0) (:113-115)
    assert np.array_equal(s2g.project_along_axis(x_in.copy(), 3), x_in)                      # along z (:118-120)
    assert np.array_equal(s2g.project_along_axis(x_in.copy(), 2), [[1.0, 0.0, 1.0]] * 2)     # along y (:123-130)
    # the reference's "along x-axis" case calls axis 2 again (:133-136): same expectation
    assert np.array_equal(s2g.project_along_axis(x_in.copy(), 2), [[1.0, 0.0, 1.0]] * 2)
    assert np.array_equal(s2g.project_along_axis(x_in.copy(), 1), [[1.0, 0.0, 1.0]] * 2)     # yz: (y, z, x)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("projection", ["xy", "xz", "yz", (20.0, -35.0, 60.0)])
def test_map_it_projection_parity(s2g, oracle, tmp_path, dtype, projection):
    pos, hsml, m, rho, q, w = random_particles(21, 20000, box=12.0, hmax=0.8, dtype=dtype, center=3.0)
    kw = dict(center=[3.0, 2.5, 3.5], x_size=8.0, y_size=6.0, z_size=4.0, Npixels=128, boxsize=40.0)
    par = s2g.mappingParameters(**kw)
    opar = oracle.mapping_parameters(**kw)
    keep = pos.copy()
    got = s2g.map_it(pos, hsml, m, rho, q, w, param=par, kernel=s2g.WendlandC4(2), reduce_image=True, parallel=False,
                     calc_mean=True, show_progress=False, projection=projection,
                     image_prefix=str(tmp_path / "img"))
    assert np.array_equal(pos, keep)  # map_it works on a copy
    ref = oracle.map_it(pos, hsml, m, rho, q, w, param=opar, kernel="WendlandC4", reduce_image=True, calc_mean=True,
                        projection=projection)
    # Euler angles: product and oracle build the rotated positions with differently ordered roundings (and neither is
    # pinned against Rotations.jl), i.e. the INPUTS of the deposit differ by an ulp -> 1e-9 instead of 1e-10
    assert_parity(got, ref, 1e-10 if isinstance(projection, str) else 1e-9, f"map_it {projection}")
    tag = projection if isinstance(projection, str) else "alpha=20.00beta=-35.00gamma=60.00"
    img, rpar, snap, units = s2g.read_fits_image(str(tmp_path / f"img.{tag}.fits"))
    assert np.array_equal(img[:, :, 0] if img.ndim == 3 else img, got[:, :, 0])
    if isinstance(projection, str) and projection != "xy":
        # the fused permutation is exact; two runs only differ by the order of the atomic adds
        rot = {"xz": s2g.rotate_to_xz_plane, "yz": s2g.rotate_to_yz_plane}[projection]
        again = s2g.sphMapping(rot(pos.copy()), hsml, m, rho, q, w, param=rot(par), kernel=s2g.WendlandC4(2),
                               reduce_image=True, calc_mean=True, show_progress=False)
        assert_parity(got, again, 1e-12, "fused vs explicit permutation")


@pytest.mark.gpu
def test_projection_on_unfused_paths(s2g, oracle, tmp_path):
    """sort_z (host filter, Q5) cannot take the fused projection: the host rotates the copy instead; same map."""
    pos, hsml, m, rho, q, w = random_particles(22, 5000, box=12.0, hmax=0.8)
    kw = dict(center=[0.0, 0.5, -0.5], x_size=8.0, y_size=8.0, z_size=6.0, Npixels=64)
    par = s2g.mappingParameters(**kw); opar = oracle.mapping_parameters(**kw)
    for projection in ("yz", (10.0, 20.0, 30.0)):
        got = s2g.map_it(pos, hsml, m, rho, q, w, param=par, kernel=s2g.Cubic(2), parallel=False, sort_z=True,
                         show_progress=False, projection=projection, write_fits=False)
        ref = oracle.map_it(pos, hsml, m, rho, q, w, param=opar, kernel="Cubic", projection=projection, sort_z=True)
        assert_parity(got, ref, 1e-9, f"sort_z {projection}")
    with pytest.raises(ValueError):
        s2g.map_it(pos, hsml, m, rho, q, w, param=par, projection="zz", show_progress=False, write_fits=False)
