"""CPU-only tests: host mirror of the reference interface, the C-ABI library's export table, and the N>1 host logic
(world_size-2 gloo).  No compute call into the CUDA library happens here."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(s2g):
    hdr = open(os.path.join(ROOT, "include", "sphtogrid_cuda.h")).read()
    declared = set(re.findall(r"S2G_API\s+[\w\s\*]+?\b(s2g_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    L = s2g.lib()
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in include/sphtogrid_cuda.h but not exported"
    assert declared == set(s2g.EXPORTED_SYMBOLS), declared ^ set(s2g.EXPORTED_SYMBOLS)
    assert b"sm_100a" in L.s2g_version()


def test_no_cpu_fallback(s2g):
    if s2g.lib().s2g_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(s2g.S2GError, match="no CUDA device"):
        s2g.Context(0)
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=8)
    one = np.ones(4)
    with pytest.raises(s2g.S2GError):
        s2g.sphMapping(np.zeros((4, 3)), one, one, one, one, param=par, kernel=s2g.Cubic(), show_progress=False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sphtogrid.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt.lower(), \
                    f"{f} mentions the oracle: the product path must never route through it"


# ---- test/runtests.jl:38-58 through the product's mappingParameters
def test_mapping_parameters_reference_cases(s2g):
    with pytest.raises(ValueError, match="Giving a center position requires extent in x, y and z direction."):
        s2g.mappingParameters()
    with pytest.raises(ValueError, match="Please specify pixelSideLength or number of pixels!"):
        s2g.mappingParameters(center=[0.0, 0.0, 0.0], x_lim=[-1.0, 1.0], y_lim=[-1.0, 1.0], z_lim=[-1.0, 1.0])
    s2g.mappingParameters(center=[0.0, 0.0, 0.0], x_lim=[-1.0, 1.0], y_lim=[-1.0, 1.0], z_lim=[-1.0, 1.0], Npixels=100)
    p = s2g.mappingParameters(center=[0.0, 0.0, 0.0], x_lim=[-1.0, 1.0], y_lim=[-1.0, 1.0], z_lim=[-1.0, 1.0],
                              pixelSideLength=0.2)
    assert p.Npixels.tolist() == [10, 10, 10]


def test_mapping_parameters_match_oracle_bitwise(s2g, oracle):
    rng = np.random.default_rng(0)
    for _ in range(300):
        c = rng.normal(size=3) * 100
        sz = rng.random(3) * 50 + 0.1
        n = int(rng.integers(1, 5000))
        box = float(rng.choice([-1.0, 100.0]))
        if rng.random() < 0.5:
            kw = dict(center=c.tolist(), x_size=sz[0], y_size=sz[1], z_size=sz[2], Npixels=n, boxsize=box)
        else:
            kw = dict(x_lim=[c[0] - sz[0], c[0] + sz[0]], y_lim=[c[1] - sz[1], c[1] + sz[1]],
                      z_lim=[c[2] - sz[2], c[2] + sz[2]], pixelSideLength=float(sz.max() / n), boxsize=box)
        a = s2g.mappingParameters(**kw)
        b = oracle.mapping_parameters(**kw)
        for f in ("x_lim", "y_lim", "z_lim", "center", "halfsize", "Npixels"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
        assert a.len2pix == b.len2pix and a.pixelSideLength == b.pixelSideLength and a.periodic == b.periodic
        a2 = s2g.recentred_parameters(a)
        _, b2 = oracle.center_particles(np.zeros((1, 3)), b)
        assert a2.len2pix == b2.len2pix and np.array_equal(a2.halfsize, b2.halfsize)
        assert a2.center.tolist() == [0.0, 0.0, 0.0]


# ---- test/runtests.jl:488-506
def test_weight_functions(s2g):
    assert s2g.part_weight_one(1)[0] == 1.0
    par = s2g.mappingParameters(center=[3.0, 3.0, 3.0], x_size=6.0, y_size=6.0, z_size=6.0, Npixels=200, boxsize=6.0)
    assert np.isclose(s2g.part_weight_physical(1, par)[0], par.pixelSideLength * 3.085678e21)
    assert np.isclose(s2g.part_weight_physical(1)[0], 3.085678e21)
    assert np.isclose(s2g.part_weight_emission([0.5, 0.5], [0.5, 0.5])[0], 0.1767766952966369)
    assert np.isclose(s2g.part_weight_spectroscopic([0.5, 0.5], [0.5, 0.5])[0], 0.4204482076268573)


def test_kernel_types(s2g):
    assert s2g.Cubic().dim == 3 and s2g.WendlandC6(2).dim == 2 and s2g.WendlandC4(float, 2).dim == 2
    ids = [k().kernel_id for k in (s2g.Cubic, s2g.Quintic, s2g.WendlandC2, s2g.WendlandC4, s2g.WendlandC6,
                                   s2g.WendlandC8)]
    assert ids == [0, 1, 2, 3, 4, 5]
    with pytest.raises(ValueError):
        s2g.Cubic(4)


def test_domain_decomposition_matches_reference(s2g, oracle):
    for n, w in [(10, 3), (7, 7), (5, 8), (1000003, 8), (0, 2)]:
        assert s2g.domain_decomposition(n, w) == oracle.domain_decomposition(n, w)


def test_filter_sort_particles_host_logic(s2g, oracle):
    rng = np.random.default_rng(2)
    n = 200
    pos = rng.normal(size=(n, 3)) * 30
    args = [rng.random(n) for _ in range(5)]
    c = [1.0, 2.0, 3.0]
    a = s2g.filter_sort_particles(pos.copy(), *args, c, [5.0, 60.0], True)
    b = oracle.filter_sort_particles(pos.copy(), *args, c, [5.0, 60.0], True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


# ---- N>1 host logic over gloo, world_size 2 (the deposit itself is stood in for by the oracle: CPU box)
def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as ge
    from oracle import oracle as orc
    s2g = ge.load_package()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(77)
    n = 1001
    pos = (rng.random((n, 3)) - 0.5) * 10; hs = rng.random(n) * 0.8 + 0.01
    m = rng.random(n) + 0.1; rho = rng.random(n) + 0.1; qq = rng.random(n); w = rng.random(n) + 0.5
    s, e = s2g.distributed.shard_range(n, world, rank)
    part, _ = orc.cic_mapping_2d(pos[s:e], hs[s:e], m[s:e], rho[s:e], qq[s:e], w[s:e], 6.4, 64, "WendlandC6", 2, True)
    if rank == 1:
        part[5, 0] = np.nan  # a poisoned partial map entry is skipped by the reference's accumulate (cic.jl:63-69)
    total = s2g.distributed.combine_partial_images(part, finite_guard=True)
    if rank == 0:
        q.put((s, e, total))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_shard_and_reduce(s2g, oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    s, e, total = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(77)
    n = 1001
    pos = (rng.random((n, 3)) - 0.5) * 10; hs = rng.random(n) * 0.8 + 0.01
    m = rng.random(n) + 0.1; rho = rng.random(n) + 0.1; qq = rng.random(n); w = rng.random(n) + 0.5
    assert (s, e) == (0, 500)
    a, _ = oracle.cic_mapping_2d(pos[:500], hs[:500], m[:500], rho[:500], qq[:500], w[:500], 6.4, 64, "WendlandC6", 2,
                                 True)
    b, _ = oracle.cic_mapping_2d(pos[500:], hs[500:], m[500:], rho[500:], qq[500:], w[500:], 6.4, 64, "WendlandC6", 2,
                                 True)
    b[5, 0] = 0.0
    assert np.array_equal(total, a + b)


# ---- the reduce-scatter exchange (exchange_reduce) and footprint-balanced shards over gloo, world_size 2
def _worker_exchange(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from oracle import oracle as orc
    s2g = ge.load_package()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(78)
    n = 1200
    pos = (rng.random((n, 3)) - 0.5) * 10; hs = rng.random(n) * 0.8 + 0.01
    hs[:300] *= 3.0                                   # clustered work: the first quarter carries most of the footprint
    m = rng.random(n) + 0.1; rho = rng.random(n) + 0.1; w = rng.random(n) + 0.5
    Q = np.stack([rng.random(n), rng.random(n) * 5.0], axis=1)
    fp = np.prod(np.diff(orc.cic_mapping_2d(pos, hs, m, rho, Q, w, 6.4, 64, "WendlandC6", 2, True, True)[1]
                         .reshape(n, 2, 2), axis=2)[..., 0] + 1, axis=1).clip(min=0)
    s, e = s2g.distributed.footprint_balanced_decomposition(fp, world)[rank]
    part, _ = orc.cic_mapping_2d(pos[s:e], hs[s:e], m[s:e], rho[s:e], Q[s:e], w[s:e], 6.4, 64, "WendlandC6", 2, True)

    def divide(qs, ws_, nn, stride, nim, dims, red):   # CPU stand-in of s2g_divide_slice_dev (plumbing under test)
        for k in range(nim):
            v = qs[k * stride:k * stride + nn]
            if red:
                v[ws_[:nn] > 0] /= ws_[:nn][ws_[:nn] > 0]

    flat = torch.from_numpy(np.ascontiguousarray(part.ravel(order="F")))
    out_all = s2g.distributed.exchange_reduce(flat.clone(), 2, 64 * 64, 2, True, divide, gather="all")
    out_root = s2g.distributed.exchange_reduce(flat.clone(), 2, 64 * 64, 2, True, divide, gather="root")
    assert (out_root is None) == (rank != 0)
    q.put((rank, s, e, float(fp[s:e].sum()), out_all.numpy().copy(), None if out_root is None else out_root.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_reduce_scatter_exchange_and_footprint_shards(s2g, oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_exchange, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, s0, e0, f0, all0, root0), (_, s1, e1, f1, all1, root1) = res
    assert s0 == 0 and e0 == s1 and e1 == 1200 and e0 < 600          # the heavy quarter gets the smaller slice
    assert abs(f0 - f1) / (f0 + f1) < 0.05                            # summed footprints balanced
    rng = np.random.default_rng(78)
    n = 1200
    pos = (rng.random((n, 3)) - 0.5) * 10; hs = rng.random(n) * 0.8 + 0.01
    hs[:300] *= 3.0
    m = rng.random(n) + 0.1; rho = rng.random(n) + 0.1; w = rng.random(n) + 0.5
    Q = np.stack([rng.random(n), rng.random(n) * 5.0], axis=1)
    a, _ = oracle.cic_mapping_2d(pos[:e0], hs[:e0], m[:e0], rho[:e0], Q[:e0], w[:e0], 6.4, 64, "WendlandC6", 2, True)
    b, _ = oracle.cic_mapping_2d(pos[e0:], hs[e0:], m[e0:], rho[e0:], Q[e0:], w[e0:], 6.4, 64, "WendlandC6", 2, True)
    tot = a + b
    ref = np.where(tot[:, 2:3] > 0, tot[:, :2] / np.where(tot[:, 2:3] > 0, tot[:, 2:3], 1.0), tot[:, :2]).ravel(order="F")
    for got in (all0, all1, root0):
        assert np.array_equal(got, ref)
    assert root1 is None


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys,
    timed on the oracle port, no GPU needed.  Tiny sample so that the test takes a second."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample", "2048", "--workload", "small"], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mparticles/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # under torchrun only rank 0 runs the arm; the other ranks exit 0 without work
    r2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                         "--warmup", "0", "--cpu-sample", "2048", "--workload", "small"], capture_output=True,
                        text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r2.returncode == 0 and not [l for l in r2.stdout.splitlines() if l.startswith("{")]


def test_flattened_footprint_index_without_integer_division():
    """csrc/s2g_cic2d.cu `unflatten`: (row, column) of a flattened footprint index e (column fastest, nj columns) from a
    single-precision reciprocal and ONE correction step instead of an integer division.  Restated in numpy float32 with
    the same operation order (float(e) + 0.5f, times fl(1/nj), round down, correct by one) and compared with divmod —
    exhaustively over the kernels' domain (footprints up to 1024 pixels) and on random large arguments."""
    def unflatten(e, nj):
        inv = (np.float32(1.0) / nj.astype(np.float32)).astype(np.float32)
        ir = np.floor((e.astype(np.float32) + np.float32(0.5)) * inv).astype(np.int64)
        jc = e - ir * nj
        lo, hi = jc < 0, jc >= nj
        ir = ir - lo + hi
        jc = jc + lo * nj - hi * nj
        return ir, jc

    nj, e = np.meshgrid(np.arange(1, 1025, dtype=np.int64), np.arange(0, 2048, dtype=np.int64), indexing="ij")
    ir, jc = unflatten(e.ravel(), nj.ravel())
    assert np.array_equal(ir, e.ravel() // nj.ravel()) and np.array_equal(jc, e.ravel() % nj.ravel())
    rng = np.random.default_rng(5)
    nj = rng.integers(1, 1 << 20, size=2_000_000)
    e = rng.integers(0, 1 << 20, size=2_000_000)
    ir, jc = unflatten(e, nj)
    assert np.array_equal(ir, e // nj) and np.array_equal(jc, e % nj)
