"""Gadget format-2 snapshot blocks (sphtogrid.jl_b200/gadget.py — the reader the reference gets from GadgetIO.jl,
call sites test/runtests.jl:185-189, docs/src/mapping.md:180-207) and the prefetching sub-file loop."""
import struct
import threading
import time

import numpy as np
import pytest


def _snapshot(tmp_path, s2g, name, ngas, ndm, seed, with_info, massarr_dm=0.5, f64=False):
    from sphtogrid_b200 import gadget
    rng = np.random.default_rng(seed)
    ft = np.float64 if f64 else np.float32
    h = gadget.SnapshotHeader(npart=[ngas, ndm, 0, 0, 0, 0], massarr=[0.0, massarr_dm, 0, 0, 0, 0], time=0.5, z=1.0,
                              nall=[ngas, ndm, 0, 0, 0, 0], num_files=1, boxsize=100.0, omega_0=0.3, omega_l=0.7, h0=0.7)
    blocks = {"POS": rng.random((ngas + ndm, 3)).astype(ft) * 100, "VEL": rng.normal(size=(ngas + ndm, 3)).astype(ft),
              "ID": np.arange(ngas + ndm, dtype=np.uint32),
              "MASS": (rng.random(ngas + (ndm if massarr_dm == 0.0 else 0)) + 1).astype(ft),
              "RHO": (rng.random(ngas) + 0.1).astype(ft), "HSML": (rng.random(ngas) + 0.5).astype(ft),
              "U": rng.random(ngas).astype(ft)}
    fn = str(tmp_path / name)
    gadget.write_snapshot(fn, h, blocks, with_info=with_info)
    return fn, h, blocks


@pytest.mark.parametrize("with_info", [False, True])
@pytest.mark.parametrize("f64", [False, True])
def test_read_blocks_round_trip(s2g, tmp_path, with_info, f64):
    from sphtogrid_b200 import gadget
    fn, h, b = _snapshot(tmp_path, s2g, "snap_000", 1000, 300, 0, with_info, f64=f64)
    hh = gadget.read_header(fn)
    assert hh.npart == h.npart and hh.massarr == h.massarr and hh.boxsize == 100.0 and hh.h0 == 0.7 and hh.z == 1.0
    pos = gadget.read_block(fn, "POS", parttype=0)
    assert pos.shape == (1000, 3) and pos.flags.c_contiguous and np.array_equal(pos, b["POS"][:1000])
    assert pos.dtype == (np.float64 if f64 else np.float32)
    assert np.array_equal(gadget.read_block(fn, "POS", parttype=1), b["POS"][1000:])
    assert np.array_equal(gadget.read_block(fn, "POS", parttype=-1), b["POS"])
    assert np.array_equal(gadget.read_block(fn, "HSML", parttype=0), b["HSML"])
    assert np.array_equal(gadget.read_block(fn, "RHO"), b["RHO"])
    assert np.array_equal(gadget.read_block(fn, "MASS", parttype=0), b["MASS"])
    m1 = gadget.read_block(fn, "MASS", parttype=1)            # massarr != 0: constant mass, not stored
    assert m1.shape == (300,) and np.all(m1 == np.float32(0.5))
    assert np.array_equal(gadget.read_block(fn, "ID", parttype=1), np.arange(1000, 1300, dtype=np.uint32))
    assert gadget.block_present(fn, "U") and not gadget.block_present(fn, "BFLD")
    with pytest.raises(KeyError):
        gadget.read_block(fn, "BFLD", parttype=0)
    with pytest.raises(KeyError):
        gadget.read_block(fn, "RHO", parttype=1)              # gas-only block


def test_on_disk_layout_and_errors(s2g, tmp_path):
    from sphtogrid_b200 import gadget
    fn, h, b = _snapshot(tmp_path, s2g, "snap_001.0", 10, 0, 1, False)
    raw = open(fn, "rb").read()
    assert struct.unpack_from("<i4sii", raw, 0) == (8, b"HEAD", 264, 8)
    assert struct.unpack_from("<i", raw, 16)[0] == 256 and struct.unpack_from("<i", raw, 20 + 256)[0] == 256
    assert struct.unpack_from("<i4sii", raw, 16 + 264) == (8, b"POS ", 10 * 12 + 8, 8)
    assert gadget.read_header(str(tmp_path / "snap_001")).npart[0] == 10    # base name resolves to sub-file .0
    bad = tmp_path / "not_a_snapshot"
    bad.write_bytes(b"\1" * 64)
    with pytest.raises(ValueError):
        gadget.read_header(str(bad))
    with pytest.raises(FileNotFoundError):
        gadget.read_header(str(tmp_path / "missing"))


def test_prefetcher_overlaps_and_propagates_errors(s2g):
    from sphtogrid_b200.gadget import SnapshotPrefetcher
    log, main = [], threading.get_ident()

    def loader(k):
        log.append((k, threading.get_ident() != main))
        time.sleep(0.05)
        return k * 10

    t0 = time.time()
    got = []
    for sf, data in SnapshotPrefetcher(range(4), loader):
        time.sleep(0.05)                                       # "device work" on the current sub-file
        got.append((sf, data))
    assert got == [(0, 0), (1, 10), (2, 20), (3, 30)]
    assert [k for k, _ in log] == [0, 1, 2, 3] and [bg for _, bg in log] == [False, True, True, True]
    assert time.time() - t0 < 0.36                             # serial would be 0.40 s

    def failing(k):
        if k == 2:
            raise OSError("disk")
        return k

    with pytest.raises(OSError):
        list(SnapshotPrefetcher(range(4), failing))


@pytest.mark.gpu
def test_distributed_cic_map_from_subfiles(s2g, oracle, tmp_path):
    """docs/src/mapping.md:152-233 end to end: four sub-files -> read_block -> sphMapping(return_both_maps) ->
    distributed_cic_map (prefetching loader) -> FITS; against the oracle on the concatenated particles."""
    from sphtogrid_b200 import gadget
    base = str(tmp_path / "snap_011")
    parts = []
    for k in range(4):
        fn, h, b = _snapshot(tmp_path, s2g, f"snap_011.{k}", 3000 + 100 * k, 50, 10 + k, with_info=bool(k % 2))
        parts.append(b)
    kw = dict(center=[50.0, 50.0, 50.0], x_size=60.0, y_size=60.0, z_size=60.0, Npixels=96)
    par = s2g.mappingParameters(**kw)
    k4 = s2g.WendlandC4(2)

    def loader(sub):
        f = f"{base}.{sub}"
        return {n: gadget.read_block(f, n, parttype=0) for n in ("POS", "HSML", "RHO", "MASS", "U")}

    def mapper(sub, d):
        m = s2g.sphMapping(d["POS"], d["HSML"], d["MASS"], d["RHO"], d["U"], d["RHO"], param=par, kernel=k4,
                           show_progress=False, parallel=False, return_both_maps=True)
        return m[:, :-1], m[:, -1]

    out = str(tmp_path / "map.fits")
    img = s2g.distributed_cic_map(out, 4, mapper, par, loader=loader, snap=11, units="erg")
    n = [3000 + 100 * k for k in range(4)]
    cat = {name: np.concatenate([p[name][:n[k]] for k, p in enumerate(parts)]) for name in ("POS", "HSML", "RHO", "MASS", "U")}
    opar = oracle.mapping_parameters(**kw)
    ref = oracle.sph_mapping(cat["POS"].copy(), cat["HSML"], cat["MASS"], cat["RHO"], cat["U"], cat["RHO"], param=opar,
                             kernel="WendlandC4", reduce_image=True)
    from util import assert_parity
    assert_parity(img, ref, 1e-10, "distributed_cic_map over sub-files")
    back, rpar, snap, units = s2g.read_fits_image(out)
    assert snap == 11 and units == "erg" and np.array_equal(back, img[:, :, 0])


def test_block_markers_are_unsigned_and_wrap_modulo_4gib(tmp_path):
    """record markers are unsigned 32-bit; a block of >= 4 GiB wraps.  A sparse file with a 4 GiB + 24 B
    'POS ' payload (marker 24) followed by a small 'ID  ' block must scan to the right offsets and sizes."""
    import struct
    from sphtogrid_b200 import gadget
    path = tmp_path / "snap_big"
    big = (1 << 32) + 24
    try:
        with open(path, "wb") as f:
            def label(name, payload_len):
                f.write(struct.pack("<I4sII", 8, name, (payload_len + 8) % (1 << 32), 8))
            label(b"HEAD", 256)
            f.write(struct.pack("<I", 256)); f.write(b"\0" * 256); f.write(struct.pack("<I", 256))
            label(b"POS ", big)
            f.write(struct.pack("<I", big % (1 << 32)))
            f.seek(big, 1)                                   # sparse payload
            f.write(struct.pack("<I", big % (1 << 32)))
            label(b"ID  ", 16)
            f.write(struct.pack("<I", 16)); f.write(b"\1" * 16); f.write(struct.pack("<I", 16))
    except OSError as e:
        pytest.skip(f"cannot create a sparse 4 GiB file here: {e}")
    with open(path, "rb") as f:
        blocks = gadget._scan_blocks(f)
    assert blocks["HEAD"] == (20, 256)
    pos_off = 20 + 256 + 4 + 16 + 4
    assert blocks["POS"] == (pos_off, big)
    assert blocks["ID"] == (pos_off + big + 4 + 16 + 4, 16)
