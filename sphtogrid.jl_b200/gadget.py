"""Gadget-2 "format 2" snapshot blocks — the data format on the near side of the deposit path.

The reference reads its inputs with GadgetIO.jl (`read_header`, `read_block(file, "POS", parttype=0)`; call sites
test/runtests.jl:185-189, 387-399 and the distributed example docs/src/mapping.md:180-207).  GadgetIO.jl is a
third-party dependency outside the reference tree, so this restates the public on-disk format:

    label record : int32 8 | char[4] name | int32 (bytes of the following data record incl. its two markers) | int32 8
    data record  : int32 nbytes | payload | int32 nbytes

`HEAD` holds the 256-byte header; `POS`, `VEL`, `ID` carry every particle type with npart > 0 (type order), `MASS` the
types with npart > 0 and massarr == 0, every other block gas only — unless an `INFO` block says otherwise.  Arrays come
back particle-major: `POS` is (N,3) C-contiguous, i.e. exactly the memory of the Julia Matrix(3,N) the C ABI expects.

`SnapshotPrefetcher` overlaps reading sub-file k+1 with the device work on sub-file k (the library calls release the
GIL), which is the double buffering the streaming entry (`distributed_cic_map`) needs on boxes larger than memory."""
from __future__ import annotations

import os
import struct
import threading
from dataclasses import dataclass, field

import numpy as np

_HEADER_FMT = "<6i6d2d2i6I2i4d2i6I"   # ... up to npartTotalHighWord; the rest of the 256 bytes is padding/extensions
_HEADER_LEN = 256
_ALL_TYPES = ("POS", "VEL", "ID")


@dataclass
class SnapshotHeader:
    npart: list = field(default_factory=lambda: [0] * 6)
    massarr: list = field(default_factory=lambda: [0.0] * 6)
    time: float = 0.0
    z: float = 0.0
    flag_sfr: int = 0
    flag_feedback: int = 0
    nall: list = field(default_factory=lambda: [0] * 6)
    flag_cooling: int = 0
    num_files: int = 1
    boxsize: float = 0.0
    omega_0: float = 0.0
    omega_l: float = 0.0
    h0: float = 1.0
    flag_stellarage: int = 0
    flag_metals: int = 0
    npartTotalHighWord: list = field(default_factory=lambda: [0] * 6)

    def pack(self) -> bytes:
        b = struct.pack(_HEADER_FMT, *self.npart, *self.massarr, self.time, self.z, self.flag_sfr, self.flag_feedback,
                        *self.nall, self.flag_cooling, self.num_files, self.boxsize, self.omega_0, self.omega_l,
                        self.h0, self.flag_stellarage, self.flag_metals, *self.npartTotalHighWord)
        return b + b"\0" * (_HEADER_LEN - len(b))

    @classmethod
    def unpack(cls, raw: bytes) -> "SnapshotHeader":
        v = struct.unpack_from(_HEADER_FMT, raw)
        return cls(list(v[0:6]), list(v[6:12]), v[12], v[13], v[14], v[15], list(v[16:22]), v[22], v[23], v[24], v[25],
                   v[26], v[27], v[28], v[29], list(v[30:36]))


def _resolve(filename):
    """`snap_011` may be a single file or the base name of `snap_011.0`, `snap_011.1`, …"""
    if os.path.isfile(filename):
        return filename
    if os.path.isfile(filename + ".0"):
        return filename + ".0"
    raise FileNotFoundError(filename)


def _scan_blocks(f):
    """{name: (offset of payload, payload bytes)} by walking the label records.  Record markers are UNSIGNED 32-bit:
    a block of >= 4 GiB (POS in Float32 from ~358 M particles, in Float64 from ~179 M) wraps modulo 2^32, so the true
    payload size is the smallest `marker + k*2^32` whose trailing marker matches and after which the file either ends
    or continues with a well-formed label record (what GadgetIO does with the header's particle counts, without
    needing the element type)."""
    out = {}
    f.seek(0, os.SEEK_END)
    size = f.tell()
    pos = 0
    while pos + 16 <= size:
        f.seek(pos)
        lead, name, nxt, trail = struct.unpack("<I4sII", f.read(16))
        if lead != 8 or trail != 8:
            raise ValueError("not a Gadget format-2 snapshot (label record marker != 8; format 1 and big-endian files "
                             "are not supported)")
        marker, = struct.unpack("<I", f.read(4))
        nbytes = None
        cand = marker
        while pos + 20 + cand + 4 <= size:
            f.seek(pos + 20 + cand)
            tail, = struct.unpack("<I", f.read(4))
            end = pos + 24 + cand
            ok = tail == marker
            if ok and end + 16 <= size:
                l2, _, _, t2 = struct.unpack("<I4sII", f.read(16))
                ok = l2 == 8 and t2 == 8
            elif ok:
                ok = end == size
            if ok:
                nbytes = cand
                break
            cand += 1 << 32
        if nbytes is None:
            raise ValueError(f"block {name!r}: no payload size congruent to its record marker {marker} (mod 2^32) "
                             "ends on a matching trailing marker")
        if (nxt - 8) % (1 << 32) != marker:
            raise ValueError(f"block {name!r}: label record and data record disagree on the payload size")
        out[name.decode("ascii").strip()] = (pos + 20, nbytes)
        pos += 24 + nbytes
    return out


def read_header(filename) -> SnapshotHeader:
    """`GadgetIO.read_header(filename)`"""
    with open(_resolve(filename), "rb") as f:
        off, n = _scan_blocks(f)["HEAD"]
        f.seek(off)
        return SnapshotHeader.unpack(f.read(n))


def block_present(filename, blockname) -> bool:
    with open(_resolve(filename), "rb") as f:
        return blockname.strip() in _scan_blocks(f)


def _info(f, blocks):
    """INFO block: per block (name[4], dtype[8], ndim int32, is_present[6] int32) = 40 bytes"""
    if "INFO" not in blocks:
        return {}
    off, n = blocks["INFO"]
    f.seek(off)
    raw = f.read(n)
    out = {}
    for k in range(n // 40):
        name, dt, ndim, *present = struct.unpack_from("<4s8si6i", raw, 40 * k)
        out[name.decode("ascii").strip()] = (dt.decode("ascii").strip(), ndim, present)
    return out


def _types_in_block(name, h: SnapshotHeader, info):
    if name in info:
        return [t for t in range(6) if info[name][2][t] and h.npart[t] > 0]
    if name in _ALL_TYPES:
        return [t for t in range(6) if h.npart[t] > 0]
    if name == "MASS":
        return [t for t in range(6) if h.npart[t] > 0 and h.massarr[t] == 0.0]
    return [0] if h.npart[0] > 0 else []


def read_block(filename, blockname, parttype=0, dtype=None):
    """`GadgetIO.read_block(filename, blockname, parttype=…)` for one (sub-)file.  `parttype=-1` returns every type the
    block holds.  Vector blocks come back (N, ndim) C-contiguous."""
    name = blockname.strip()
    with open(_resolve(filename), "rb") as f:
        blocks = _scan_blocks(f)
        off, n = blocks["HEAD"]
        f.seek(off)
        h = SnapshotHeader.unpack(f.read(n))
        info = _info(f, blocks)
        if name == "MASS" and parttype >= 0 and h.massarr[parttype] != 0.0:
            return np.full(h.npart[parttype], h.massarr[parttype], dtype=dtype or np.float32)
        if name not in blocks:
            raise KeyError(f"Block {name} not present!")
        types = _types_in_block(name, h, info)
        ntot = sum(h.npart[t] for t in types)
        off, nbytes = blocks[name]
        if ntot == 0:
            return np.zeros(0, dtype=dtype or np.float32)
        if name in info:
            dt = {"FLOAT": "<f4", "FLOATN": "<f4", "DOUBLE": "<f8", "DOUBLEN": "<f8", "LONG": "<u4", "LLONG": "<u8"}[info[name][0]]
            ndim = info[name][1]
        else:
            ndim = 3 if name in ("POS", "VEL", "BFLD", "ACCE") else 1
            item = nbytes // (ntot * ndim)
            dt = ("<u4" if item == 4 else "<u8") if name == "ID" else ("<f4" if item == 4 else "<f8")
        if nbytes != ntot * ndim * int(dt[2]):
            raise ValueError(f"block {name}: {nbytes} bytes do not match {ntot} particles x {ndim} x {dt}")
        if parttype >= 0:
            if parttype not in types:
                raise KeyError(f"Block {name} not present for particle type {parttype}")
            first = sum(h.npart[t] for t in types if t < parttype)
            count = h.npart[parttype]
        else:
            first, count = 0, ntot
        f.seek(off + first * ndim * int(dt[2]))
        a = np.fromfile(f, dtype=dt, count=count * ndim)
    a = a.reshape(count, ndim) if ndim > 1 else a
    return a.astype(dtype) if dtype is not None else a.astype(dt[1:])


def write_snapshot(filename, header: SnapshotHeader, blocks: dict, with_info=False):
    """Writes a format-2 (sub-)file: `blocks` maps a block name to the payload array already concatenated over the
    particle types the block carries.  Used by the tests and for synthetic inputs."""
    def record(f, name, payload: bytes):
        m32 = 1 << 32   # record markers are unsigned 32-bit and wrap for blocks >= 4 GiB
        f.write(struct.pack("<I4sII", 8, f"{name:<4}".encode("ascii"), (len(payload) + 8) % m32, 8))
        f.write(struct.pack("<I", len(payload) % m32))
        f.write(payload)
        f.write(struct.pack("<I", len(payload) % m32))

    with open(filename, "wb") as f:
        record(f, "HEAD", header.pack())
        if with_info:
            raw = b""
            for name, arr in blocks.items():
                arr = np.asarray(arr)
                kind = {"f4": "FLOAT", "f8": "DOUBLE", "u4": "LONG", "u8": "LLONG"}[arr.dtype.str[1:]]
                if arr.ndim > 1 and kind in ("FLOAT", "DOUBLE"):
                    kind += "N"
                present = [1 if t in _types_in_block(name, header, {}) else 0 for t in range(6)]
                raw += struct.pack("<4s8si6i", f"{name:<4}".encode(), f"{kind:<8}".encode(),
                                   arr.shape[1] if arr.ndim > 1 else 1, *present)
            record(f, "INFO", raw)
        for name, arr in blocks.items():
            arr = np.ascontiguousarray(arr)
            record(f, name, arr.astype(arr.dtype.newbyteorder("<")).tobytes())


class SnapshotPrefetcher:
    """Iterates `loader(subfile)` over `subfiles`, reading one sub-file ahead on a background thread so that the file
    system works while the GPU deposits the previous one (double buffering; at most two loaded sub-files alive)."""

    def __init__(self, subfiles, loader):
        self._subfiles = list(subfiles)
        self._loader = loader

    def __iter__(self):
        nxt, box = None, {}

        def work(sf):
            try:
                box["v"] = self._loader(sf)
            except BaseException as e:  # re-raised on the consumer side
                box["e"] = e

        for k, sf in enumerate(self._subfiles):
            if nxt is None:
                work(sf)
            else:
                nxt.join()
            if "e" in box:
                raise box.pop("e")
            cur = box.pop("v")
            nxt = None
            if k + 1 < len(self._subfiles):
                nxt = threading.Thread(target=work, args=(self._subfiles[k + 1],), daemon=True)
                nxt.start()
            yield sf, cur
