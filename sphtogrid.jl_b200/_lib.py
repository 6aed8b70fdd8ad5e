"""ctypes binding of libsphtogrid_cuda.so (C ABI declared in include/sphtogrid_cuda.h).

This is the boundary a Julia `ccall` would cross; nothing here computes anything.  There is NO fallback: if the
shared library is missing, or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsphtogrid_cuda.so")
CSRC = os.path.join(_HERE, "csrc")

S2G_OK, S2G_EINVAL, S2G_ECUDA, S2G_ENOMEM, S2G_EUNSUPPORTED, S2G_EINTERNAL = 0, -1, -2, -3, -4, -5
F32, F64 = 0, 1
STRATEGY = {"auto": 0, "scatter": 1, "gather": 2}
ACCUMULATE = {"f64": 0, "f32": 1}


class S2GError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsphtogrid_cuda error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_in", "n_mapped", "footprint_pixels", "touched_pixels", "n_fallback",
                                         "n_pairs", "n_scatter", "n_gather", "n_launches")] + \
               [(n, C.c_double) for n in ("ms_h2d", "ms_compute", "ms_d2h", "ms_total", "ms_prep", "ms_sort",
                                          "ms_norm", "ms_deposit", "ms_epilogue")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a with nvcc (works without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "sphtogrid_cuda.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", CSRC, "-j8"] + (["-B"] if force else []), capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout[-4000:], r.stderr[-4000:])
        if r.returncode != 0:
            raise RuntimeError("building libsphtogrid_cuda.so failed")
    return LIB_PATH


_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_dp = C.POINTER(C.c_double)
_SIGS = {
    "s2g_device_count": (C.c_int, []),
    "s2g_init": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "s2g_shutdown": (C.c_int, [_vp]),
    "s2g_sync": (C.c_int, [_vp]),
    "s2g_last_error": (C.c_char_p, []),
    "s2g_version": (C.c_char_p, []),
    "s2g_set_stream": (C.c_int, [_vp, _vp]),
    "s2g_set_strategy": (C.c_int, [_vp, C.c_int]),
    "s2g_set_exact_norm": (C.c_int, [_vp, C.c_int]),
    "s2g_set_accumulate_mode": (C.c_int, [_vp, C.c_int]),
    "s2g_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "s2g_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_uint64]),
    "s2g_host_free": (C.c_int, [_vp]),
    "s2g_dev_alloc": (C.c_int, [_vp, C.POINTER(_vp), C.c_uint64]),
    "s2g_dev_free": (C.c_int, [_vp, _vp]),
    "s2g_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "s2g_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "s2g_memset_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_uint64]),
    "s2g_deposit_2d": (C.c_int, [_vp] + [_vp] * 6 + [_i64, _i32, _i32, _f64, _i64, _i64, _i32, _i32, _vp,
                                                    C.POINTER(Stats)]),
    "s2g_deposit_2d_dev": (C.c_int, [_vp] + [_vp] * 6 + [_i64, _i32, _i32, _f64, _i64, _i64, _i32, _i32, _i32, _vp]),
    "s2g_deposit_2d_rm": (C.c_int, [_vp] + [_vp] * 7 + [_i64, _i32, _i32, _f64, _i64, _i64, _i32, _i32, _i32, _vp,
                                                       C.POINTER(Stats)]),
    "s2g_deposit_2d_rm_dev": (C.c_int, [_vp] + [_vp] * 7 + [_i64, _i32, _i32, _f64, _i64, _i64, _i32, _i32, _i32,
                                                           _vp]),
    "s2g_deposit_3d": (C.c_int, [_vp] + [_vp] * 6 + [_i64, _i32, _f64, _i64, _i32, _i32, _vp, C.POINTER(Stats)]),
    "s2g_deposit_3d_dev": (C.c_int, [_vp] + [_vp] * 6 + [_i64, _i32, _f64, _i64, _i32, _i32, _i32, _vp]),
    "s2g_footprints": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _f64, _i64, _i32, _vp]),
    "s2g_reduce_image_2d": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "s2g_reduce_image_2d_dev": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "s2g_reduce_image_3d": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "s2g_reduce_image_3d_dev": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "s2g_center_filter": (C.c_int, [_vp, _vp, _i64, _i32, _dp, _i32, _f64, _dp, _dp, _vp, _vp]),
    "s2g_sphmap": (C.c_int, [_vp, _i32] + [_vp] * 6 + [_i64, _i32, _i32, _dp, _i32, _f64, _dp, _f64, _i64, _i32, _i32,
                                                      _i32, _i32, _vp, _vp, C.POINTER(Stats)]),
    "s2g_sphmap_dev": (C.c_int, [_vp, _i32] + [_vp] * 6 + [_i64, _i32, _i32, _dp, _i32, _f64, _dp, _f64, _i64, _i32,
                                                          _i32, _i32, _vp]),
    "s2g_sphmap_projected": (C.c_int, [_vp, _i32] + [_vp] * 6 + [_i64, _i32, _i32, C.POINTER(C.c_int32), _dp, _dp, _i32,
                                                                _f64, _dp, _f64, _i64, _i32, _i32, _i32, _i32, _vp, _vp,
                                                                C.POINTER(Stats)]),
    "s2g_sphmap_projected_dev": (C.c_int, [_vp, _i32] + [_vp] * 6 + [_i64, _i32, _i32, C.POINTER(C.c_int32), _dp, _dp,
                                                                    _i32, _f64, _dp, _f64, _i64, _i32, _i32, _i32,
                                                                    _vp]),
    "s2g_healpix_deposit": (C.c_int, [_vp] + [_vp] * 6 + [_i64, _i32, _i64, _i32, _i32, _vp, _vp, C.POINTER(Stats)]),
    "s2g_healpix_deposit_dev": (C.c_int, [_vp] + [_vp] * 6 + [_i64, _i32, _i64, _i32, _i32, _i32, _vp, _vp]),
    "s2g_healpix_map": (C.c_int, [_vp] + [_vp] * 6 + [_i64, _dp, _dp, _i64, _i32, _i32, _vp, _vp, _vp, C.POINTER(Stats)]),
    "s2g_healpix_map_dev": (C.c_int, [_vp] + [_vp] * 6 + [_i64, _dp, _dp, _i64, _i32, _i32, _i32, _vp, _vp,
                                                        C.POINTER(C.c_int64)]),
    "s2g_healpix_pixels": (C.c_int, [_vp, _dp, _f64, _i64, _vp, _i64, C.POINTER(C.c_int64)]),
    "s2g_stencil_deposit": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i64, _i32, _f64, _i64, _i32, _vp, C.POINTER(Stats)]),
    "s2g_stencil_deposit_dev": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i64, _i32, _f64, _i64, _i32, _i32, _vp]),
    "s2g_accumulate_finite_dev": (C.c_int, [_vp, _vp, _vp, _i64]),
    "s2g_divide_slice_dev": (C.c_int, [_vp, _i32, _vp, _vp, _i64, _i64, _i32, _i32]),
    "s2g_domain_decomposition": (C.c_int, [_i64, _i32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "s2g_group_init": (C.c_int, [C.POINTER(C.c_int32), _i32, C.POINTER(_vp)]),
    "s2g_group_shutdown": (C.c_int, [_vp]),
    "s2g_group_size": (C.c_int, [_vp]),
    "s2g_group_peer_access": (C.c_int, [_vp]),
    "s2g_group_context": (C.c_int, [_vp, _i32, C.POINTER(_vp)]),
    "s2g_group_sphmap": (C.c_int, [_vp, _i32] + [_vp] * 6 + [_i64, _i32, _i32, _dp, _i32, _f64, _dp, _f64, _i64, _i32,
                                                            _i32, _i32, _i32, _vp, _vp, C.POINTER(Stats)]),
    "s2g_group_healpix_map": (C.c_int, [_vp] + [_vp] * 6 + [_i64, _dp, _dp, _i64, _i32, _i32, _vp, _vp, _vp,
                                                              C.POINTER(Stats)]),
    "s2g_synth_particles_dev": (C.c_int, [_vp, C.c_uint64, _i64, _i64, _i64, _f64, _f64, _f64, _i32] + [_vp] * 5),
    "s2g_microbench": (C.c_int, [_vp, _i32, C.c_uint64, _i32, _dp]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def lib():
    """The loaded library.  Raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build() / make -C {CSRC}")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != S2G_OK:
        raise S2GError(rc, lib().s2g_last_error().decode(errors="replace"))


def ptr(a):
    """void* of a numpy array / int address / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(_vp)
    return _vp(int(a))


def dbl3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


class Context:
    """One CUDA device + stream + scratch arena (s2g_ctx)."""

    def __init__(self, device: int = 0, strategy: str = "auto", exact_norm: bool = False):
        self._h = _vp()
        check(lib().s2g_init(int(device), C.byref(self._h)))
        self.device = device
        if strategy != "auto":
            self.set_strategy(strategy)
        if exact_norm:
            self.set_exact_norm(True)

    @property
    def handle(self):
        if not self._h:
            raise S2GError(S2G_EINVAL, "context was shut down")
        return self._h

    def set_strategy(self, strategy: str):
        check(lib().s2g_set_strategy(self.handle, STRATEGY[strategy]))

    def set_exact_norm(self, on: bool):
        check(lib().s2g_set_exact_norm(self.handle, int(bool(on))))

    def set_accumulate_mode(self, mode: str):
        """"f64" (default, parity bar 1e-10) or "f32" (FP32 partial sums in the 2D tile-gather kernel, bar 1e-5)."""
        check(lib().s2g_set_accumulate_mode(self.handle, ACCUMULATE[mode]))

    def set_stream(self, cuda_stream: int):
        check(lib().s2g_set_stream(self.handle, _vp(int(cuda_stream))))

    def sync(self):
        check(lib().s2g_sync(self.handle))

    def stats(self) -> dict:
        s = Stats()
        check(lib().s2g_get_stats(self.handle, C.byref(s)))
        return s.asdict()

    def close(self):
        if self._h:
            lib().s2g_shutdown(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _BorrowedContext(Context):
    """A context owned by a DeviceGroup (never shut down on its own)."""

    def __init__(self, handle, device):
        self._h = handle
        self.device = device

    def close(self):
        self._h = _vp()


class DeviceGroup:
    """Several GPUs driven by ONE process (s2g_group_*): what `parallel=true` means for a caller without a process
    group.  `devices=None` takes every visible device.  A device may be listed more than once."""

    def __init__(self, devices=None, strategy: str = "auto", exact_norm: bool = False):
        if devices is None:
            devices = list(range(lib().s2g_device_count()))
        devices = [int(d) for d in devices]
        self._h = _vp()
        arr = (C.c_int32 * max(len(devices), 1))(*devices)
        check(lib().s2g_group_init(arr, len(devices), C.byref(self._h)))
        self.devices = devices
        for r in range(len(devices)):
            c = self.context(r)
            if strategy != "auto":
                c.set_strategy(strategy)
            if exact_norm:
                c.set_exact_norm(True)

    @property
    def handle(self):
        if not self._h:
            raise S2GError(S2G_EINVAL, "device group was shut down")
        return self._h

    def __len__(self):
        return len(self.devices)

    @property
    def peer_access(self) -> bool:
        return bool(lib().s2g_group_peer_access(self.handle))

    def context(self, rank: int) -> Context:
        h = _vp()
        check(lib().s2g_group_context(self.handle, int(rank), C.byref(h)))
        return _BorrowedContext(h, self.devices[rank])

    def new_stats(self):
        return (Stats * len(self.devices))()

    def close(self):
        if self._h:
            lib().s2g_group_shutdown(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def domain_decomposition_c(n: int, n_parts: int):
    """s2g_domain_decomposition: (starts, counts) of parallel/domain_decomp.jl:7-17, 0-based; needs no device."""
    st = (C.c_int64 * n_parts)()
    ct = (C.c_int64 * n_parts)()
    check(lib().s2g_domain_decomposition(int(n), int(n_parts), st, ct))
    return list(st), list(ct)


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("LOCAL_RANK", "0")) if lib().s2g_device_count() > 1 else 0
        _default_ctx = Context(dev)
    return _default_ctx
