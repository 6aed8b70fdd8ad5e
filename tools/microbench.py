#!/usr/bin/env python
"""Measures the roofline denominators MEASURED_PEAKS.json lacks: FP64 DFMA rate, red.global.add.f64 rates, smem f64
atomics, and a copy for cross-checking hbm_gbs.  Prints one JSON line; run under gpurun."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

s2g = ge.load_package()
from sphtogrid_b200 import _lib

ctx = s2g.Context(0)
out = {}


def run(which, nbytes, iters):
    r = C.c_double(0)
    _lib.check(s2g.lib().s2g_microbench(ctx.handle, which, nbytes, iters, C.byref(r)))
    return r.value


out["fp64_gflops"] = run(0, 0, 20000)
for mb in (8, 128, 1024, 4096):
    out[f"red_rows_gps_{mb}MB"] = run(1, mb << 20, 2000)
    out[f"red_random_gps_{mb}MB"] = run(2, mb << 20, 500)
out["smem_atomic_f64_gps"] = run(4, 0, 20000)
out["copy_gbs_1GB"] = run(3, 1 << 30, 5)
print(json.dumps(out))
