#!/usr/bin/env python
"""red.global.add.f64 (per-lane) vs cp.reduce.async.bulk .add.f64 (TMA, 256 B per op) on the deposit's access shape."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
s2g = ge.load_package()
from sphtogrid_b200 import _lib
ctx = s2g.Context(0)
out = {}
for which, name in ((1, "red_rows"), (5, "bulk256B_rows"), (6, "bulk2KiB_rows")):
    for mb in (64, 1024):
        r = C.c_double(0)
        _lib.check(s2g.lib().s2g_microbench(ctx.handle, which, mb << 20, 2000, C.byref(r)))
        out[f"{name}_gadds_{mb}MB"] = r.value
print(json.dumps(out))
