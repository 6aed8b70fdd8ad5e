#!/usr/bin/env python
"""One-process multi-GPU path (DeviceGroup, s2g_group_sphmap) timed end to end from host arrays against the
single-context call on the same inputs: S2G_GROUP_BENCH_N (4 Mi) particles of the c2 stream -> S2G_GROUP_BENCH_NPIX
(2048)^2, WendlandC6, calc_mean T map; times are wall clock around the public call, host numpy arrays in and out.
With one GPU the group lists device 0 twice (functional check of sharding + peer sum; no speed-up expected)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

s2g = ge.load_package()
nd = s2g.lib().s2g_device_count()
n = int(os.environ.get("S2G_GROUP_BENCH_N", 4 << 20))
wl = dict(bench.WORKLOADS["c2"])
pos, hsml, m, rho, temp = bench.host_particles(wl, n)
npix = int(os.environ.get("S2G_GROUP_BENCH_NPIX", 2048))
par = s2g.mappingParameters(center=[0.5, 0.5, 0.5], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=npix)
kern = s2g.WendlandC6(2)


def run(**kw):
    best, out = 1e30, None
    for _ in range(3):
        p = pos.copy()
        t0 = time.perf_counter()
        out = s2g.sphMapping(p, hsml, m, rho, temp, rho, param=par, kernel=kern, calc_mean=True, show_progress=False,
                             return_stats=True, **kw)
        best = min(best, time.perf_counter() - t0)
    return best, out


t1, (a, st1) = run(ctx=s2g.Context(0))
devs = list(range(nd)) if nd > 1 else [0, 0]
grp = s2g.DeviceGroup(devs)
tg, (b, stg) = run(parallel=True, group=grp)
err = float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300)))
print(json.dumps({"particles": n, "npix": npix, "devices": devs, "peer_access": grp.peer_access,
                  "single_ms": t1 * 1e3, "group_ms": tg * 1e3, "max_rel_diff_group_vs_single": err,
                  "single_stats_ms": {k: st1[k] for k in ("ms_h2d", "ms_compute", "ms_epilogue", "ms_d2h")},
                  "group_stats_ms": [{k: s[k] for k in ("ms_h2d", "ms_compute", "ms_epilogue", "ms_d2h")} for s in stg]}))
