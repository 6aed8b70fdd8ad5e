#!/usr/bin/env python
"""FP32-accumulate mode: per-kernel error statistics against the FP64 result of the same context (GPU needed).
Prints, per kernel: worst pixel of the 1e-5 bar with its values, and quantiles of the per-pixel relative error."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from util import KERNELS, kern, random_particles  # noqa: E402

s2g = ge.load_package()
pos, hsml, m, rho, q, w = random_particles(21, 5000, box=11.0, hmin=0.3, hmax=2.5)
npix = 256
par = s2g.mappingParameters(center=[0, 0, 0], x_size=10.0, y_size=10.0, z_size=10.0, Npixels=npix)
ctx = s2g.Context(0, strategy="gather")
for kname in KERNELS:
    ctx.set_accumulate_mode("f64")
    ref = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kname), calc_mean=True, ctx=ctx)
    ctx.set_accumulate_mode("f32")
    got = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, param=par, kernel=kern(s2g, kname), calc_mean=True, ctx=ctx)
    rec = {"kernel": kname}
    for k, nm in ((0, "quantity"), (1, "weight")):
        a, b = got[:, k], ref[:, k]
        pm = float(np.max(np.abs(b)))
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        bar = np.abs(a - b) / (1e-5 * np.maximum(np.abs(a), np.abs(b)) + 1e-9 * pm)
        i = int(np.argmax(bar))
        big = np.abs(b) > 1e-3 * pm
        rec[nm] = {"worst_bar": float(bar[i]), "pixel": [i // npix, i % npix], "got": float(a[i]), "ref": float(b[i]),
                   "ref_over_planemax": float(abs(b[i]) / pm),
                   "rel_q50_q99_max(px > 1e-3 max)": [float(np.quantile(rel[big], 0.5)), float(np.quantile(rel[big], 0.99)),
                                                      float(rel[big].max())],
                   "sum_rel": float(abs(a.sum() - b.sum()) / abs(b.sum()))}
    print(json.dumps(rec), flush=True)
