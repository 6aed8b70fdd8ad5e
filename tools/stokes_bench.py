"""Times the ordered Stokes/Faraday compositing path (s2g_deposit_2d_rm, stokes=1) against the plain deposit on the
same particles and prints one JSON line.  Usage: python tools/stokes_bench.py [n_particles] [npix]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    npix = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    s2g = load_package()
    rng = np.random.default_rng(3)
    pos = (rng.random((n, 3)) - 0.5)
    nngb = 295.0
    rho = np.exp(1.0 * rng.normal(size=n)) * n
    m = np.ones(n)
    hsml = (3 * nngb * m / (4 * np.pi * rho)) ** (1 / 3)
    q = rng.normal(size=(n, 2))
    w = np.ones(n)
    order = np.argsort(pos[:, 2], kind="stable")[::-1]  # far -> near like sphMapping(stokes=true)
    pos, hsml, m, rho, q = pos[order].copy(), hsml[order].copy(), m[order].copy(), rho[order].copy(), q[order].copy()
    par = s2g.mappingParameters(center=[0, 0, 0], x_size=1.0, y_size=1.0, z_size=1.0, Npixels=npix)
    k = s2g.WendlandC6(2)
    out = {}
    for name, kw in (("plain", dict()), ("stokes", dict(stokes=True))):
        rm = None if name == "plain" else rng.normal(size=n) * 1e-9
        best = None
        for _ in range(3):
            _, st = s2g.cic_mapping_2D(pos, hsml, m, rho, q, w, rm, param=par, kernel=k, return_stats=True, **kw)
            best = st if best is None or st["ms_compute"] < best["ms_compute"] else best
        out[name] = {"ms_compute": best["ms_compute"], "ms_norm": best["ms_norm"], "ms_sort": best["ms_sort"],
                     "ms_deposit": best["ms_deposit"], "footprint_pixels": best["footprint_pixels"],
                     "pairs": best["n_pairs"],
                     "Gpix_per_s": best["footprint_pixels"] / best["ms_compute"] / 1e6}
    out["n"] = n
    out["npix"] = npix
    print(json.dumps(out))


if __name__ == "__main__":
    main()
