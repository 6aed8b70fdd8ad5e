#!/usr/bin/env python
"""Generates tests/golden/oracle_vectors.npz: seeded inputs and the CPU oracle's outputs for one small case per path
(2D multi-image, 3D, HEALPix, CIC/TSC, Stokes/Faraday compositing).  The vectors pin the oracle against silent drift (CPU test) and give the GPU
parity tests a committed target that does not depend on the oracle being rebuilt on the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc


def inputs():
    rng = np.random.default_rng(20261017)
    n = 300
    pos = (rng.random((n, 3)) - 0.5) * 11.0
    hsml = 0.02 + rng.random(n) ** 2 * 1.4
    hsml[:20] *= 0.02
    m = rng.random(n) + 0.1
    rho = rng.random(n) + 0.1
    q = rng.random(n) * 100.0
    q[5:12] = 0.0
    w = rng.random(n) + 0.5
    return pos, hsml, m, rho, q, w


def main():
    pos, hsml, m, rho, q, w = inputs()
    out = dict(pos=pos, hsml=hsml, m=m, rho=rho, q=q, w=w)
    npix2, npix3, nside = 64, 20, 16
    Q = np.stack([q, np.sqrt(q + 1.0)], axis=1)
    out["map2d_WendlandC6"] = orc.cic_mapping_2d(pos, hsml, m, rho, Q, w, npix2 / 10.0, npix2, "WendlandC6", 2, True)[0]
    out["map2d_Cubic_nomean"] = orc.cic_mapping_2d(pos, hsml, m, rho, q, w, npix2 / 10.0, npix2, "Cubic", 2, False)[0]
    out["map3d_WendlandC4"] = orc.cic_mapping_3d(pos, hsml, m, rho, q, w, npix3 / 10.0, npix3, "WendlandC4", 3, False)[0]
    hp = pos * 20.0 + np.array([3.0, -2.0, 1.0])
    a, wm, _ = orc.healpix_deposit(hp, hsml * 12.0, m, rho, q, w, nside, "WendlandC4", 2, True)
    out["hp_pos"] = hp
    out["hp_map"] = a
    out["hp_wmap"] = wm
    # the same particles through the extended-precision arbiter (oracle/s2g_oracle_exact.c, long double, one thread):
    # the yardstick the CUDA maps are held to (the literal Float64 acos form above is 1e-9 away from it at Nside 16)
    ea, ew, est = orc.healpix_deposit(hp, hsml * 12.0, m, rho, q, w, nside, "WendlandC4", 2, True, n_workers=1,
                                      exact="sens")
    out["hp_map_exact"] = ea
    out["hp_wmap_exact"] = ew
    out["hp_sens"] = est["sens"]
    out["hp_sens_q"] = est["sens_q"]
    out["cic3d"] = orc.stencil_deposit(2, 3, pos, q, npix3 / 10.0, npix3, False)
    out["tsc2d"] = orc.stencil_deposit(3, 2, pos, q, npix2 / 10.0, npix2, True)
    # ordered Stokes/Faraday compositing (cic_mapping_2D with RM, stokes=true): particles far -> near, RM*pw = O(1) rad
    rng = np.random.default_rng(4)
    order = np.argsort(pos[:, 2], kind="stable")[::-1]
    QU = rng.normal(size=(pos.shape[0], 2))
    rm = rng.normal(size=pos.shape[0]) * 0.3
    out["stokes_order"] = order
    out["stokes_qu"] = QU
    out["stokes_rm"] = rm
    out["stokes_WendlandC4"] = orc.cic_mapping_2d_rm(pos[order], hsml[order], m[order], rho[order], QU[order], w[order],
                                                     rm[order], npix2 / 10.0, npix2, "WendlandC4", 2, True, True)[0]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz"), **out)
    print({k: (v.shape, float(np.nansum(v))) for k, v in out.items() if k.startswith(("map", "hp_m", "hp_w", "cic", "tsc", "stokes_W"))})


if __name__ == "__main__":
    main()
