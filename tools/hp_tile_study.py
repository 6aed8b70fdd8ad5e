#!/usr/bin/env python
"""Round-2 planning aid (CPU only): how would the discs of the C4 workload (healpix_map, Nside 2048, WendlandC4)
fall onto tiles of a tile-owning HEALPix gather kernel?

For a sample of the synthetic C4 stream (same recipe as bench.py's host stand-in) the oracle's query_disc gives the
exact pixel list of every particle; the pixels are mapped to tiles = (band of RB consecutive rings) x (sector of the
ring: floor(in_ring_index * NSECT(ring) / ring_len), with NSECT = ceil(ring_len / SW)), and the script reports, per
tile shape: (tile, particle) pairs per particle, pixels of the disc per pair (how full the tiles are), and the
pairs a conservative rectangular bound (ring range x azimuth range at the widest ring of the band) would generate.
"""
import ctypes as C
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402

NSIDE = 2048
N_STREAM = 128 * 1024 * 1024
N_NGB = 200.0
SIGMA = 1.5


def ring_of(pix, nside):
    """ring index (1-based) and in-ring index of RING pixels (vectorised)."""
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    pix = np.asarray(pix, dtype=np.int64)
    ring = np.empty_like(pix); j = np.empty_like(pix); rlen = np.empty_like(pix)
    north = pix < ncap
    r = ((1 + np.sqrt(1 + 2 * pix[north].astype(np.float64))) / 2).astype(np.int64)
    r = np.where(2 * r * (r - 1) > pix[north], r - 1, r)
    r = np.where(2 * (r + 1) * r <= pix[north], r + 1, r)
    ring[north] = r; j[north] = pix[north] - 2 * r * (r - 1); rlen[north] = 4 * r
    belt = (~north) & (pix < npix - ncap)
    ip = pix[belt] - ncap
    ring[belt] = ip // (4 * nside) + nside; j[belt] = ip % (4 * nside); rlen[belt] = 4 * nside
    south = pix >= npix - ncap
    ip = npix - 1 - pix[south]
    r = ((1 + np.sqrt(1 + 2 * ip.astype(np.float64))) / 2).astype(np.int64)
    r = np.where(2 * r * (r - 1) > ip, r - 1, r)
    r = np.where(2 * (r + 1) * r <= ip, r + 1, r)
    ring[south] = 4 * nside - r; rlen[south] = 4 * r
    j[south] = 4 * r - 1 - (ip - 2 * r * (r - 1))
    return ring, j, rlen


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    rng = np.random.default_rng(4)
    L = orc.lib()
    pos = rng.random((n * 4, 3)) - 0.5
    g = rng.normal(size=n * 4)
    rho = np.exp(SIGMA * g - 0.5 * SIGMA ** 2)
    hsml = np.cbrt(3.0 * N_NGB / N_STREAM / (4.0 * np.pi * rho))
    d = np.linalg.norm(pos, axis=1)
    sel = (d >= 0.05) & (d <= 0.5) & (d > hsml)
    pos, hsml, d = pos[sel][:n], hsml[sel][:n], d[sel][:n]
    cap = 12 * NSIDE * NSIDE
    buf = np.zeros(4_000_000, dtype=np.int64)
    shapes = [(32, 64), (16, 128), (64, 32), (32, 128), (16, 64)]
    acc = {s: [0, 0, 0] for s in shapes}  # pairs, pixels, bound-pairs
    npx = 0
    for k in range(len(d)):
        th = math.acos(pos[k, 2] / d[k]); ph = math.atan2(pos[k, 1], pos[k, 0]) % (2 * math.pi)
        r = math.asin(hsml[k] / d[k])
        cnt = L.s2go_hp_query_disc_ring(NSIDE, th, ph, r, buf.ctypes.data_as(C.POINTER(C.c_int64)), len(buf))
        if cnt <= 0 or cnt > len(buf):
            continue
        ring, j, rlen = ring_of(buf[:cnt], NSIDE)
        npx += cnt
        for (rb, sw) in shapes:
            nsect = -(-rlen // sw)
            sect = (j * nsect) // rlen
            band = ring // rb
            key = band * 100000 + sect
            u = np.unique(key)
            acc[(rb, sw)][0] += len(u)
            acc[(rb, sw)][1] += cnt
            # conservative rectangle per band: sectors from min..max (with wrap -> all) of the band's pixels
            bp = 0
            for b in np.unique(band):
                s_b = sect[band == b]; ns_b = int(nsect[band == b].max())
                span = int(s_b.max() - s_b.min() + 1)
                if span > ns_b // 2:  # wrapped run
                    occupied = np.zeros(ns_b, dtype=bool); occupied[s_b % ns_b] = True
                    gaps = np.diff(np.flatnonzero(np.concatenate([occupied, occupied])))
                    span = ns_b - (int(gaps.max()) - 1) if occupied.sum() < ns_b else ns_b
                bp += span
            acc[(rb, sw)][2] += bp
    print(f"C4-like sample: {len(d)} particles in the shell, mean disc = {npx / len(d):.0f} pixels (Nside {NSIDE})")
    print("tile (rings x pixels)  pairs/particle  disc pixels per pair  tile fill  rectangle-bound pairs/particle")
    for (rb, sw) in shapes:
        p, px, bp = acc[(rb, sw)]
        print(f"  {rb:3d} x {sw:4d}          {p / len(d):8.2f}        {px / p:10.0f}       {px / p / (rb * sw):6.1%}"
              f"      {bp / len(d):8.2f}")


if __name__ == "__main__":
    main()
