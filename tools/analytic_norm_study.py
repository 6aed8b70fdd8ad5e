#!/usr/bin/env python
"""How well does the analytic kernel integral h^2 * I2 (I2 = ∫ w(u) 2πu du = 1/norm_2D) approximate the discrete
pass-A sum  Σ_ij w(u_ij) dA_ij  of cic_2D.jl:11-72 for a particle whose footprint is NOT clipped by the image?
(Poisson summation: the difference is the kernel's Fourier transform at multiples of the pixel frequency, which
falls off with the kernel's smoothness.)  Prints, per kernel, the worst relative difference over random sub-pixel
offsets as a function of h [pixels]; the thresholds in csrc/s2g_gather2d.cu are where this stays below 5e-12."""
import math
import numpy as np

PI = math.pi
NORM2 = {"Cubic": 40 / (7 * PI), "Quintic": 3 ** 7 * 7 / (478 * PI), "WendlandC2": 7 / PI, "WendlandC4": 9 / PI,
         "WendlandC6": 78 / (7 * PI), "WendlandC8": 8 / (3 * PI)}


def shape(k, u):
    t = 1 - u
    if k == "Cubic":
        w = np.where(u < 0.5, 1 + 6 * (u - 1) * u * u, 2 * t ** 3)
    elif k == "Quintic":
        w = t ** 5 - 6 * np.maximum(2 / 3 - u, 0) ** 5 + 15 * np.maximum(1 / 3 - u, 0) ** 5
    elif k == "WendlandC2":
        w = t ** 4 * (1 + 4 * u)
    elif k == "WendlandC4":
        w = t ** 6 * (1 + 6 * u + 35 / 3 * u * u)
    elif k == "WendlandC6":
        w = t ** 8 * (1 + 8 * u + 25 * u * u + 32 * u ** 3)
    else:
        w = t ** 10 * (5 + 50 * u + 210 * u * u + 450 * u ** 3 + 429 * u ** 4)
    return np.where(u < 1, w, 0.0)


def worst(k, h, trials, rng):
    wmax = 0.0
    for _ in range(trials):
        x = 1000 + rng.random(); y = 1000 + rng.random(); hh = h * (1 + 0.25 * rng.random())
        i = np.arange(math.floor(x - hh), math.floor(x + hh) + 1)
        j = np.arange(math.floor(y - hh), math.floor(y + hh) + 1)
        dx = np.minimum(x + hh, i + 1) - np.maximum(x - hh, i)
        dy = np.minimum(y + hh, j + 1) - np.maximum(y - hh, j)
        u = np.sqrt(((x - i - 0.5) ** 2)[:, None] + ((y - j - 0.5) ** 2)[None, :]) / hh
        s = math.fsum((shape(k, u) * dx[:, None] * dy[None, :]).ravel().tolist())
        wmax = max(wmax, abs(s / (hh * hh / NORM2[k]) - 1))
    return wmax


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    hs = [16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 96, 128]
    print("kernel      " + " ".join(f"{h:8d}" for h in hs))
    for k in NORM2:
        print(f"{k:11s} " + " ".join(f"{worst(k, h, 60, rng):8.1e}" for h in hs))
