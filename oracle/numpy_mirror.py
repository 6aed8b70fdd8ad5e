"""
numpy_mirror.py — second, independent restatement (pure Python loops) of the reference's deposit
loops, used ONLY to cross-check oracle/s2g_oracle.c on small cases.  TEST INFRASTRUCTURE.

Follows: src/cic_interpolation/cic_2D.jl:11-244, cic_3D.jl:13-209, cic_shared.jl:9-121,
src/healpix_interpolation/main.jl:143-213, pixel_weights.jl:6-140, constributing_pixels.jl:7-22.
Written against the Julia text, not against the C file, so that a typo in one shows up as a diff.
"""
import math

import numpy as np

PI = math.pi


def W(kernel, dim, u, h_inv):
    """SPHKernels.jl v2 kernel values (third-party; shapes restated)."""
    norms = {
        "Cubic": {2: 40.0 / (7.0 * PI), 3: 8.0 / PI},
        "Quintic": {2: 3 ** 7 * 7.0 / (478.0 * PI), 3: 3 ** 7 / (40.0 * PI)},
        "WendlandC2": {2: 7.0 / PI, 3: 21.0 / (2.0 * PI)},
        "WendlandC4": {2: 9.0 / PI, 3: 495.0 / (32.0 * PI)},
        "WendlandC6": {2: 78.0 / (7.0 * PI), 3: 1365.0 / (64.0 * PI)},
        "WendlandC8": {2: 8.0 / (3.0 * PI), 3: 357.0 / (64.0 * PI)},
    }
    n = norms[kernel][dim] * h_inv ** dim
    if u >= 1.0:
        return 0.0
    if kernel == "Cubic":
        w = 1.0 + 6.0 * (u - 1.0) * u ** 2 if u < 0.5 else 2.0 * (1.0 - u) ** 3
    elif kernel == "Quintic":
        w = (1 - u) ** 5 - 6 * max(2 / 3 - u, 0.0) ** 5 + 15 * max(1 / 3 - u, 0.0) ** 5
    elif kernel == "WendlandC2":
        w = (1 - u) ** 4 * (1 + 4 * u)
    elif kernel == "WendlandC4":
        w = (1 - u) ** 6 * (1 + 6 * u + 35 / 3 * u ** 2)
    elif kernel == "WendlandC6":
        w = (1 - u) ** 8 * (1 + 8 * u + 25 * u ** 2 + 32 * u ** 3)
    elif kernel == "WendlandC8":
        w = (1 - u) ** 10 * (5 + 50 * u + 210 * u ** 2 + 450 * u ** 3 + 429 * u ** 4)
    else:
        raise KeyError(kernel)
    return w * n


def _minmax(x, h, n):
    return max(math.floor(x - h), 0), min(math.floor(x + h), n - 1)


def _x_dx(x, h, i):
    return x - i - 0.5, min(x + h, i + 1) - max(x - h, i)


def _jl_mod(x, y):
    # Julia mod(x::Float64, y::Float64): rem, then move into the sign of y
    r = math.fmod(x, y)
    if r == 0:
        return math.copysign(r, y)
    if (r > 0) != (y > 0):
        return r + y
    return r


def cic_mapping_2d(pos, hsml, m, rho, binq, w, len2pix, npix, kernel, kdim=2, calc_mean=True, rm=None, stokes=False):
    """rm/stokes: the Faraday-rotation branch of cic_2D.jl:201-217 + cic_shared.jl:129-159."""
    pos = np.asarray(pos, float); binq = np.asarray(binq, float)
    n = len(hsml)
    nim = 1 if binq.ndim == 1 else binq.shape[1]
    img = np.zeros((npix * npix, nim + 1))
    touched = set()
    for p in range(n):
        bq = np.atleast_1d(binq[p])
        allzero = bool(np.all(bq == 0))
        if allzero and not calc_mean:
            continue
        h = hsml[p] * len2pix
        hinv = 1.0 / h
        area = (2 * h) ** 2
        rr = rho[p] * (1.0 / (len2pix * len2pix * len2pix))
        dz = m[p] / rr / area
        x = pos[p, 0] * len2pix + 0.5 * npix
        y = pos[p, 1] * len2pix + 0.5 * npix
        i0, i1 = _minmax(x, h, npix)
        j0, j1 = _minmax(y, h, npix)
        wk = {}
        A = {}
        ndist = ntot = 0
        dw = da = 0.0
        for i in range(i0, i1 + 1):
            xd, dx = _x_dx(x, h, i)
            for j in range(j0, j1 + 1):
                yd, dy = _x_dx(y, h, j)
                u = math.sqrt(xd * xd + yd * yd) * hinv
                dA = dx * dy
                idx = i * npix + j
                A[idx] = dA
                da += dA
                ntot += 1
                if u <= 1:
                    k = W(kernel, kdim, u, hinv)
                    dw += k * dA
                    ndist += 1
                    wk[idx] = k
                else:
                    wk[idx] = 0.0
        if dw == 0.0:
            ndist = ntot
            for key in wk:
                wk[key] = 1.0
            wpp = ndist / da if da != 0 else 1.0
        else:
            wpp = ndist / dw
        if ndist == 0:
            continue  # empty footprint: nothing to write (area/0 never used)
        kn = area / ndist
        an = kn * wpp * w[p] * dz
        for idx in wk:
            pw = wk[idx] * A[idx] * an
            if rm is not None and idx in touched and stokes:
                ang = _jl_mod(rm[p] * pw, math.pi)
                Q, U = img[idx, 0], img[idx, 1]
                ip = math.sqrt(Q * Q + U * U)
                with np.errstate(all="ignore"):
                    psi = 0.5 * math.atan(np.float64(U) / np.float64(Q))
                img[idx, 0] = ip * math.cos(2 * (psi + ang))
                img[idx, 1] = ip * math.sin(2 * (psi + ang))
            if pw != 0.0:
                touched.add(idx)
                img[idx, nim] += pw
                if allzero:
                    img[idx, 0] += 0.0 * pw
                else:
                    for q in range(nim):
                        img[idx, q] += bq[q] * pw
    return img


def cic_mapping_3d(pos, hsml, m, rho, binq, w, len2pix, npix, kernel, kdim=3, calc_mean=False):
    pos = np.asarray(pos, float)
    n = len(hsml)
    img = np.zeros((npix ** 3, 2))
    for p in range(n):
        bq = float(binq[p])
        if bq == 0 and not calc_mean:
            continue
        h = hsml[p] * len2pix
        hinv = 1.0 / h
        rr = rho[p] / len2pix ** 3
        vol = m[p] / rr
        x = pos[p, 0] * len2pix + 0.5 * npix
        y = pos[p, 1] * len2pix + 0.5 * npix
        z = pos[p, 2] * len2pix + 0.5 * npix
        i0, i1 = _minmax(x, h, npix)
        j0, j1 = _minmax(y, h, npix)
        k0, k1 = _minmax(z, h, npix)
        wk = {}
        V = {}
        ndist = ntot = 0
        dw = dv = 0.0
        for i in range(i0, i1 + 1):
            xd, dx = _x_dx(x, h, i)
            for j in range(j0, j1 + 1):
                yd, dy = _x_dx(y, h, j)
                for k in range(k0, k1 + 1):
                    zd, dzz = _x_dx(z, h, k)
                    idx = i * npix * npix + j * npix + k
                    dV = dx * dy * dzz
                    u = math.sqrt(xd * xd + yd * yd + zd * zd) * hinv
                    V[idx] = dV
                    dv += dV
                    ntot += 1
                    if u <= 1:
                        kk = W(kernel, kdim, u, hinv)
                        dw += kk * dV
                        ndist += 1
                        wk[idx] = kk
                    else:
                        wk[idx] = 0.0
        if dw == 0.0:
            ndist = ntot
            for key in wk:
                wk[key] = 1.0
            wpp = ndist / dv if dv != 0 else 1.0
        else:
            wpp = ndist / dw
        if ndist == 0:
            continue
        kn = vol / ndist
        vn = kn * wpp * w[p] * len2pix
        for idx in wk:
            pw = wk[idx] * V[idx] * vn
            if pw != 0:
                img[idx, 1] += pw
                img[idx, 0] += bq * pw
    return img


# ------------------------------------------------------------------ HEALPix (RING), 0-based pixels
def hp_ring_info(nside, ring):
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    if ring < nside:
        return 2 * ring * (ring - 1), 4 * ring, True
    if ring <= 3 * nside:
        return ncap + (ring - nside) * 4 * nside, 4 * nside, ((ring - nside) & 1) == 0
    nr = 4 * nside - ring
    return npix - 2 * nr * (nr + 1), 4 * nr, True


def hp_ring2z(nside, ring):
    if ring < nside:
        return 1.0 - ring * ring * (4.0 / (12 * nside * nside))
    if ring <= 3 * nside:
        return (2 * nside - ring) * (2 * nside * (4.0 / (12 * nside * nside)))
    r = 4 * nside - ring
    return r * r * (4.0 / (12 * nside * nside)) - 1.0


def hp_pix_center(nside, pix):
    """(z, phi) of a pixel centre from its ring/in-ring position — independent of pix2ang_ring's sqrt inversion."""
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    if pix < ncap:
        ring = int((1 + math.isqrt(1 + 2 * pix)) // 2)
    elif pix < npix - ncap:
        ring = (pix - ncap) // (4 * nside) + nside
    else:
        r = int((1 + math.isqrt(2 * (npix - pix) - 1)) // 2)
        ring = 4 * nside - r
    start, nr, shifted = hp_ring_info(nside, ring)
    iphi = pix - start  # 0-based in ring
    phi = (iphi + (0.5 if shifted else 0.0)) * 2 * PI / nr
    return hp_ring2z(nside, ring), phi, ring


def hp_brute_disc(nside, theta, phi, radius):
    """Brute force: all pixels whose centre is within `radius` of (theta,phi)."""
    v = np.array([math.sin(theta) * math.cos(phi), math.sin(theta) * math.sin(phi), math.cos(theta)])
    out = []
    for pix in range(12 * nside * nside):
        z, ph, _ = hp_pix_center(nside, pix)
        s = math.sqrt(max(0.0, (1 - z) * (1 + z)))
        c = np.array([s * math.cos(ph), s * math.sin(ph), z])
        if math.acos(max(-1.0, min(1.0, float(v @ c)))) < radius:
            out.append(pix)
    return out


def hp_pix2vec(nside, pix):
    z, ph, _ = hp_pix_center(nside, pix)
    s = math.sqrt(max(0.0, (1.0 - z) * (1.0 + z)))
    return (s * math.cos(ph), s * math.sin(ph), z)


def hp_ang2pix_brute(nside, theta, phi):
    """pixel containing the direction, found as the nearest pixel centre among the pixels of the ring(s) around it —
    independent of ang2pix_ring's index arithmetic.  (HEALPix pixels are not Voronoi cells of their centres; the test
    that uses this mirror keeps the particle directions well inside a pixel, where the two agree.)"""
    v = (math.sin(theta) * math.cos(phi), math.sin(theta) * math.sin(phi), math.cos(theta))
    best, bd = -1, -2.0
    for pix in range(12 * nside * nside):
        c = hp_pix2vec(nside, pix)
        d = v[0] * c[0] + v[1] * c[1] + v[2] * c[2]
        if d > bd:
            best, bd = pix, d
    return best


def healpix_deposit(pos, hsml, m, rho, binq, w, nside, kernel, kdim=2, calc_mean=True, centre_pixels=None):
    """Second, independent restatement of the particle loop of healpix_map (src/healpix_interpolation/main.jl:143-213)
    with calculate_weights / weight_per_index / contributing_area / distance_to_pixel_center
    (pixel_weights.jl:6-140), contributing_pixels (constributing_pixels.jl:7-22: brute-force disc here),
    particle_area_and_depth (main.jl:56-63) and update_image! (main.jl:25-45).  Pure Python loops: small Nside only.
    `centre_pixels[p]` may supply ang2pix of particle p (otherwise the nearest pixel centre is used)."""
    npix = 12 * nside * nside
    amap = np.zeros(npix); wmap = np.zeros(npix)
    ang_pix = math.sqrt(4.0 * PI / npix)
    for p in range(len(hsml)):
        if not calc_mean and binq[p] == 0:
            continue
        x = [float(pos[p, 0]), float(pos[p, 1]), float(pos[p, 2])]
        dX = math.sqrt(x[0] ** 2 + x[1] ** 2 + x[2] ** 2)
        if dX < hsml[p]:
            continue
        proj = math.asin(hsml[p] / dX)
        theta = math.acos(x[2] / dX)
        phi = math.atan2(x[1], x[0])
        if phi < 0:
            phi += 2 * PI
        pix = hp_brute_disc(nside, theta, phi, proj)
        cpix = centre_pixels[p] if centre_pixels is not None else hp_ang2pix_brute(nside, theta, phi)
        if cpix not in pix:
            pix.append(cpix)
        dz = 2.0 * hsml[p]
        area = (m[p] / rho[p]) / dz
        dz /= (ang_pix * dX) ** 2
        hinv = 1.0 / proj
        A = []; wk = []
        n_tot = n_distr = 0
        d_area = d_w = 0.0
        for q in pix:
            c = hp_pix2vec(nside, q)
            dot = x[0] * c[0] + x[1] * c[1] + x[2] * c[2]
            dx = math.acos(min(dot / dX, 1.0))
            u = dx * hinv
            a = max(0.0, min(ang_pix, abs(proj - (dx - 0.5 * ang_pix)))) / ang_pix
            a /= (ang_pix * dX) ** 2
            d_area += a
            n_tot += 1
            if u <= 1:
                k = W(kernel, kdim, u, hinv)
                d_w += k * a
                n_distr += 1
            else:
                k = 0.0
            A.append(a); wk.append(k)
        if d_w == 0.0:
            n_distr = n_tot
            wk = [1.0] * len(pix)
            wpp = n_distr / d_area if d_area != 0 else 1.0
        else:
            wpp = n_distr / d_w
        an = (area / n_distr) * wpp * w[p] * dz
        for q, a, k in zip(pix, A, wk):
            pw = an * k * a
            amap[q] += binq[p] * pw
            wmap[q] += pw
    return amap, wmap
