/*
 * s2g_oracle_exact.c — EXTENDED-PRECISION ARBITER for the HEALPix particle loop.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as s2g_oracle.c: only tests/, smoke() and bench.py's CPU legs).
 *
 * Why it exists.  weight_per_index (src/healpix_interpolation/pixel_weights.jl:34-76) measures the angle between the
 * particle and a pixel centre as  dx = acos(min(pos·c/Δx, 1))  (distance_to_pixel_center, :16-22).  In Float64 that
 * expression carries a relative error of about ε/dx² (acos is ill-conditioned at 1): 1e-8 for dx ~ 1e-4 rad.  The
 * literal Float64 restatement in s2g_oracle.c inherits that error, so a comparison "GPU vs Float64 oracle" cannot say
 * which side is right when the two differ by 1e-8.  This file evaluates THE SAME FORMULAS in extended precision:
 *
 *   mode 1 (long double, 64-bit mantissa, ε = 5.4e-20; fast enough for 1e9 pixel evaluations):
 *       pixel centre from the exact ring geometry (z, sinθ without the acos round trip, φ = (iφ-off)·π/den),
 *       p̂ = pos/Δx, dx = 2·asinl(|p̂-ĉ|/2)   (same angle as acos(p̂·ĉ), conditioned like dx itself)
 *   mode 2 (__float128, ε = 1e-34; slow, for small samples):
 *       the reference's expression LITERALLY:  dx = acosq(min((pos·c)/Δx, 1))  with the exact pixel centre.
 *   The two agree to ~1e-18 (tests/test_oracle_healpix.py), i.e. both are "the exact value" at the 1e-10 bar.
 *
 * Everything that is integer work stays the Float64 algorithm of s2g_oracle.c (the pixel list of
 * contributing_pixels = query_disc ∪ centre pixel is taken from there, bit for bit), so that the arbiter answers one
 * question only: what are the weights of THOSE pixels in exact arithmetic.  The `u <= 1` test and the sums of
 * calculate_weights (pixel_weights.jl:87-140) run in the extended type as well.
 *
 * Build: oracle/Makefile links this into libs2g_oracle.so (needs -lquadmath).
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define S2GO_API __attribute__((visibility("default")))

int64_t s2go_hp_contributing_pixels(int64_t nside, const double pos[3], double radius, int64_t* out, int64_t cap);

typedef long double ld;
static const ld PI_L = 3.14159265358979323846264338327950288L;
#define PI_Q 3.14159265358979323846264338327950288419716939937510Q

/* ring (1..4nside-1) and 1-based in-ring index of a RING pixel; exact integer arithmetic */
static void pix_ring_iphi(int64_t nside, int64_t pix, int64_t* ring, int64_t* iphi, int64_t* nr, int* off_half);
void s2go_exact_ring_of(int64_t nside, int64_t pix, int64_t* ring, int64_t* iphi, int64_t* nr, int* off_half)
{
    pix_ring_iphi(nside, pix, ring, iphi, nr, off_half);
}
static void pix_ring_iphi(int64_t nside, int64_t pix, int64_t* ring, int64_t* iphi, int64_t* nr, int* off_half)
{
    const int64_t npix = 12 * nside * nside, ncap = 2 * nside * (nside - 1), nl4 = 4 * nside;
    if (pix < ncap) {
        int64_t r = (int64_t)((1.0 + sqrt(1.0 + 2.0 * (double)pix)) / 2.0);
        while (2 * r * (r - 1) > pix) --r;
        while (2 * r * (r + 1) <= pix) ++r;
        *ring = r; *iphi = pix - 2 * r * (r - 1) + 1; *nr = 4 * r; *off_half = 1;
    } else if (pix < npix - ncap) {
        const int64_t ip = pix - ncap;
        *ring = ip / nl4 + nside; *iphi = ip % nl4 + 1; *nr = nl4;
        *off_half = ((*ring + nside) & 1) ? 0 : 1; /* fodd = 0.5*(1 + ((ring+nside)&1)): 0.5 when even, 1 when odd */
    } else {
        const int64_t rem = npix - 1 - pix;
        int64_t r = (int64_t)((1.0 + sqrt(1.0 + 2.0 * (double)rem)) / 2.0);
        while (2 * r * (r - 1) > rem) --r;
        while (2 * r * (r + 1) <= rem) ++r;
        *ring = nl4 - r; *nr = 4 * r; *off_half = 1;
        *iphi = pix - (npix - 2 * r * (r + 1)) + 1;
    }
}

/* walker over a pixel list: consecutive pixels of one ring share (sinθ, cosθ) and step the azimuth by a rotation
 * (exact sinl/cosl re-seed every 32 steps: the recurrence adds < 1e-18); anything else is computed from scratch */
typedef struct {
    int64_t nside, last_pix, ring, nr;
    int steps;
    ld st, ct, cphi, sphi, cd, sd;
} pixwalk;

static void pix2vec_l(int64_t nside, int64_t pix, ld c[3]);

static void walk_init(pixwalk* w, int64_t nside)
{
    w->nside = nside; w->last_pix = -10; w->steps = 0; w->ring = -1; w->nr = 0;
    w->st = w->ct = w->cphi = w->sphi = w->cd = w->sd = 0.0L;
}

static void walk_vec(pixwalk* w, int64_t pix, ld c[3])
{
    if (pix == w->last_pix + 1 && w->steps < 32) {
        /* same ring?  the next pixel number stays in the ring unless last_pix was the ring's last pixel */
        int64_t ring, iphi, nr; int oh;
        (void)oh;
        /* cheap test: recompute ring only at ring boundaries, detected through the cached ring extent */
        if (w->ring >= 0 && pix < w->nr /* nr holds the first pixel of the NEXT ring here */) {
            const ld cn = w->cphi * w->cd - w->sphi * w->sd, sn = w->sphi * w->cd + w->cphi * w->sd;
            w->cphi = cn; w->sphi = sn; w->last_pix = pix; w->steps++;
            c[0] = w->st * cn; c[1] = w->st * sn; c[2] = w->ct;
            return;
        }
        (void)ring; (void)iphi; (void)nr;
    }
    /* from scratch */
    int64_t ring, iphi, nr; int oh;
    extern void s2go_exact_ring_of(int64_t, int64_t, int64_t*, int64_t*, int64_t*, int*);
    s2go_exact_ring_of(w->nside, pix, &ring, &iphi, &nr, &oh);
    pix2vec_l(w->nside, pix, c);
    const ld n = (ld)w->nside;
    (void)n;
    w->ct = c[2];
    w->st = sqrtl(c[0] * c[0] + c[1] * c[1]);
    if (w->st > 0) { w->cphi = c[0] / w->st; w->sphi = c[1] / w->st; } else { w->cphi = 1; w->sphi = 0; }
    const ld dphi = 2.0L * 3.14159265358979323846264338327950288L / (ld)nr;
    w->cd = cosl(dphi); w->sd = sinl(dphi);
    w->ring = ring;
    w->nr = pix - (iphi - 1) + nr;   /* first pixel of the next ring */
    w->last_pix = pix; w->steps = 0;
}

/* exact pixel centre (unit vector), long double */
static void pix2vec_l(int64_t nside, int64_t pix, ld c[3])
{
    int64_t ring, iphi, nr; int oh;
    pix_ring_iphi(nside, pix, &ring, &iphi, &nr, &oh);
    const ld n = (ld)nside;
    ld ct, st;
    if (ring < nside) {
        const ld omz = (ld)(ring * ring) / (3.0L * n * n); /* 1 - z, no cancellation */
        ct = 1.0L - omz; st = sqrtl(omz * (2.0L - omz));
    } else if (ring <= 3 * nside) {
        ct = (ld)(2 * nside - ring) / (1.5L * n); st = sqrtl((1.0L - ct) * (1.0L + ct));
    } else {
        const int64_t rs = 4 * nside - ring;
        const ld opz = (ld)(rs * rs) / (3.0L * n * n);
        ct = opz - 1.0L; st = sqrtl(opz * (2.0L - opz));
    }
    /* phi = (iphi - off) * pi / (nr/2),  off = 0.5 (shifted ring) or 1 */
    const ld phi = ((ld)iphi - (oh ? 0.5L : 1.0L)) * PI_L / ((ld)nr * 0.5L);
    c[0] = st * cosl(phi); c[1] = st * sinl(phi); c[2] = ct;
}

static void pix2vec_q(int64_t nside, int64_t pix, __float128 c[3])
{
    int64_t ring, iphi, nr; int oh;
    pix_ring_iphi(nside, pix, &ring, &iphi, &nr, &oh);
    const __float128 n = (__float128)nside;
    __float128 ct, st;
    if (ring < nside) {
        const __float128 omz = (__float128)(ring * ring) / (3.0Q * n * n);
        ct = 1.0Q - omz; st = sqrtq(omz * (2.0Q - omz));
    } else if (ring <= 3 * nside) {
        ct = (__float128)(2 * nside - ring) / (1.5Q * n); st = sqrtq((1.0Q - ct) * (1.0Q + ct));
    } else {
        const int64_t rs = 4 * nside - ring;
        const __float128 opz = (__float128)(rs * rs) / (3.0Q * n * n);
        ct = opz - 1.0Q; st = sqrtq(opz * (2.0Q - opz));
    }
    const __float128 phi = ((__float128)iphi - (oh ? 0.5Q : 1.0Q)) * PI_Q / ((__float128)nr * 0.5Q);
    c[0] = st * cosq(phi); c[1] = st * sinq(phi); c[2] = ct;
}

/* angle between the particle direction and the centre of `pix` in three ways (for the conditioning study):
 *   out[0] long double, chord form;  out[1] __float128, the reference's acos(min(d/r,1)) literally;
 *   out[2] the same acos expression evaluated in long double (shows how much of long double the acos form eats) */
S2GO_API void s2go_hp_angdist_exact(int64_t nside, int64_t pix, const double pos[3], double out[3], long double* out_ld)
{
    ld c[3];
    pix2vec_l(nside, pix, c);
    const ld Dx = sqrtl((ld)pos[0] * pos[0] + (ld)pos[1] * pos[1] + (ld)pos[2] * pos[2]);
    ld ch2 = 0.0L, dot = 0.0L;
    for (int k = 0; k < 3; k++) { const ld e = (ld)pos[k] / Dx - c[k]; ch2 += e * e; dot += (ld)pos[k] * c[k]; }
    const ld hc = 0.5L * sqrtl(ch2);
    const ld dx_l = 2.0L * asinl(hc < 1.0L ? hc : 1.0L);
    __float128 cq[3];
    pix2vec_q(nside, pix, cq);
    __float128 d = 0.0Q, r2 = 0.0Q;
    for (int k = 0; k < 3; k++) { d += (__float128)pos[k] * cq[k]; r2 += (__float128)pos[k] * (__float128)pos[k]; }
    __float128 t = d / sqrtq(r2);
    if (t > 1.0Q) t = 1.0Q;
    const __float128 dx_q = acosq(t);
    ld tl = dot / Dx;
    if (tl > 1.0L) tl = 1.0L;
    out[0] = (double)dx_l; out[1] = (double)dx_q; out[2] = (double)acosl(tl);
    if (out_ld) { out_ld[0] = dx_l; out_ld[1] = (ld)dx_q; out_ld[2] = acosl(tl); }
}

static inline ld ipow(ld t, int n)
{
    ld r = 1.0L;
    for (int i = 0; i < n; i++) r *= t;
    return r;
}

static ld shape_l(int kid, ld u)
{
    if (!(u < 1.0L)) return 0.0L;
    const ld t = 1.0L - u;
    switch (kid) {
    case 0: return u < 0.5L ? 1.0L + 6.0L * (u - 1.0L) * (u * u) : 2.0L * (t * t * t);
    case 1: {
        ld b = 2.0L / 3.0L - u, c = 1.0L / 3.0L - u;
        b = b > 0 ? b : 0; c = c > 0 ? c : 0;
        return ipow(t, 5) - 6.0L * ipow(b, 5) + 15.0L * ipow(c, 5);
    }
    case 2: return ipow(t, 4) * (1.0L + 4.0L * u);
    case 3: return ipow(t, 6) * (1.0L + 6.0L * u + (35.0L / 3.0L) * u * u);
    case 4: return ipow(t, 8) * (1.0L + 8.0L * u + 25.0L * u * u + 32.0L * u * u * u);
    case 5: return ipow(t, 10) * (5.0L + 50.0L * u + 210.0L * u * u + 450.0L * u * u * u + 429.0L * u * u * u * u);
    }
    return 0.0L;
}

/* `sens_map` (optional, may be NULL; 2*npix doubles: weight map first, then quantity map): Σ over the contributions to a pixel of |∂(pix_weight)/∂dx| · 1 rad, i.e. how much
 * the exact weight-map value moves per radian of error in the angular distances.  Multiplied by the resolution of
 * Float64 unit vectors (a few ulp of 1 = a few 1e-16 rad) it is the part of a pixel value that NO Float64 evaluation
 * can resolve (kernel-rim contributions (1-u)^k -> 0 and A -> 0 edges have unbounded relative sensitivity); the
 * parity tests add exactly that term to the 1e-10 bar instead of a global absolute floor. */
S2GO_API int s2go_healpix_deposit_exact(const double* pos, const double* hsml, const double* m, const double* rho,
                                        const double* binq, const double* w, int64_t n, int64_t nside, int kid,
                                        int calc_mean, int n_workers, double* allsky_map, double* weight_map,
                                        int64_t* stats4, double* sens_map)
{
    const int64_t npix = 12 * nside * nside;
    const ld ang_pix = sqrtl(4.0L * PI_L / (ld)npix);
    int64_t s_mapped = 0, s_foot = 0, s_fb = 0;
    if (n_workers < 1) n_workers = 1;
#pragma omp parallel num_threads(n_workers) reduction(+ : s_mapped, s_foot, s_fb)
    {
        int64_t cap = 4096;
        int64_t* pixidx = (int64_t*)malloc(sizeof(int64_t) * (size_t)cap);
        ld* wk = (ld*)malloc(sizeof(ld) * (size_t)cap);
        ld* A = (ld*)malloc(sizeof(ld) * (size_t)cap);
        ld* D = (ld*)malloc(sizeof(ld) * (size_t)cap);
#pragma omp for schedule(dynamic, 16)
        for (int64_t ip = 0; ip < n; ip++) {
            if (!calc_mean && binq[ip] == 0.0) continue;
            const double* P = pos + 3 * ip;
            /* the skip test and the pixel list use the reference's Float64 values (discrete decisions) */
            double dx2 = 0.0;
            for (int d = 0; d < 3; d++) dx2 += P[d] * P[d];
            const double Dx64 = sqrt(dx2);
            if (Dx64 < hsml[ip]) continue;
            const double proj64 = asin(hsml[ip] / Dx64);
            int64_t np;
            for (;;) {
                np = s2go_hp_contributing_pixels(nside, P, proj64, pixidx, cap);
                if (np >= 0) break;
                cap = -np + 16;
                pixidx = (int64_t*)realloc(pixidx, sizeof(int64_t) * (size_t)cap);
                wk = (ld*)realloc(wk, sizeof(ld) * (size_t)cap);
                A = (ld*)realloc(A, sizeof(ld) * (size_t)cap);
                D = (ld*)realloc(D, sizeof(ld) * (size_t)cap);
            }
            const ld Dx = sqrtl((ld)P[0] * P[0] + (ld)P[1] * P[1] + (ld)P[2] * P[2]);
            const ld ph = asinl((ld)hsml[ip] / Dx), hinv = 1.0L / ph;
            const ld ux = (ld)P[0] / Dx, uy = (ld)P[1] / Dx, uz = (ld)P[2] / Dx;
            ld dz = 2.0L * (ld)hsml[ip];
            const ld area = ((ld)m[ip] / (ld)rho[ip]) / dz;
            const ld aD = ang_pix * Dx;
            dz /= aD * aD;
            int64_t n_distr = 0, n_tot = 0;
            ld dw = 0.0L, da = 0.0L;
            pixwalk wlk;
            walk_init(&wlk, nside);
            for (int64_t k = 0; k < np; k++) {
                ld c[3];
                walk_vec(&wlk, pixidx[k], c);
                const ld ex = ux - c[0], ey = uy - c[1], ez = uz - c[2];
                const ld hc = 0.5L * sqrtl(ex * ex + ey * ey + ez * ez);
                const ld ddx = 2.0L * asinl(hc < 1.0L ? hc : 1.0L);
                const ld u = ddx * hinv;
                ld inner = fabsl(ph - (ddx - 0.5L * ang_pix));
                ld mn = ang_pix < inner ? ang_pix : inner;
                ld a_ = (mn > 0.0L ? mn : 0.0L) / ang_pix;
                a_ /= aD * aD;
                da += a_; n_tot++;
                ld w_ = 0.0L;
                if (u <= 1.0L) { w_ = shape_l(kid, u); dw += w_ * a_; n_distr++; } /* the kernel norm cancels (Q14) */
                A[k] = a_; wk[k] = w_;
                if (sens_map) D[k] = ddx;
            }
            ld* dA = NULL; ld* dW = NULL;
            if (sens_map) { /* second sweep: derivatives of A and w with respect to dx (cheap: sens is a test aid) */
                dA = (ld*)malloc(sizeof(ld) * (size_t)np); dW = (ld*)malloc(sizeof(ld) * (size_t)np);
                for (int64_t k = 0; k < np; k++) {
                    const ld ddx = D[k];
                    const ld u = ddx * hinv, h = 1e-7L;
                    const ld inner = fabsl(ph - (ddx - 0.5L * ang_pix));
                    dA[k] = (inner < ang_pix) ? 1.0L / ang_pix / (aD * aD) : 0.0L;
                    dW[k] = (u <= 1.0L) ? fabsl(shape_l(kid, u + h) - shape_l(kid, u - h > 0 ? u - h : 0.0L)) /
                                              (u - h > 0 ? 2.0L * h : u + h) * hinv
                                        : 0.0L;
                }
            }
            ld wpp;
            int fb = 0;
            if (dw == 0.0L) {
                fb = 1; n_distr = n_tot;
                wpp = (da != 0.0L) ? (ld)n_distr / da : 1.0L;
                s_fb++;
            } else
                wpp = (ld)n_distr / dw;
            const ld area_norm = area / (ld)n_distr * wpp * (ld)w[ip] * dz;
            for (int64_t k = 0; k < np; k++) {
                const ld pw = area_norm * (fb ? 1.0L : wk[k]) * A[k];
                const double a1 = (double)((ld)binq[ip] * pw), a2 = (double)pw;
#pragma omp atomic
                allsky_map[pixidx[k]] += a1;
#pragma omp atomic
                weight_map[pixidx[k]] += a2;
                if (sens_map) {
                    const double sv = (double)(fabsl(area_norm) * (fb ? dA[k] : (dW[k] * A[k] + wk[k] * dA[k])));
#pragma omp atomic
                    sens_map[pixidx[k]] += sv;
                    const double sq = fabs(binq[ip]) * sv;
#pragma omp atomic
                    sens_map[npix + pixidx[k]] += sq;
                }
            }
            free(dA); free(dW);
            s_mapped++; s_foot += np;
        }
        free(pixidx); free(wk); free(A); free(D);
    }
    if (stats4) { stats4[0] = s_mapped; stats4[1] = s_foot; stats4[2] = s_foot; stats4[3] = s_fb; }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Float64 evaluation of the SAME weights with the chord formulation the CUDA kernels use
 * (csrc/s2g_healpix.cu: pixel centre from the ring geometry without acos, dx = 2 asin(|p̂-ĉ|/2)).
 * Only for the conditioning study in tests/test_oracle_healpix.py: it shows on the CPU, without a GPU, that the chord
 * form in Float64 stays within 1e-10 of the extended-precision value where the literal acos form is off by 1e-8.
 * ------------------------------------------------------------------------------------------------------------ */
static void pix2vec_chord64(int64_t nside, int64_t pix, double c[3])
{
    int64_t ring, iphi, nr; int oh;
    pix_ring_iphi(nside, pix, &ring, &iphi, &nr, &oh);
    const double n = (double)nside;
    double ct, st;
    if (ring < nside) {
        const double omz = (double)(ring * ring) / (3.0 * n * n);
        ct = 1.0 - omz; st = sqrt(omz * (2.0 - omz));
    } else if (ring <= 3 * nside) {
        ct = (double)(2 * nside - ring) / (1.5 * n); st = sqrt((1.0 - ct) * (1.0 + ct));
    } else {
        const int64_t rs = 4 * nside - ring;
        const double opz = (double)(rs * rs) / (3.0 * n * n);
        ct = opz - 1.0; st = sqrt(opz * (2.0 - opz));
    }
    const double phi = ((double)iphi - (oh ? 0.5 : 1.0)) * 3.14159265358979323846 / ((double)nr * 0.5);
    c[0] = st * cos(phi); c[1] = st * sin(phi); c[2] = ct;
}

static double shape_d(int kid, double u) { return (double)shape_l(kid, (ld)u); }

S2GO_API int s2go_healpix_deposit_chord64(const double* pos, const double* hsml, const double* m, const double* rho,
                                          const double* binq, const double* w, int64_t n, int64_t nside, int kid,
                                          int calc_mean, double* allsky_map, double* weight_map)
{
    const int64_t npix = 12 * nside * nside;
    const double ang_pix = sqrt(4.0 * 3.14159265358979323846 / (double)npix);
    int64_t cap = 4096;
    int64_t* pixidx = (int64_t*)malloc(sizeof(int64_t) * (size_t)cap);
    double* wk = (double*)malloc(sizeof(double) * (size_t)cap);
    double* A = (double*)malloc(sizeof(double) * (size_t)cap);
    for (int64_t ip = 0; ip < n; ip++) {
        if (!calc_mean && binq[ip] == 0.0) continue;
        const double* P = pos + 3 * ip;
        double dx2 = 0.0;
        for (int d = 0; d < 3; d++) dx2 += P[d] * P[d];
        const double Dx = sqrt(dx2);
        if (Dx < hsml[ip]) continue;
        const double ph = asin(hsml[ip] / Dx), hinv = 1.0 / ph;
        int64_t np;
        for (;;) {
            np = s2go_hp_contributing_pixels(nside, P, ph, pixidx, cap);
            if (np >= 0) break;
            cap = -np + 16;
            pixidx = (int64_t*)realloc(pixidx, sizeof(int64_t) * (size_t)cap);
            wk = (double*)realloc(wk, sizeof(double) * (size_t)cap);
            A = (double*)realloc(A, sizeof(double) * (size_t)cap);
        }
        const double ux = P[0] / Dx, uy = P[1] / Dx, uz = P[2] / Dx;
        double dz = 2.0 * hsml[ip];
        const double area = (m[ip] / rho[ip]) / dz;
        const double aD = ang_pix * Dx;
        dz /= aD * aD;
        int64_t n_distr = 0, n_tot = 0;
        double dw = 0.0, da = 0.0;
        for (int64_t k = 0; k < np; k++) {
            double c[3];
            pix2vec_chord64(nside, pixidx[k], c);
            const double ex = c[0] - ux, ey = c[1] - uy, ez = c[2] - uz;
            const double hc = 0.5 * sqrt(ex * ex + ey * ey + ez * ez);
            const double ddx = 2.0 * asin(hc < 1.0 ? hc : 1.0);
            const double u = ddx * hinv;
            double inner = fabs(ph - (ddx - 0.5 * ang_pix));
            double mn = ang_pix < inner ? ang_pix : inner;
            double a_ = (mn > 0.0 ? mn : 0.0) / ang_pix / (aD * aD);
            da += a_; n_tot++;
            double w_ = 0.0;
            if (u <= 1.0) { w_ = shape_d(kid, u); dw += w_ * a_; n_distr++; }
            A[k] = a_; wk[k] = w_;
        }
        double wpp;
        int fb = 0;
        if (dw == 0.0) { fb = 1; n_distr = n_tot; wpp = (da != 0.0) ? (double)n_distr / da : 1.0; }
        else wpp = (double)n_distr / dw;
        const double area_norm = area / (double)n_distr * wpp * w[ip] * dz;
        for (int64_t k = 0; k < np; k++) {
            const double pw = area_norm * (fb ? 1.0 : wk[k]) * A[k];
            allsky_map[pixidx[k]] += binq[ip] * pw;
            weight_map[pixidx[k]] += pw;
        }
    }
    free(pixidx); free(wk); free(A);
    return 0;
}
