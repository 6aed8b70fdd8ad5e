// s2g_healpix.cuh — HEALPix RING geometry shared by the scatter walk (s2g_healpix.cu) and the tile-gather
// (s2g_hpgather.cu).  The RING arithmetic is the published HEALPix algorithm (ring_above / ring2z / get_ring_info /
// pix2ang_ring / ang2pix_ring / query_disc) that Healpix.jl ports; call sites in the reference:
// src/healpix_interpolation/constributing_pixels.jl:10-16, pixel_weights.jl:42.  Pixel numbers are 0-based.
#pragma once
#include "s2g_common.cuh"

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;

struct HpGeom {
    long long nside, npix, ncap, nl2, nl4;
    double fact1_r2z, fact2_r2z;  // ring2z:   fact2 = 4/npix, fact1 = 2*nside*fact2
    double fact1_p2a, fact2_p2a;  // pix2ang:  fact1 = 1.5*nside, fact2 = 3*nside^2
    double ang_pix;               // sqrt(4π/npix)  (main.jl:144)
};

__host__ __device__ inline HpGeom make_hp(long long nside)
{
    HpGeom g;
    g.nside = nside;
    g.npix = 12 * nside * nside;
    g.ncap = 2 * nside * (nside - 1);
    g.nl2 = 2 * nside;
    g.nl4 = 4 * nside;
    g.fact2_r2z = 4.0 / (double)g.npix;
    g.fact1_r2z = (double)(2 * nside) * g.fact2_r2z;
    g.fact1_p2a = 1.5 * (double)nside;
    g.fact2_p2a = 3.0 * (double)nside * (double)nside;
    g.ang_pix = sqrt(4.0 * kPi / (double)g.npix);
    return g;
}

__device__ __forceinline__ void hp_ring_info(const HpGeom& g, long long ring, long long& startpix, long long& ringpix,
                                             bool& shifted)
{
    if (ring < g.nside) {
        ringpix = 4 * ring; startpix = 2 * ring * (ring - 1); shifted = true;
    } else if (ring <= 3 * g.nside) {
        ringpix = g.nl4; startpix = g.ncap + (ring - g.nside) * g.nl4; shifted = (((ring - g.nside) & 1) == 0);
    } else {
        const long long nr = g.nl4 - ring;
        ringpix = 4 * nr; startpix = g.npix - 2 * nr * (nr + 1); shifted = true;
    }
}

__device__ __forceinline__ long long hp_ring_above(const HpGeom& g, double z)
{
    const double az = fabs(z);
    if (az <= 2.0 / 3.0) return (long long)__dmul_rn((double)g.nside, __dadd_rn(2.0, -__dmul_rn(1.5, z)));
    const long long iring = (long long)__dmul_rn((double)g.nside, __dsqrt_rn(__dmul_rn(3.0, __dadd_rn(1.0, -az))));
    return (z > 0) ? iring : g.nl4 - iring - 1;
}

__device__ __forceinline__ double hp_ring2z(const HpGeom& g, long long ring)
{
    if (ring < g.nside) return __dadd_rn(1.0, -__dmul_rn((double)(ring * ring), g.fact2_r2z));
    if (ring <= 3 * g.nside) return __dmul_rn((double)(g.nl2 - ring), g.fact1_r2z);
    ring = g.nl4 - ring;
    return __dadd_rn(__dmul_rn((double)(ring * ring), g.fact2_r2z), -1.0);
}

__device__ __forceinline__ long long hp_ang2pix_ring(const HpGeom& g, double theta, double phi)
{
    const double z = cos(theta), za = fabs(z);
    double tt = fmod(phi, kTwoPi);
    if (tt < 0) tt = __dadd_rn(tt, kTwoPi);
    tt = __ddiv_rn(tt, 0.5 * kPi);
    if (za <= 2.0 / 3.0) {
        const double temp1 = __dmul_rn((double)g.nside, __dadd_rn(0.5, tt));
        const double temp2 = __dmul_rn(__dmul_rn((double)g.nside, z), 0.75);
        const long long jp = (long long)floor(__dadd_rn(temp1, -temp2));
        const long long jm = (long long)floor(__dadd_rn(temp1, temp2));
        const long long ir = g.nside + 1 + jp - jm;
        const long long kshift = 1 - (ir & 1);
        long long ip = (jp + jm - g.nside + kshift + 1) / 2;
        ip = ((ip % g.nl4) + g.nl4) % g.nl4;
        return g.ncap + (ir - 1) * g.nl4 + ip;
    }
    const double tp = __dadd_rn(tt, -floor(tt));
    const double tmp = __dmul_rn((double)g.nside, __dsqrt_rn(__dmul_rn(3.0, __dadd_rn(1.0, -za))));
    const long long jp = (long long)floor(__dmul_rn(tp, tmp));
    const long long jm = (long long)floor(__dmul_rn(__dadd_rn(1.0, -tp), tmp));
    const long long ir = jp + jm + 1;
    long long ip = (long long)floor(__dmul_rn(tt, (double)ir));
    ip = ((ip % (4 * ir)) + 4 * ir) % (4 * ir);
    if (z > 0) return 2 * ir * (ir - 1) + ip;
    return g.npix - 2 * ir * (ir + 1) + ip;
}

// colatitude-dependent part of pix2ang_ring for a whole ring: theta = acos(z_ring), returns sin/cos(theta) and the
// azimuth step so that phi(j) = (j + 1 - off) * kPi / den  for the 0-based in-ring index j
struct RingTrig {
    double st, ct, off, inv_den;
};
__device__ __forceinline__ RingTrig hp_ring_trig(const HpGeom& g, long long ring)
{
    RingTrig t;
    double z, den;
    // pix2vecRing = (sinθ cosφ, sinθ sinφ, cosθ) with θ = acos(z): cosθ = z and sinθ = sqrt((1-z)(1+z)) without the
    // round trip through acos.  In the polar caps 1-|z| = ring²/(3 Nside²) is taken directly (no cancellation: next
    // to a pole 1-z ~ 1e-7 and 1 - (1 - x) would keep only 9 digits of sinθ); DESIGN.md "HEALPix conditioning".
    if (ring < g.nside) {
        const double omz = __ddiv_rn((double)(ring * ring), g.fact2_p2a);
        z = 1.0 - omz;
        t.st = sqrt(omz * (2.0 - omz));
        t.off = 0.5; den = __dmul_rn(2.0, (double)ring);
    } else if (ring <= 3 * g.nside) {
        z = __ddiv_rn((double)(g.nl2 - ring), g.fact1_p2a);
        t.st = sqrt((1.0 - z) * (1.0 + z));
        t.off = 0.5 * (double)(1 + ((ring + g.nside) & 1));
        den = __dmul_rn(2.0, (double)g.nside);
    } else {
        const long long rs = g.nl4 - ring;
        const double opz = __ddiv_rn((double)(rs * rs), g.fact2_p2a);
        z = opz - 1.0;
        t.st = sqrt(opz * (2.0 - opz));
        t.off = 0.5; den = __dmul_rn(2.0, (double)rs);
    }
    t.ct = z;
    t.inv_den = 1.0 / den;
    return t;
}

// per-particle disc description
struct Disc {
    double px, py, pz, Dx;        // position relative to the observer and its norm (shared.jl:1-10)
    double proj_h, hinv;          // asin(h/Dx) and its inverse
    double theta, phi;            // vec2ang
    double z0, xa, cosr;          // query_disc constants
    long long ring_first, ring_last;      // all rings walked (cap rings + disc rings)
    long long irmin, irmax;               // disc rings (others are full cap rings)
    long long cpix;                       // pixel containing the centre (ang2pix)
    bool full_sky;
    double ux, uy, uz;            // unit vector to the particle
    double inv_ang, inv_aD2;      // 1/ang_pix, 1/(ang_pix*Dx)^2
    bool small;                   // disc (+ one pixel) below 0.2 rad: asin by its series
};

__device__ __forceinline__ void make_disc(const HpGeom& g, Disc& d)
{
    // vec2ang (Healpix.jl): theta = acos(z/norm), phi = atan(y,x) (+2π if negative)
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d.px, d.px), __dmul_rn(d.py, d.py)), __dmul_rn(d.pz, d.pz)));
    d.theta = acos(__ddiv_rn(d.pz, nrm));
    double ph = atan2(d.py, d.px);
    if (ph < 0) ph = __dadd_rn(ph, kTwoPi);
    d.phi = ph;
    d.cpix = hp_ang2pix_ring(g, d.theta, d.phi);
    const double r = d.proj_h;
    d.full_sky = (r >= kPi);
    if (d.full_sky) {
        d.ring_first = 1; d.ring_last = g.nl4 - 1; d.irmin = g.nl4; d.irmax = 0;
        return;
    }
    d.cosr = cos(r);
    d.z0 = cos(d.theta);
    d.xa = __ddiv_rn(1.0, __dsqrt_rn(__dmul_rn(__dadd_rn(1.0, -d.z0), __dadd_rn(1.0, d.z0))));
    const double rlat1 = __dadd_rn(d.theta, -r);
    d.irmin = hp_ring_above(g, cos(rlat1)) + 1;
    d.ring_first = d.irmin;
    if ((rlat1 <= 0) && (d.irmin > 1)) d.ring_first = 1;  // north pole inside the disc: rings 1..irmin-1 entirely
    const double rlat2 = __dadd_rn(d.theta, r);
    d.irmax = hp_ring_above(g, cos(rlat2));
    d.ring_last = d.irmax;
    if ((rlat2 >= kPi) && (d.irmax + 1 < g.nl4)) d.ring_last = g.nl4 - 1;  // south pole inside the disc
}

// pixel run of one ring: start index (0-based in ring, may need mod) and count (0 = ring not touched)
__device__ __forceinline__ void ring_run(const HpGeom& g, const Disc& d, long long ring, long long nr, bool shifted,
                                         long long& j0, long long& cnt)
{
    if (d.full_sky || ring < d.irmin || ring > d.irmax) {  // cap ring: the whole ring
        j0 = 0; cnt = nr;
        return;
    }
    const double z = hp_ring2z(g, ring);
    const double x = __dmul_rn(__dadd_rn(d.cosr, -__dmul_rn(z, d.z0)), d.xa);
    const double ysq = __dadd_rn(__dadd_rn(1.0, -__dmul_rn(z, z)), -__dmul_rn(x, x));
    const double dphi = (ysq <= 0) ? 0.0 : atan2(__dsqrt_rn(ysq), x);
    if (!(dphi > 0)) { j0 = 0; cnt = 0; return; }
    const double shift = shifted ? 0.5 : 0.0;
    const double f = __ddiv_rn((double)nr, kTwoPi);
    long long ip_lo = (long long)floor(__dadd_rn(__dmul_rn(f, __dadd_rn(d.phi, -dphi)), -shift)) + 1;
    long long ip_hi = (long long)floor(__dadd_rn(__dmul_rn(f, __dadd_rn(d.phi, dphi)), -shift));
    if (ip_hi >= nr) { ip_lo -= nr; ip_hi -= nr; }
    long long c = ip_hi - ip_lo + 1;
    if (c <= 0) { j0 = 0; cnt = 0; return; }
    if (c > nr) c = nr;  // the reference de-duplicates with unique! (constributing_pixels.jl:19)
    j0 = ip_lo < 0 ? ip_lo + nr : ip_lo;
    cnt = c;
}

// weight_per_index (pixel_weights.jl:34-76) for the pixel at in-ring index j of a ring with trig constants rt.
// The angular distance to the pixel centre is evaluated from the CHORD between the two unit vectors,
// dx = 2 asin(|p̂ - ĉ| / 2), instead of the reference's acos(min(p·c/Δx, 1)): same angle, but well conditioned at the
// sub-degree separations that matter here (acos loses ε/dx² relative accuracy) and without a division or acos call.
template <int KID>
__device__ __forceinline__ void pixel_weight(const HpGeom& g, const Disc& d, const RingTrig& rt, long long j, double& A,
                                             double& wk, bool& inside)
{
    double sp, cp;
    sincospi(((double)(j + 1) - rt.off) * rt.inv_den, &sp, &cp);  // phi = (iphi - off) * pi / den
    const double ex = fma(rt.st, cp, -d.ux), ey = fma(rt.st, sp, -d.uy), ez = rt.ct - d.uz;
    const double c2 = fma(ex, ex, fma(ey, ey, ez * ez));
    const double hc = 0.5 * sqrt(c2);  // half chord = sin(dx/2)
    double dx;
    if (d.small) {
        const double x2 = hc * hc;
        double pser = fma(x2, 135135.0 / 9676800.0, 10395.0 / 599040.0);
        pser = fma(pser, x2, 945.0 / 42240.0);
        pser = fma(pser, x2, 105.0 / 3456.0);
        pser = fma(pser, x2, 15.0 / 336.0);
        pser = fma(pser, x2, 3.0 / 40.0);
        pser = fma(pser, x2, 1.0 / 6.0);
        dx = 2.0 * fma(hc * x2, pser, hc);
    } else
        dx = 2.0 * asin(fmin(hc, 1.0));
    const double u = dx * d.hinv;
    // contributing_area (pixel_weights.jl:6-8) then / (ang_pix*Dx)^2 (:53)
    const double inner = fabs(d.proj_h - (dx - 0.5 * g.ang_pix));
    A = fmax(0.0, fmin(g.ang_pix, inner)) * d.inv_ang * d.inv_aD2;
    inside = (u <= 1.0);
    wk = inside ? kernel_shape<KID>(u) : 0.0;
}

// ---- fast ring walk ---------------------------------------------------------------------------------------
// polynomial coefficients live in the constant bank so that they are FP64-instruction operands (c[bank][off]) instead of
// being re-materialised with UMOV/IMAD.MOV pairs inside the pixel loop
__constant__ double kAsinC[7] = {1.0 / 6.0, 3.0 / 40.0, 15.0 / 336.0, 105.0 / 3456.0, 945.0 / 42240.0,
                                 10395.0 / 599040.0, 135135.0 / 9676800.0};
__constant__ double kRsq[2] = {0.375, 0.5};
__constant__ double kWC4[3] = {56.0 / 3.0, -88.0 / 3.0, 35.0 / 3.0};
__constant__ double kWC6[4] = {66.0, -154.0, 121.0, -32.0};
// 1/sqrt(s), s > 0, ~2 ulp (MUFU.RSQ64H seed + one third-order Newton step)
__device__ __forceinline__ double hp_rsqrt(double s)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double e = fma(-s, y * y, 1.0);
    return fma(fma(e, kRsq[0], kRsq[1]), e * y, y);
}

// kernel shape from t = 1 - u, 0 < t <= 1 (no range test)
template <int KID>
__device__ __forceinline__ double hp_shape_t(double t)
{
    if (KID == S2G_KERNEL_CUBIC) {
        const double a = fma(fma(fma(-6.0, t, 12.0), t, -6.0), t, 1.0), b = 2.0 * (t * t * t);
        return t > 0.5 ? a : b;
    } else if (KID == S2G_KERNEL_QUINTIC) {
        const double b0 = t - 1.0 / 3.0, c0 = t - 2.0 / 3.0;
        const double b = b0 > 0.0 ? b0 : 0.0, c = c0 > 0.0 ? c0 : 0.0;
        const double a2 = t * t, b2 = b * b, c2 = c * c;
        return fma(15.0 * c, c2 * c2, fma(-6.0 * b, b2 * b2, a2 * a2 * t));
    } else if (KID == S2G_KERNEL_WENDLAND_C2) {
        const double t2 = t * t;
        return (t2 * t2) * fma(-4.0, t, 5.0);
    } else if (KID == S2G_KERNEL_WENDLAND_C4) {
        const double t2 = t * t;
        return (t2 * t2 * t2) * fma(fma(kWC4[2], t, kWC4[1]), t, kWC4[0]);
    } else if (KID == S2G_KERNEL_WENDLAND_C6) {
        const double t2 = t * t, t4 = t2 * t2;
        return (t4 * t4) * fma(fma(fma(kWC6[3], t, kWC6[2]), t, kWC6[1]), t, kWC6[0]);
    } else {
        const double t2 = t * t, t4 = t2 * t2, u = 1.0 - t;
        return (t4 * t4 * t2) * fma(fma(fma(fma(429.0, u, 450.0), u, 210.0), u, 50.0), u, 5.0);
    }
}


}  // namespace

// ---- tile-gather (s2g_hpgather.cu): one record per particle, written by the ring-walk pass A (s2g_healpix.cu, REC mode)
struct __align__(16) HRec {
    double ux, uy, uz;   // unit vector from the observer to the particle
    double ph;           // proj_hsml = asin(hsml / Δx)                      (main.jl:182)
    double an, anq;      // area_norm / (ang_pix Δx)² (main.jl:32-33, pixel_weights.jl:53) and the same times Bin_q
    int rmin, rmax;      // rings of the disc; rmin > rmax: record unused (particle handed back to the scatter walk)
    int ntot;            // length of the reference's pixel list (n_tot_pix, pixel_weights.jl:57)
    int pad;
};
static_assert(sizeof(HRec) == 64, "HRec is one 64-byte line");

int s2g_hp_classify(s2g_ctx* ctx, const s2g_particles& P, long long nside, int calc_mean, const unsigned char* take,
                    double heavy_radius, double gather_radius, int gather_on, unsigned char* heavy, unsigned char* gath,
                    unsigned char* gath_heavy, unsigned char* skip);
int s2g_hp_launch_records(s2g_ctx* ctx, const s2g_particles& P, long long nside, int kernel, int calc_mean,
                          const unsigned* list, long long n_list, HRec* recs, unsigned char* skip, int coop);
int s2g_hp_gather_pipeline(s2g_ctx* ctx, const s2g_particles& P, long long nside, int kernel, int calc_mean,
                           const unsigned* list, long long n_list, unsigned char* skip, double* amap, double* wmap,
                           int coop_records, int big, int nt);
