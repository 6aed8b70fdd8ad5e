// s2g_cic2d.cuh — per-particle pixel-space record of the 2D Smac deposit and the exact (bit-for-bit)
// footprint arithmetic shared by the scatter and gather kernels.
#pragma once
#include "s2g_common.cuh"

#ifdef __CUDACC__
struct Rec2 {
    double x, y;        // pixel coordinates (get_xyz, cic_shared.jl:85-100)
    double h, hinv;     // hsml in pixels and its inverse (get_quantities_2D, cic_2D.jl:80-91)
    double area;        // (2h)^2
    double dz;          // m / rho_pix / area
    double w;           // los weight
    int iMin, iMax, jMin, jMax;  // pix_index_min_max (cic_shared.jl:46-52)
    bool all_zero;      // bin_q == 0 for every image
};

// floor(Integer, v) clamped to int range (npix < 2^31 is enforced by the API)
__device__ __forceinline__ int floor_to_int(double v)
{
    long long f = __double2ll_rd(v);
    f = f < -2147483647LL ? -2147483647LL : f;
    f = f > 2147483647LL ? 2147483647LL : f;
    return (int)f;
}

// Builds the record of particle p.  Returns false when the particle does not deposit anything:
//   - bin_q == 0 (all images) and !calc_mean          (cic_2D.jl:166)
//   - fused sphMapping filter rejects its centre       (filter_shift.jl:50-58)
//   - its clipped footprint is empty                   (loops iMin:iMax / jMin:jMax do not execute)
// No FMA contraction anywhere on the path to the integer bounds: __dmul_rn/__dadd_rn are never fused.
__device__ __forceinline__ bool make_rec2(const s2g_particles& P, const s2g_geom& G, long long p, Rec2& r)
{
    bool all_zero = true;
    for (int q = 0; q < G.n_images; ++q)
        if (ld_in(P.binq, (long long)G.n_images * p + q, P.in_dtype) != 0.0) all_zero = false;
    if (all_zero && !G.calc_mean) return false;
    r.all_zero = all_zero;

    const double px = ld_pos(P, p, 0), py = ld_pos(P, p, 1);
    if (P.fuse_center) {
        const double pz = ld_pos(P, p, 2);
        if (!in_image(P, px, py, pz)) return false;
    }
    const double hs = ld_in(P.hsml, p, P.in_dtype);
    const double mm = ld_in(P.m, p, P.in_dtype);
    const double rh = ld_in(P.rho, p, P.in_dtype);
    r.w = ld_in(P.w, p, P.in_dtype);

    r.h = __dmul_rn(hs, G.len2pix);
    r.hinv = __ddiv_rn(1.0, r.h);
    const double h2 = __dmul_rn(2.0, r.h);
    r.area = __dmul_rn(h2, h2);
    const double rho_p = __dmul_rn(rh, G.inv_l3);
    r.dz = __ddiv_rn(__ddiv_rn(mm, rho_p), r.area);

    r.x = __dadd_rn(__dmul_rn(px, G.len2pix), G.half_n);
    r.y = __dadd_rn(__dmul_rn(py, G.len2pix), G.half_n);

    const int n1 = (int)G.npix - 1;
    r.iMin = max(floor_to_int(__dadd_rn(r.x, -r.h)), 0);
    r.iMax = min(floor_to_int(__dadd_rn(r.x, r.h)), n1);
    r.jMin = max(floor_to_int(__dadd_rn(r.y, -r.h)), 0);
    r.jMax = min(floor_to_int(__dadd_rn(r.y, r.h)), n1);
    return (r.iMin <= r.iMax) && (r.jMin <= r.jMax);
}

// get_dxyz (cic_shared.jl:60-62): overlap length of [x-h, x+h] with pixel [i, i+1]
__device__ __forceinline__ double overlap_1d(double x, double h, int i)
{
    const double a = __dadd_rn(x, h), b = (double)(i + 1);
    const double c = __dadd_rn(x, -h), d = (double)i;
    // plain selects (fmin/fmax carry NaN handling that costs ~8 instructions each; a NaN here is garbage anyway)
    return __dadd_rn(a < b ? a : b, -(c > d ? c : d));
}

// x - i - 0.5 evaluated left to right (get_x_dx, cic_shared.jl:73)
__device__ __forceinline__ double center_dist(double x, double id) { return __dadd_rn(__dadd_rn(x, -id), -0.5); }

// u = sqrt(dx*dx + dy*dy) * hinv  (get_d_hsml, distances.jl:6-8), every operation individually rounded so that
// the classification u <= 1 agrees with the reference for every pixel.
__device__ __forceinline__ double u_of(double xd2, double yd2, double hinv)
{
    return __dmul_rn(__dsqrt_rn(__dadd_rn(xd2, yd2)), hinv);
}

// ---- fast path helpers (coordinates in units of h; used where the pixel count is large) -------------------------
// 1/sqrt(s) for s > 0 to ~2 ulp: MUFU.RSQ64H seed + one third-order Newton step (no IEEE fix-up, no slow path)
__device__ __forceinline__ double rsqrt_fast(double s)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double e = fma(-s, y * y, 1.0);
    const double c = fma(e, 0.375, 0.5);
    return fma(c, e * y, y);
}

// kernel shape as a function of t = 1 - u for 0 < t <= 1, WITHOUT range test (callers mask u >= 1).
// The polynomial factor of the Wendland kernels is re-expanded in t (saves forming u); WendlandC8 keeps the u form
// (its t-expansion has alternating coefficients ~4e3 and would lose ~3 digits to cancellation).
template <int KID>
__device__ __forceinline__ double shape_t(double t)
{
    if (KID == S2G_KERNEL_CUBIC) {
        // u < 0.5: 1 + 6(u-1)u^2 = 1 - 6t + 12t^2 - 6t^3 ; else 2 t^3
        const double a = fma(fma(fma(-6.0, t, 12.0), t, -6.0), t, 1.0), b = 2.0 * (t * t * t);
        return t > 0.5 ? a : b;
    } else if (KID == S2G_KERNEL_QUINTIC) {
        const double b = fmax(t - 1.0 / 3.0, 0.0), c = fmax(t - 2.0 / 3.0, 0.0);
        const double a2 = t * t, b2 = b * b, c2 = c * c;
        return fma(15.0 * c, c2 * c2, fma(-6.0 * b, b2 * b2, a2 * a2 * t));
    } else if (KID == S2G_KERNEL_WENDLAND_C2) {
        const double t2 = t * t;
        return (t2 * t2) * fma(-4.0, t, 5.0);
    } else if (KID == S2G_KERNEL_WENDLAND_C4) {
        const double t2 = t * t;
        return (t2 * t2 * t2) * fma(fma(35.0 / 3.0, t, -88.0 / 3.0), t, 56.0 / 3.0);
    } else if (KID == S2G_KERNEL_WENDLAND_C6) {
        const double t2 = t * t, t4 = t2 * t2;
        return (t4 * t4) * fma(fma(fma(-32.0, t, 121.0), t, -154.0), t, 66.0);
    } else {
        const double t2 = t * t, t4 = t2 * t2, u = 1.0 - t;
        return (t4 * t4 * t2) * fma(fma(fma(fma(429.0, u, 450.0), u, 210.0), u, 50.0), u, 5.0);
    }
}

// w(sqrt(s)) for 0 < s < 1 (garbage for s >= 1: mask it): t = 1 - s * rsqrt(s)
template <int KID>
__device__ __forceinline__ double shape_s(double s)
{
    return shape_t<KID>(fma(-s, rsqrt_fast(s), 1.0));
}

// ---- one-constant-per-step forms for the tile-gather inner loop.  A DFMA takes ONE immediate; a Horner step with
// two literal constants (fma(-32, t, 121)) makes ptxas re-materialise the second one with two IMAD.MOVs per group of
// pixels.  Dividing the polynomial factor by its leading coefficient where that is a power of two turns the first
// step into a DADD with an immediate and leaves every other step with a single immediate;
//     shape_t<KID>(t) == shape_scale<KID>() * shape_t_scaled<KID>(t)   bit for bit
// (scaling by 2^k commutes with every rounding), and the caller folds shape_scale into area_norm once per record.
template <int KID>
__host__ __device__ constexpr double shape_scale()
{
    return KID == S2G_KERNEL_WENDLAND_C6 ? 32.0 : (KID == S2G_KERNEL_WENDLAND_C2 ? 4.0 : 1.0);
}

template <int KID>
__device__ __forceinline__ double shape_t_scaled(double t)
{
    if (KID == S2G_KERNEL_WENDLAND_C2) {
        const double t2 = t * t;
        return (t2 * t2) * (1.25 - t);
    } else if (KID == S2G_KERNEL_WENDLAND_C6) {
        const double t2 = t * t, t4 = t2 * t2;
        return (t4 * t4) * fma(fma(3.78125 - t, t, -4.8125), t, 2.0625);
    } else {
        return shape_t<KID>(t);
    }
}

template <int KID>
__device__ __forceinline__ double shape_s_scaled(double s)
{
    return shape_t_scaled<KID>(fma(-s, rsqrt_fast(s), 1.0));
}

// ++counter if flag, as ONE predicated add (the C++ form compiles to SEL + IADD3; the tile-gather kernel is close to
// issue bound — FP64 is only 52 % of its instructions — so per-pixel integer instructions count)
__device__ __forceinline__ void count_if(bool flag, unsigned& counter)
{
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, 0;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(counter) : "r"((int)flag));
}

// s < 1.0 for s > 0 (or NaN -> false), decided on the integer pipe from the high word: 1.0 = 0x3ff00000'00000000
__device__ __forceinline__ bool below_one(double s) { return __double2hiint(s) < 0x3ff00000; }

// v if flag else +0.0, as a data select (keeps the four per-group dependency chains in one straight-line block;
// a C++ ?: around the kernel evaluation invites the compiler to branch around it per lane)
__device__ __forceinline__ double select_or_zero(bool flag, double v)
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tselp.f64 %0, %1, 0d0000000000000000, p;\n\t}"
        : "=d"(r)
        : "d"(v), "r"((int)flag));
    return r;
}

// ---- FP32-accumulate mode (s2g_set_accumulate_mode): the shape functions in single precision, as functions of
// u = sqrt(s) and t = 1 - u.  The Wendland polynomial factors are kept in their u form here (all coefficients
// positive): the t-expanded forms of the FP64 path cancel ~300:1 near t = 1, which costs nothing at 1e-16 but would
// cost 2e-5 at FP32's 6e-8.
template <int KID>
__device__ __forceinline__ float shape_uf(float u, float t)
{
    if (KID == S2G_KERNEL_CUBIC) {
        const float a = fmaf(6.0f * (u - 1.0f), u * u, 1.0f), b = 2.0f * (t * t * t);
        return u < 0.5f ? a : b;
    } else if (KID == S2G_KERNEL_QUINTIC) {
        const float b = fmaxf(t - 1.0f / 3.0f, 0.0f), c = fmaxf(t - 2.0f / 3.0f, 0.0f);
        const float a2 = t * t, b2 = b * b, c2 = c * c;
        return fmaf(15.0f * c, c2 * c2, fmaf(-6.0f * b, b2 * b2, a2 * a2 * t));
    } else if (KID == S2G_KERNEL_WENDLAND_C2) {
        const float t2 = t * t;
        return (t2 * t2) * fmaf(4.0f, u, 1.0f);
    } else if (KID == S2G_KERNEL_WENDLAND_C4) {
        const float t2 = t * t;
        return (t2 * t2 * t2) * fmaf(fmaf(35.0f / 3.0f, u, 6.0f), u, 1.0f);
    } else if (KID == S2G_KERNEL_WENDLAND_C6) {
        const float t2 = t * t, t4 = t2 * t2;
        return (t4 * t4) * fmaf(fmaf(fmaf(32.0f, u, 25.0f), u, 8.0f), u, 1.0f);
    } else {
        const float t2 = t * t, t4 = t2 * t2;
        return (t4 * t4 * t2) * fmaf(fmaf(fmaf(fmaf(429.0f, u, 450.0f), u, 210.0f), u, 50.0f), u, 5.0f);
    }
}

// w(sqrt(s)) for 1e-30 <= s < 1 (garbage for s >= 1: mask it); sqrt from the MUFU.RSQ seed + one Newton step
template <int KID>
__device__ __forceinline__ float shape_sf(float s)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));
    const float sy = s * y;                        // ~sqrt(s), 2 ulp
    const float e = fmaf(-sy, y, 1.0f);            // 1 - s*y^2
    const float u = fmaf(0.5f * sy, e, sy);        // one Newton step on sqrt(s): removes the seed's bias
    return shape_uf<KID>(u, 1.0f - u);
}

// ---- the same in packed single precision (Blackwell FFMA2 / FMUL2 / FADD2: one instruction, two pixels).  The
// tile-gather kernel is close to issue bound, so in FP32-accumulate mode the instruction count per pixel, not the FP32
// rate, sets the speed: two rows of a thread's column share every arithmetic instruction of the chain.
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }

template <int KID>
__device__ __forceinline__ float2 shape_uf2(float2 u, float2 t)
{
    if (KID == S2G_KERNEL_CUBIC) {
        // u < 0.5: 1 + 6(u-1)u^2 = 1 - 6 u^2 t ; else 2 t^3
        const float2 a = __ffma2_rn(__fmul2_rn(f2(-6.0f), t), __fmul2_rn(u, u), f2(1.0f));
        const float2 b = __fmul2_rn(__fmul2_rn(f2(2.0f), t), __fmul2_rn(t, t));
        return make_float2(u.x < 0.5f ? a.x : b.x, u.y < 0.5f ? a.y : b.y);
    } else if (KID == S2G_KERNEL_QUINTIC) {
        const float2 b = make_float2(fmaxf(t.x - 1.0f / 3.0f, 0.0f), fmaxf(t.y - 1.0f / 3.0f, 0.0f));
        const float2 c = make_float2(fmaxf(t.x - 2.0f / 3.0f, 0.0f), fmaxf(t.y - 2.0f / 3.0f, 0.0f));
        const float2 a2 = __fmul2_rn(t, t), b2 = __fmul2_rn(b, b), c2 = __fmul2_rn(c, c);
        return __ffma2_rn(__fmul2_rn(f2(15.0f), c), __fmul2_rn(c2, c2),
                          __ffma2_rn(__fmul2_rn(f2(-6.0f), b), __fmul2_rn(b2, b2), __fmul2_rn(__fmul2_rn(a2, a2), t)));
    } else if (KID == S2G_KERNEL_WENDLAND_C2) {
        const float2 t2 = __fmul2_rn(t, t);
        return __fmul2_rn(__fmul2_rn(t2, t2), __ffma2_rn(f2(4.0f), u, f2(1.0f)));
    } else if (KID == S2G_KERNEL_WENDLAND_C4) {
        const float2 t2 = __fmul2_rn(t, t);
        return __fmul2_rn(__fmul2_rn(__fmul2_rn(t2, t2), t2),
                          __ffma2_rn(__ffma2_rn(f2(35.0f / 3.0f), u, f2(6.0f)), u, f2(1.0f)));
    } else if (KID == S2G_KERNEL_WENDLAND_C6) {
        const float2 t2 = __fmul2_rn(t, t), t4 = __fmul2_rn(t2, t2);
        return __fmul2_rn(__fmul2_rn(t4, t4),
                          __ffma2_rn(__ffma2_rn(__ffma2_rn(f2(32.0f), u, f2(25.0f)), u, f2(8.0f)), u, f2(1.0f)));
    } else {
        const float2 t2 = __fmul2_rn(t, t), t4 = __fmul2_rn(t2, t2);
        return __fmul2_rn(
            __fmul2_rn(__fmul2_rn(t4, t4), t2),
            __ffma2_rn(__ffma2_rn(__ffma2_rn(__ffma2_rn(f2(429.0f), u, f2(450.0f)), u, f2(210.0f)), u, f2(50.0f)), u,
                       f2(5.0f)));
    }
}

// w(sqrt(s)) for two pixels, 1e-30 <= s < 1 (garbage for s >= 1: mask it); the two MUFU.RSQ seeds are the only scalar
// instructions; signs are arranged so that no packed negation is needed (em = s*y^2 - 1 = -e)
template <int KID>
__device__ __forceinline__ float2 shape_sf2(float2 s)
{
    float2 y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.x) : "f"(s.x));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y.y) : "f"(s.y));
    const float2 sy = __fmul2_rn(s, y);                                   // ~sqrt(s)
    const float2 em = __ffma2_rn(sy, y, f2(-1.0f));                       // s*y^2 - 1
    const float2 u = __ffma2_rn(__fmul2_rn(sy, f2(-0.5f)), em, sy);       // one Newton step on sqrt(s)
    return shape_uf2<KID>(u, __ffma2_rn(u, f2(-1.0f), f2(1.0f)));
}

__device__ __forceinline__ float select_or_zero_f(bool flag, float v)
{
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tselp.f32 %0, %1, 0f00000000, p;\n\t}"
        : "=f"(r)
        : "f"(v), "r"((int)flag));
    return r;
}

__device__ __forceinline__ bool nonzero_bits(double v)
{
    return ((__double2hiint(v) & 0x7fffffff) | __double2loint(v)) != 0;
}

#endif
