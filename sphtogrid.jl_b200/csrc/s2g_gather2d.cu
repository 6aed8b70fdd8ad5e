// s2g_gather2d.cu — 2D Smac deposit, gather strategy, and the per-particle strategy dispatch.
//
// Same arithmetic as the scatter path (cic_2D.jl:11-72, :103-244) but organised so that NO global atomics are
// needed for the bulk of the work:
//   1. k_classify   : per particle -> class (skip / scatter / gather) and number of 64x64 image tiles it touches
//   2. k_norm2d     : pass A (weight sums over the whole footprint), one warp per particle -> GRec with area_norm
//   3. k_expand     : (tile, particle) pairs ; cub radix sort by tile ; tile ranges
//   4. k_gather2d   : one CTA per (tile, chunk of its particle list); every thread OWNS 16 pixels of the tile and
//                     accumulates weight and quantity in registers while the particle records stream through
//                     shared memory; a single coalesced red.add flush per work item at the end.
// Small footprints (few pixels) go to the scatter kernel instead: the per-pair cost of the gather kernel would
// dominate there (threshold: S2G_GATHER_MIN_PIXELS, default below).
#include <cub/cub.cuh>
#include <cstdlib>

#include "s2g_cic2d.cuh"

int s2g_launch_scatter_2d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, const int* list,
                          long long n_list, double* image);
int s2g_launch_scatter_2d_tiny(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, const int* list,
                               long long n_list, double* image);

namespace {

// CTA geometry of k_gather2d: a warp covers 16 columns x 16 rows (2 interleaved rows per pass, RPT rows per thread);
// a CTA is NCG warps side by side (columns) x NRB warps on top of each other (rows).  Compile-time knobs so that
// other shapes can be measured (-DS2G_G2D_NCG=2 -DS2G_G2D_NRB=5 -DS2G_G2D_CTAS=2: 320 threads, 96 registers).
#ifndef S2G_G2D_NCG
#define S2G_G2D_NCG 4
#endif
#ifndef S2G_G2D_NRB
#define S2G_G2D_NRB 2
#endif
#ifndef S2G_G2D_CTAS
#define S2G_G2D_CTAS 3
#endif
constexpr int NCG = S2G_G2D_NCG, NRB = S2G_G2D_NRB, G2D_CTAS = S2G_G2D_CTAS;
constexpr int G2D_THREADS = 32 * NCG * NRB;
constexpr int RPT = 8;                 // rows per thread
constexpr int TILE_W = 16 * NCG;       // tile width  (j, the contiguous image axis)
constexpr int TILE_H = 2 * RPT * NRB;  // tile height (i)
constexpr int BATCH = 256;     // particle records staged in shared memory at a time
constexpr int CHUNK = 4096;    // max pairs per work item

struct __align__(16) GRec {
    int iMin, iMax, jMin, jMax;  // footprint; iMin > iMax: record unused (particle re-routed to the scatter kernel)
    double x, y;                 // pixel coordinates of the particle
    double hinv, an;             // 1/h [1/pixel], area_norm (cic_2D.jl:188)
    double dx_lo, dx_hi;         // overlap lengths of the first / last row    (interior rows: 1)
    double dy_lo, dy_hi;         //                     first / last column
    double h;
    int p;                       // particle index (for the per-image quantity)
    int f32_ok;                  // FP32-accumulate mode may evaluate this particle in single precision (k_norm2d)
};

// integral of the kernel shape over the unit disc, ∫ w(u) 2πu du = 1 / norm_2D(kernel)
__host__ __device__ constexpr double shape_integral_2d(int kid)
{
    constexpr double pi = 3.14159265358979323846;
    return kid == S2G_KERNEL_CUBIC         ? 7.0 * pi / 40.0
           : kid == S2G_KERNEL_QUINTIC     ? 478.0 * pi / 15309.0
           : kid == S2G_KERNEL_WENDLAND_C2 ? pi / 7.0
           : kid == S2G_KERNEL_WENDLAND_C4 ? pi / 9.0
           : kid == S2G_KERNEL_WENDLAND_C6 ? 7.0 * pi / 78.0
                                           : 3.0 * pi / 8.0;
}

// smallest h [pixels] from which the discrete pass-A sum of an UNCLIPPED footprint equals h^2 * shape_integral to
// better than 5e-12 relative — a 20th of the 1e-10 parity budget (tools/analytic_norm_study.py; Poisson summation:
// the difference is the kernel's Fourier transform at the pixel frequency).  Infinity: never use the integral.
__host__ __device__ constexpr double analytic_norm_min_h(int kid)
{
    return kid == S2G_KERNEL_WENDLAND_C8   ? 20.0
           : kid == S2G_KERNEL_WENDLAND_C6 ? 32.0
           : kid == S2G_KERNEL_WENDLAND_C4 ? 56.0
           : kid == S2G_KERNEL_QUINTIC     ? 64.0
                                           : 1e300;
}

// does the kernel support (circle of radius h around (x,y)) possibly reach a pixel centre of tile (ti,tj)?
// conservative (never rejects a tile that holds a pixel with u < 1); evaluated on the GRec fields by BOTH the pair
// count (k_norm2d) and the pair expansion (k_expand), so the two always agree.
__device__ __forceinline__ bool tile_hit(const GRec& g, int ti, int tj)
{
    const double lo_i = ti * (double)TILE_H + 0.5, hi_i = lo_i + (TILE_H - 1);
    const double lo_j = tj * (double)TILE_W + 0.5, hi_j = lo_j + (TILE_W - 1);
    const double ddx = fmax(fmax(lo_i - g.x, 0.0), g.x - hi_i);
    const double ddy = fmax(fmax(lo_j - g.y, 0.0), g.y - hi_j);
    const double hh = g.h * (1.0 + 1e-9) + 1e-9;
    return ddx * ddx + ddy * ddy <= hh * hh;
}

// class: 0 = nothing to do, 1 = scatter (warp per particle), 2 = gather, 3 = tiny scatter (8 lanes per particle)
// — the footprint-size bins of the deposit (F_p = product of the pix_index_min_max ranges, cic_shared.jl:46-52)
__global__ void __launch_bounds__(256) k_classify(s2g_particles P, s2g_geom G, long long p0, long long nb,
                                                  long long gather_min_pixels, long long tiny_max_pixels, int force,
                                                  int* __restrict__ cls, unsigned* __restrict__ npairs)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nb) return;
    Rec2 r;
    int c = 0;
    unsigned np = 0;
    if (make_rec2(P, G, p0 + t, r)) {
        const long long fp = (long long)(r.iMax - r.iMin + 1) * (r.jMax - r.jMin + 1);
        const bool gather = force == S2G_STRATEGY_GATHER || (force == S2G_STRATEGY_AUTO && fp >= gather_min_pixels);
        c = gather ? 2 : (fp <= tiny_max_pixels ? 3 : 1);
        if (gather)  // upper bound of the number of (tile, particle) pairs
            np = (unsigned)((r.iMax / TILE_H - r.iMin / TILE_H + 1) * (r.jMax / TILE_W - r.jMin / TILE_W + 1));
    }
    cls[t] = c;
    npairs[t] = np;
}

// compact lists of scatter and gather particles (indices relative to the whole particle set)
__global__ void __launch_bounds__(256) k_build_lists(const int* __restrict__ cls, const unsigned* __restrict__ pos_s,
                                                     const unsigned* __restrict__ pos_g,
                                                     const unsigned* __restrict__ pos_t, long long p0, long long nb,
                                                     int* __restrict__ list_s, int* __restrict__ list_g,
                                                     int* __restrict__ list_t)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nb) return;
    const int c = cls[t];
    if (c == 1) list_s[pos_s[t]] = (int)(p0 + t);
    if (c == 2) list_g[pos_g[t]] = (int)(p0 + t);
    if (c == 3) list_t[pos_t[t]] = (int)(p0 + t);
}

// Σ w(u)·dA over the pixel centres inside the kernel, for the pixel rectangle [ia,ib] x [ja,jb] (indices may lie
// outside the image: "virtual" pixels), optionally EXCLUDING the image [0,npix)².  Lanes own columns, rows are walked
// four at a time.  dA = dx·dy with dx, dy = 1 except in the first/last row/column of the UNCLIPPED footprint
// (iEl, iEh, jEl, jEh), where the overlap lengths dxl, dxh, dyl, dyh apply (get_dxyz, cic_shared.jl:60-62).
// Returns this lane's partial sum; c counts the pixels inside.
template <int KID>
__device__ __forceinline__ double pass_a_region(double x, double y, double h, double hinv, int lane, int ia, int ib,
                                                int ja, int jb, int iEl, int iEh, double dxl, double dxh, int jEl,
                                                int jEh, double dyl, double dyh, bool excl, int npix, int& c)
{
    double acc = 0.0;
    const int nj = jb - ja + 1;
    const double xb = center_dist(x, (double)ia) * hinv;  // a(i) = xb - (i - ia)*hinv  (origin at ia: no cancellation)
    for (int jc = lane; jc < ((nj + 31) & ~31); jc += 32) {
        const int j = ja + jc;
        const bool col = jc < nj;
        const double b = center_dist(y, (double)j) * hinv;
        const double b2 = col ? b * b : 4.0;
        const double dy = (j == jEl) ? dyl : ((j == jEh) ? dyh : 1.0);
        // rows that can reach this 32-column group: a² < 1 - min_lanes(b²)
        double m = b2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (m >= 1.0) continue;
        const double reach = sqrt(1.0 - m) * h + 1.0;
        const int r_lo = max(ia, (int)floor(x - 0.5 - reach));
        const int r_hi = min(ib, (int)ceil(x - 0.5 + reach));
        const bool jin = excl && (j >= 0) && (j < npix);
        // all columns of the group inside the image: only the rows above and below the image remain
        const bool all_in = excl && __all_sync(0xffffffffu, jin || !col);
        const double b2s = fmax(b2, 1e-300);  // keeps s > 0 when a pixel centre sits on the particle
        double colsum = 0.0;
        for (int part = 0; part < 2; ++part) {
            int lo = r_lo, hi = r_hi;
            if (all_in) {
                if (part == 0) hi = min(hi, -1); else lo = max(lo, npix);
            } else if (part == 1)
                break;
            for (int i0 = lo; i0 <= hi; i0 += 4) {
                double wk[4];
                bool in[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = i0 + k;
                    const double a = fma(-(double)(i - ia), hinv, xb);
                    const double s = fma(a, a, b2s);
                    in[k] = below_one(s) && (i <= hi) && !(jin && i >= 0 && i < npix);
                    wk[k] = shape_s<KID>(s);
                }
                if (i0 <= iEl || i0 + 3 >= iEh) {  // group touches the first / last row: partial overlaps
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = i0 + k;
                        const double dx = (i == iEl) ? dxl : ((i == iEh) ? dxh : 1.0);
                        colsum = fma(select_or_zero(in[k], wk[k]), dx, colsum);
                        c += in[k] ? 1 : 0;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        colsum += select_or_zero(in[k], wk[k]);
                        c += in[k] ? 1 : 0;
                    }
                }
            }
        }
        acc = fma(colsum, dy, acc);
    }
    return acc;
}

// ---- pass A for gather particles (cic_2D.jl:11-72): one warp per particle.
// distr_weight = Σ w(u)·dA over the pixel centres inside the kernel.  Work in units of h: with a = (x-i-0.5)/h,
// b = (y-j-0.5)/h a pixel is inside iff a²+b² < 1.  Lanes own columns, rows are walked four at a time (independent
// dependency chains keep the FP64 pipe busy).  Unclipped, well-resolved footprints take the closed form
// h²·∫w instead (analytic_norm_min_h).  Particles in the "no pixel centre covered" branch are re-routed to the
// scatter kernel (they are tiny or clipped; their wk := 1 deposit does not fit the gather kernel's inner loop).
// `slow` (optional): the entries of `list` the warp kernel still has to do, *n_slow of them (built by k_norm2d_fast).
template <int KID>
__global__ void __launch_bounds__(256) k_norm2d(s2g_particles P, s2g_geom G, const int* __restrict__ list,
                                                long long n_list, int exact_norm, GRec* __restrict__ recs,
                                                unsigned* __restrict__ npairs_g, int* __restrict__ reroute,
                                                unsigned long long* __restrict__ counters,
                                                const unsigned* __restrict__ slow)
{
    const int lane = threadIdx.x & 31;
    unsigned long long mapped = 0, fpx = 0;
    const long long n_work = slow ? (long long)counters[CNT_AUX] : n_list;
    for (;;) {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(&counters[CNT_WORK], 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_work) break;
        if (slow) t = (long long)slow[t];
        const long long p = list[t];
        Rec2 r;
        make_rec2(P, G, p, r);  // known valid
        const int ni = r.iMax - r.iMin + 1, nj = r.jMax - r.jMin + 1;
        const double dx_lo = overlap_1d(r.x, r.h, r.iMin), dx_hi = overlap_1d(r.x, r.h, r.iMax);
        const double dy_lo = overlap_1d(r.y, r.h, r.jMin), dy_hi = overlap_1d(r.y, r.h, r.jMax);
        const int n1 = (int)G.npix - 1;
        // clipped by the image border?  (unclipped: the bounds are the plain floors of x±h)
        const bool unclipped = floor_to_int(r.x - r.h) >= 0 && floor_to_int(r.x + r.h) <= n1 &&
                               floor_to_int(r.y - r.h) >= 0 && floor_to_int(r.y + r.h) <= n1;
        // first/last row and column of the UNCLIPPED footprint and their overlap lengths
        const int iEl = floor_to_int(r.x - r.h), iEh = floor_to_int(r.x + r.h);
        const int jEl = floor_to_int(r.y - r.h), jEh = floor_to_int(r.y + r.h);
        const double dxl = overlap_1d(r.x, r.h, iEl), dxh = overlap_1d(r.x, r.h, iEh);
        const double dyl = overlap_1d(r.y, r.h, jEl), dyh = overlap_1d(r.y, r.h, jEh);
        const bool resolved = !exact_norm && r.h >= analytic_norm_min_h(KID);
        const double closed = r.h * r.h * shape_integral_2d(KID);
        double sw = -1.0;
        long long cnt = 1;
        if (resolved && unclipped) {
            sw = closed;
        } else if (resolved) {
            // clipped by the image border: the sum over ALL lattice pixels of the unclipped footprint equals the closed
            // form; subtract the (smaller) numerical sum over the virtual pixels that lie OUTSIDE the image
            int c = 0;
            const double out = warp_sum(pass_a_region<KID>(r.x, r.y, r.h, r.hinv, lane, iEl, iEh, jEl, jEh, iEl, iEh,
                                                           dxl, dxh, jEl, jEh, dyl, dyh, true, (int)G.npix, c));
            // keep the subtraction well conditioned: if little of the kernel is left inside, sum directly instead
            if (out < 0.75 * closed) sw = closed - out;
        }
        if (sw < 0.0) {
            int c = 0;
            sw = warp_sum(pass_a_region<KID>(r.x, r.y, r.h, r.hinv, lane, r.iMin, r.iMax, r.jMin, r.jMax, iEl, iEh, dxl,
                                             dxh, jEl, jEh, dyl, dyh, false, (int)G.npix, c));
            cnt = __reduce_add_sync(0xffffffffu, c);
        }
        GRec g;
        g.x = r.x; g.y = r.y; g.h = r.h; g.hinv = r.hinv;
        g.dx_lo = dx_lo; g.dx_hi = dx_hi; g.dy_lo = dy_lo; g.dy_hi = dy_hi;
        g.iMin = r.iMin; g.iMax = r.iMax; g.jMin = r.jMin; g.jMax = r.jMax;
        g.p = (int)p;
        // FP32-accumulate mode: a footprint that keeps less than half of the kernel integral inside the image is
        // renormalised by the reference to carry the particle's whole weight anyway, so contributions from the kernel
        // rim (1-u)^k, whose relative error in single precision is ~k*6e-8/(1-u), can dominate a border pixel (measured:
        // 2.8e-5 on a pixel at u = 0.983).  Those particles keep the FP64 chain.
        g.f32_ok = (sw >= 0.5 * closed) ? 1 : 0;
        unsigned np = 0;
        const double an_probe = (r.area / sw) * r.w * r.dz;
        if (sw == 0.0 || !isfinite(an_probe)) {
            // cic_2D.jl:51-66 branch, or an Inf/NaN normalisation (rho = 0, NaN weight) -> scatter kernel
            g.iMin = 1; g.iMax = 0; g.an = 0.0;
            if (lane == 0) reroute[atomicAdd(&counters[CNT_PAIRS], 1ull)] = (int)p;
        } else {
            // kernel_norm * weight_per_pix = (area/N) * (N/sw); N cancels to rounding (cic_2D.jl:187-188)
            const double n_distr = (double)cnt;
            const double kernel_norm = r.area / n_distr;
            g.an = kernel_norm * (n_distr / sw) * r.w * r.dz;
            const int ti0 = r.iMin / TILE_H, tj0 = r.jMin / TILE_W;
            const int nti = r.iMax / TILE_H - ti0 + 1, ntj = r.jMax / TILE_W - tj0 + 1;
            for (int q = lane; q < nti * ntj; q += 32)
                if (tile_hit(g, ti0 + q / ntj, tj0 + q % ntj)) ++np;
            np = __reduce_add_sync(0xffffffffu, np);
            if (lane == 0) {
                ++mapped;
                fpx += (unsigned long long)ni * (unsigned long long)nj;
            }
        }
        if (lane == 0) {
            recs[t] = g;
            npairs_g[t] = np;
        }
    }
    if (lane == 0) {
        if (mapped) { atomicAdd(&counters[CNT_MAPPED], mapped); atomicAdd(&counters[CNT_GATHER], mapped); }
        if (fpx) atomicAdd(&counters[CNT_FOOTPRINT], fpx);
    }
}

// ---- pass A, one THREAD per particle, for what needs no sum: an unclipped footprint resolved well enough for the closed
// form (the bulk of a gather class).  The record, the normalisation and the tile count are the per-particle arithmetic
// of k_norm2d's closed-form branch, bit for bit; a warp per particle left 31 lanes idle there and paid one memory
// latency per particle (C5: 2.7 s of a 25.8-s map).  Everything else — clipped, under-resolved, degenerate — is
// appended to `slow` for the warp kernel.
template <int KID>
__global__ void __launch_bounds__(256) k_norm2d_fast(s2g_particles P, s2g_geom G, const int* __restrict__ list,
                                                     long long n_list, GRec* __restrict__ recs,
                                                     unsigned* __restrict__ npairs_g, unsigned* __restrict__ slow,
                                                     unsigned long long* __restrict__ counters)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool is_slow = false;
    unsigned long long mapped = 0, fpx = 0;
    if (t < n_list) {
        const long long p = list[t];
        Rec2 r;
        make_rec2(P, G, p, r);  // known valid
        const int n1 = (int)G.npix - 1;
        const bool unclipped = floor_to_int(r.x - r.h) >= 0 && floor_to_int(r.x + r.h) <= n1 &&
                               floor_to_int(r.y - r.h) >= 0 && floor_to_int(r.y + r.h) <= n1;
        const bool resolved = r.h >= analytic_norm_min_h(KID);
        const double sw = r.h * r.h * shape_integral_2d(KID);
        const double an_probe = (r.area / sw) * r.w * r.dz;
        if (!(resolved && unclipped) || sw == 0.0 || !isfinite(an_probe)) {
            is_slow = true;
        } else {
            GRec g;
            g.x = r.x; g.y = r.y; g.h = r.h; g.hinv = r.hinv;
            g.dx_lo = overlap_1d(r.x, r.h, r.iMin); g.dx_hi = overlap_1d(r.x, r.h, r.iMax);
            g.dy_lo = overlap_1d(r.y, r.h, r.jMin); g.dy_hi = overlap_1d(r.y, r.h, r.jMax);
            g.iMin = r.iMin; g.iMax = r.iMax; g.jMin = r.jMin; g.jMax = r.jMax;
            g.p = (int)p;
            g.f32_ok = 1;
            const double n_distr = 1.0;
            const double kernel_norm = r.area / n_distr;
            g.an = kernel_norm * (n_distr / sw) * r.w * r.dz;
            const int ti0 = r.iMin / TILE_H, ti1 = r.iMax / TILE_H, tj0 = r.jMin / TILE_W, tj1 = r.jMax / TILE_W;
            unsigned np = 0;
            for (int ti = ti0; ti <= ti1; ++ti)
                for (int tj = tj0; tj <= tj1; ++tj)
                    if (tile_hit(g, ti, tj)) ++np;
            recs[t] = g;
            npairs_g[t] = np;
            mapped = 1;
            fpx = (unsigned long long)(r.iMax - r.iMin + 1) * (unsigned long long)(r.jMax - r.jMin + 1);
        }
    }
    const unsigned sm = __ballot_sync(0xffffffffu, is_slow);
    if (sm) {
        unsigned long long base = 0;
        if (lane == __ffs(sm) - 1) base = atomicAdd(&counters[CNT_AUX], (unsigned long long)__popc(sm));
        base = __shfl_sync(0xffffffffu, base, __ffs(sm) - 1);
        if (is_slow) slow[base + __popc(sm & ((1u << lane) - 1u))] = (unsigned)t;
    }
    mapped = (unsigned long long)warp_sum_ll((long long)mapped);
    fpx = (unsigned long long)warp_sum_ll((long long)fpx);
    if (lane == 0 && mapped) {
        atomicAdd(&counters[CNT_MAPPED], mapped);
        atomicAdd(&counters[CNT_GATHER], mapped);
        atomicAdd(&counters[CNT_FOOTPRINT], fpx);
    }
}

// ---- (tile, record) pairs
__global__ void __launch_bounds__(256) k_expand(const GRec* __restrict__ recs, const unsigned* __restrict__ off,
                                                long long n_list, int ntile_j, unsigned* __restrict__ keys,
                                                unsigned* __restrict__ vals)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_list) return;
    const GRec g = recs[t];
    if (g.iMin > g.iMax) return;
    const int ti0 = g.iMin / TILE_H, ti1 = g.iMax / TILE_H, tj0 = g.jMin / TILE_W, tj1 = g.jMax / TILE_W;
    unsigned o = off[t];
    for (int ti = ti0; ti <= ti1; ++ti)
        for (int tj = tj0; tj <= tj1; ++tj)
            if (tile_hit(g, ti, tj)) {
                keys[o] = (unsigned)(ti * ntile_j + tj);
                vals[o] = (unsigned)t;
                ++o;
            }
}

// tile ranges of the sorted pair list by boundary detection (no atomics): [tile_beg[t], tile_end[t])
__global__ void __launch_bounds__(256) k_tile_bounds(const unsigned* __restrict__ keys, long long m,
                                                     unsigned* __restrict__ tile_beg, unsigned* __restrict__ tile_end)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= m) return;
    const unsigned k = keys[t];
    if (t == 0 || keys[t - 1] != k) tile_beg[k] = (unsigned)t;
    if (t == m - 1 || keys[t + 1] != k) tile_end[k] = (unsigned)(t + 1);
}

__global__ void __launch_bounds__(256) k_tile_chunks(const unsigned* __restrict__ tile_beg,
                                                     const unsigned* __restrict__ tile_end, int ntiles,
                                                     unsigned* __restrict__ nchunks)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    nchunks[t] = (tile_end[t] - tile_beg[t] + CHUNK - 1) / CHUNK;
}

// ---- the gather kernel: pass B (cic_2D.jl:193-222) without atomics.
// A CTA owns a TILE_H x TILE_W tile for the duration of a work item (a chunk of the tile's particle list).  Every
// thread owns RPT pixels of the tile and keeps their weight and quantity sums in registers; particle records stream
// through shared memory.  One coalesced red.add flush per work item.
// Lane layout: a warp covers 16 columns x 2 interleaved rows (half-warp rows halve the idle lanes at the two ends
// of each chord of the kernel disc compared with 32-wide rows); warp w: column group w&3, row block w>>2;
// thread rows: i = i0 + (w>>2)*2*RPT + 2*r + (lane>>4), r = 0..RPT-1.
// F32 = FP32-accumulate mode (s2g_set_accumulate_mode): the per-(record, thread) preamble stays FP64 (tile-relative
// coordinates in units of h, so the conversion to float costs 6e-8 of a quantity of order one), the per-pixel chain
// (s, 1 - sqrt(s), polynomial, two FMAs) runs in FP32, and the FP32 partial sums are folded into the FP64 register
// accumulators after every batch of 256 records: per-pixel relative error ~1e-7 median, 1e-6 worst (bar of the mode:
// 1e-5).  Records flagged !f32_ok (most of the kernel clipped away by the image border) take the FP64 chain.
template <int KID, bool F32>
__global__ void __launch_bounds__(G2D_THREADS, G2D_CTAS) k_gather2d(const GRec* __restrict__ recs, const unsigned* __restrict__ vals,
                                                     const unsigned* __restrict__ tile_beg,
                                                     const unsigned* __restrict__ tile_end,
                                                     const unsigned* __restrict__ chunk_begin,  // ntiles+1
                                                     int ntiles, int ntile_j, unsigned total_chunks,
                                                     const void* __restrict__ binq, int in_dtype, int n_images,
                                                     int image_k, long long npix, double* __restrict__ image,
                                                     unsigned long long* __restrict__ counters)
{
    __shared__ GRec s_rec[BATCH];
    __shared__ double s_q[BATCH];
    __shared__ float4 s_f[F32 ? BATCH : 1];  // FP32-accumulate mode: (2/h, dx_lo, dx_hi, -) converted once per record
    __shared__ unsigned s_work[3];

    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    const int jl = (wq % NCG) * 16 + (lane & 15);  // column inside the tile
    const int rpar = lane >> 4;                   // row parity inside the warp
    const int rblk = (wq / NCG) * (2 * RPT);      // first row of the warp's row block
    unsigned touched = 0;

    for (;;) {
        if (tid == 0) {
            const unsigned w = (unsigned)atomicAdd(&counters[CNT_WORK], 1ull);
            unsigned tile = 0xffffffffu, b = 0, e = 0;
            if (w < total_chunks) {
                int lo = 0, hi = ntiles;  // last tile with chunk_begin[tile] <= w
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (chunk_begin[mid] <= w) lo = mid; else hi = mid;
                }
                tile = (unsigned)lo;
                const unsigned c = w - chunk_begin[lo];
                b = tile_beg[lo] + c * CHUNK;
                e = min(b + CHUNK, tile_end[lo]);
            }
            s_work[0] = tile; s_work[1] = b; s_work[2] = e;
        }
        __syncthreads();
        const unsigned tile = s_work[0], wb = s_work[1], we = s_work[2];
        if (tile == 0xffffffffu) break;
        const int i0 = (int)(tile / ntile_j) * TILE_H, j0 = (int)(tile % ntile_j) * TILE_W;
        const int j = j0 + jl;
        const int jw0 = j0 + (wq % NCG) * 16;  // first column of this warp (uniform)
        const int wbase = i0 + rblk;          // first row of this warp (uniform)
        const int ibase = wbase + rpar;       // first row of this thread; its rows are ibase + 2r
        // this thread's column / first row centre inside the tile.  Opaque to the compiler on purpose: as plain functions
        // of threadIdx ptxas re-materialises them for every record (LOP3 + I2F.F64 + DADD, twice) instead of keeping
        // four registers
        double jld = (double)jl + 0.5, ild = (double)(rblk + rpar) + 0.5;
        asm volatile("" : "+d"(jld), "+d"(ild));

        // FP64 chain: in the FP64-only kernel the polynomial factor is scaled by 1/shape_scale (folded into g.an above)
        auto shape_fp64 = [](double s) { return F32 ? shape_s<KID>(s) : shape_s_scaled<KID>(s); };
        // FP64 mode: the thread's 8 pixels x 2 planes live in FP64 registers for the whole work item.
        // FP32-accumulate mode: packed FP32 partial sums of the current batch of 256 records (row pairs), added to the
        // FP64 image at the end of every batch — no FP64 accumulator registers, which is what lets this variant keep
        // 3 CTAs/SM; the accumulation error does not grow with the length of a tile's particle list.
        double acc_w[F32 ? 1 : RPT], acc_q[F32 ? 1 : RPT];
        float2 facc_w[F32 ? RPT / 2 : 1], facc_q[F32 ? RPT / 2 : 1];
#pragma unroll
        for (int r = 0; r < (F32 ? 1 : RPT); ++r) { acc_w[r] = 0.0; acc_q[r] = 0.0; }
#pragma unroll
        for (int r = 0; r < (F32 ? RPT / 2 : 1); ++r) { facc_w[r] = f2(0.0f); facc_q[r] = f2(0.0f); }
        const long long npl = npix * npix;
        // one pixel of this thread's column to the image: half-warps write 16 consecutive doubles (128 B) of a row
        auto flush_px = [&](int i, double w, double q) {
            // FP32-accumulate mode flushes inside the batch loop: keep the address arithmetic HERE (hoisted out of the
            // loop as loop-invariant it pins 26 registers for the 16 addresses and spills the rest)
            if constexpr (F32) asm volatile("" : "+r"(i));
            if (i < npix && (w != 0.0 || q != 0.0)) {
                const long long idx = (long long)i * npix + j;
                if (image_k == 0) red_add(image + npl * n_images + idx, w);
                red_add(image + npl * image_k + idx, q);
            }
        };

        for (unsigned b = wb; b < we; b += BATCH) {
            const int nb = (int)min((unsigned)BATCH, we - b);
            __syncthreads();  // previous batch fully consumed
            for (int t = tid; t < nb; t += G2D_THREADS) {
                // the shared-memory copy is made TILE-RELATIVE by the loading thread: x := x - i0, y := y - j0 (exact: a
                // multiple of ulp(x) that is smaller than x, or a difference of at most h).  The per-(record, thread)
                // offset is then ONE subtraction of the thread's constant (row + 0.5), a single rounding of the exact
                // x - i - 0.5 where get_x_dx (cic_shared.jl:73) rounds twice — never further from the exact value.
                // (Scaling by 1/h BEFORE subtracting the thread's offset saves another instruction but cancels: the
                // error grows to 64/h ulp, which the kernel rim (1-u)^k amplifies beyond the 1e-10 bar — measured.)
                //   h := 2/h (row stride of a thread: its rows are two apart);  s_q := area_norm * quantity
                GRec g = recs[vals[b + t]];
                if (!F32) g.an *= shape_scale<KID>();  // exact (power of two), see shape_t_scaled
                g.x = g.x - (double)i0;
                g.y = g.y - (double)j0;
                g.h = g.hinv + g.hinv;
                s_rec[t] = g;
                s_q[t] = g.an * ld_in(binq, (long long)n_images * g.p + image_k, in_dtype);
                if constexpr (F32) s_f[t] = make_float4((float)g.h, (float)g.dx_lo, (float)g.dx_hi, 0.0f);
            }
            __syncthreads();
            for (int e = 0; e < nb; ++e) {
                const GRec& g = s_rec[e];
                // warp-uniform integer culls: the warp's 16 rows x 16 columns against the footprint box
                const int rlo = max(g.iMin, wbase), rhi = min(g.iMax, wbase + 2 * RPT - 1);
                if (rlo > rhi || g.jMax < jw0 || g.jMin > jw0 + 15) continue;
                const double hinv = g.hinv;
                const double bq = (g.y - jld) * hinv;
                const double b2 = fma(bq, bq, 1e-300);  // s > 0 even when a pixel centre sits on the particle
                const double dy = (j == g.jMin) ? g.dy_lo : ((j == g.jMax) ? g.dy_hi : 1.0);
                const double dyan = dy * g.an;
                // pix_weight != 0 test of cic_2D.jl:211 hoisted: wk > 0 inside the disc, so only dy*area_norm decides
                const bool live = (j >= g.jMin) && (j <= g.jMax) && below_one(b2) && nonzero_bits(dyan);
                if (!__any_sync(0xffffffffu, live)) continue;
                const double dyanq = dy * s_q[e];
                const double xb = (g.x - ild) * hinv;  // a of this thread's first row
                const double hinv2 = g.h;
                // pixel r of this thread += wk * (dyan, dyanq), FP64 chain (in the FP32-accumulate kernel: the records
                // whose kernel is mostly clipped away are EVALUATED in FP64; their products join the FP32 partial sums)
                auto add_px = [&](int r, double wk) {
                    if constexpr (F32) {
                        const float pw = (float)(wk * dyan), pq = (float)(wk * dyanq);
                        if (r & 1) { facc_w[r >> 1].y += pw; facc_q[r >> 1].y += pq; }
                        else       { facc_w[r >> 1].x += pw; facc_q[r >> 1].x += pq; }
                    } else {
                        acc_w[r] = fma(wk, dyan, acc_w[r]);
                        acc_q[r] = fma(wk, dyanq, acc_q[r]);
                    }
                };
                if (F32 && g.f32_ok) {
                    // two rows per instruction (FFMA2): pairs (r4, r4+1) and (r4+2, r4+3) of each group of four
                    const float4 gf = s_f[F32 ? e : 0];
                    const float2 b22 = f2(fmaxf((float)b2, 1e-30f)), xb2 = f2((float)xb), hinv22 = f2(gf.x);
                    const float2 dyan2 = f2((float)dyan), dyanq2 = f2((float)dyanq);
                    const float dxlo = gf.y, dxhi = gf.z;
#pragma unroll
                    for (int r4 = 0; r4 < RPT; r4 += 4) {
                        const int g_lo = wbase + 2 * r4, g_hi = g_lo + 7;
                        if (g_lo > rhi || g_hi < rlo) continue;  // uniform
                        float2 sf[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float2 a = __ffma2_rn(make_float2(-(float)(r4 + 2 * h), -(float)(r4 + 2 * h + 1)),
                                                        hinv22, xb2);
                            sf[h] = __ffma2_rn(a, a, b22);
                        }
                        bool in[4];
                        in[0] = live && (sf[0].x < 1.0f); in[1] = live && (sf[0].y < 1.0f);
                        in[2] = live && (sf[1].x < 1.0f); in[3] = live && (sf[1].y < 1.0f);
                        if ((g_lo > g.iMin) && (g_hi < g.iMax)) {  // uniform: no first/last row in this group
                            if (!__any_sync(0xffffffffu, in[0] || in[1] || in[2] || in[3])) continue;
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                float2 wk = shape_sf2<KID>(sf[h]);
                                wk.x = select_or_zero_f(in[2 * h], wk.x);
                                wk.y = select_or_zero_f(in[2 * h + 1], wk.y);
                                facc_w[(r4 >> 1) + h] = __ffma2_rn(wk, dyan2, facc_w[(r4 >> 1) + h]);
                                facc_q[(r4 >> 1) + h] = __ffma2_rn(wk, dyanq2, facc_q[(r4 >> 1) + h]);
                                count_if(in[2 * h], touched);
                                count_if(in[2 * h + 1], touched);
                            }
                        } else {
                            float dx[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int i = ibase + 2 * (r4 + k);
                                in[k] = in[k] && (i >= g.iMin) && (i <= g.iMax);
                                dx[k] = (i == g.iMin) ? dxlo : ((i == g.iMax) ? dxhi : 1.0f);
                            }
                            if (!__any_sync(0xffffffffu, in[0] || in[1] || in[2] || in[3])) continue;
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                float2 wk = __fmul2_rn(shape_sf2<KID>(sf[h]), make_float2(dx[2 * h], dx[2 * h + 1]));
                                wk.x = select_or_zero_f(in[2 * h], wk.x);
                                wk.y = select_or_zero_f(in[2 * h + 1], wk.y);
                                facc_w[(r4 >> 1) + h] = __ffma2_rn(wk, dyan2, facc_w[(r4 >> 1) + h]);
                                facc_q[(r4 >> 1) + h] = __ffma2_rn(wk, dyanq2, facc_q[(r4 >> 1) + h]);
                                count_if(in[2 * h] && dx[2 * h] != 0.0f, touched);
                                count_if(in[2 * h + 1] && dx[2 * h + 1] != 0.0f, touched);
                            }
                        }
                    }
                    continue;
                }
#pragma unroll
                for (int r4 = 0; r4 < RPT; r4 += 4) {
                    const int g_lo = wbase + 2 * r4, g_hi = g_lo + 7;  // rows of this group (both parities)
                    if (g_lo > rhi || g_hi < rlo) continue;            // uniform
                    double s[4];
                    bool in[4];
                    bool any_in = false;
                    const bool interior = (g_lo > g.iMin) && (g_hi < g.iMax);  // uniform: no first/last row inside
                    if (interior) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const double a = fma(-(double)(r4 + k), hinv2, xb);
                            s[k] = fma(a, a, b2);
                            in[k] = live && below_one(s[k]);
                            any_in = any_in || in[k];
                        }
                        if (!__any_sync(0xffffffffu, any_in)) continue;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            add_px(r4 + k, select_or_zero(in[k], shape_fp64(s[k])));
                            count_if(in[k], touched);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int i = ibase + 2 * (r4 + k);
                            const double a = fma(-(double)(r4 + k), hinv2, xb);
                            s[k] = fma(a, a, b2);
                            in[k] = live && below_one(s[k]) && (i >= g.iMin) && (i <= g.iMax);
                            any_in = any_in || in[k];
                        }
                        if (!__any_sync(0xffffffffu, any_in)) continue;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int i = ibase + 2 * (r4 + k);
                            const double dx = (i == g.iMin) ? g.dx_lo : ((i == g.iMax) ? g.dx_hi : 1.0);
                            add_px(r4 + k, select_or_zero(in[k], shape_fp64(s[k]) * dx));
                            count_if(in[k] && nonzero_bits(dx), touched);
                        }
                    }
                }
            }
            if constexpr (F32) {  // the batch's FP32 partial sums join the FP64 image
                if (j < npix) {
#pragma unroll
                    for (int r = 0; r < RPT / 2; ++r) {
                        flush_px(ibase + 4 * r, (double)facc_w[r].x, (double)facc_q[r].x);
                        flush_px(ibase + 4 * r + 2, (double)facc_w[r].y, (double)facc_q[r].y);
                    }
                }
#pragma unroll
                for (int r = 0; r < RPT / 2; ++r) { facc_w[r] = f2(0.0f); facc_q[r] = f2(0.0f); }
            }
        }
        if constexpr (!F32) {  // one flush per work item
            if (j < npix) {
#pragma unroll
                for (int r = 0; r < RPT; ++r) flush_px(ibase + 2 * r, acc_w[r], acc_q[r]);
            }
        }
        if (image_k == 0 && touched > 0x7f000000u) {  // keep the 32-bit per-thread counter from wrapping
            atomicAdd(&counters[CNT_TOUCHED], (unsigned long long)touched);
            touched = 0;
        }
        __syncthreads();  // s_work reuse
    }
    if (image_k == 0) {
        unsigned long long tt = (unsigned long long)warp_sum_ll((long long)touched);
        if (lane == 0 && tt) atomicAdd(&counters[CNT_TOUCHED], tt);
    }
}

struct KLaunch {
    int (*norm)(s2g_ctx*, const s2g_particles&, const s2g_geom&, const int*, long long, int, GRec*, unsigned*, int*);
    int (*gather)(s2g_ctx*, const GRec*, const unsigned*, const unsigned*, const unsigned*, const unsigned*, int, int,
                  unsigned, const s2g_particles&, const s2g_geom&, int, double*);
};

template <int KID>
int launch_norm(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, const int* list, long long n_list,
                int exact_norm, GRec* recs, unsigned* npairs_g, int* reroute)
{
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_PAIRS, 0, sizeof(unsigned long long), ctx->stream));
    const unsigned* slow = nullptr;
    if (!exact_norm && analytic_norm_min_h(KID) < 1e200 && n_list > 0) {
        // thread per particle for the closed-form particles; the warp kernel does the listed rest
        void* d_slow;
        S2G_TRY(s2g_scratch(ctx, "g_slow", sizeof(unsigned) * (size_t)n_list, &d_slow));
        S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_AUX, 0, sizeof(unsigned long long), ctx->stream));
        k_norm2d_fast<KID><<<(int)((n_list + 255) / 256), 256, 0, ctx->stream>>>(P, G, list, n_list, recs, npairs_g,
                                                                                  (unsigned*)d_slow, ctx->d_counters);
        S2G_CUDA(cudaGetLastError());
        ctx->launches += 1;
        slow = (const unsigned*)d_slow;
    }
    const int blocks = (int)std::min<long long>((n_list + 7) / 8, (long long)ctx->sm_count * 8);
    k_norm2d<KID><<<max(blocks, 1), 256, 0, ctx->stream>>>(P, G, list, n_list, exact_norm, recs, npairs_g, reroute,
                                                          ctx->d_counters, slow);
    S2G_CUDA(cudaGetLastError());
    return S2G_OK;
}

template <int KID>
int launch_gather(s2g_ctx* ctx, const GRec* recs, const unsigned* vals, const unsigned* tile_beg,
                  const unsigned* tile_end, const unsigned* chunk_begin, int ntiles, int ntile_j, unsigned total_chunks,
                  const s2g_particles& P, const s2g_geom& G, int image_k, double* image)
{
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    const int blocks = (int)std::min<long long>((long long)total_chunks, (long long)ctx->sm_count * G2D_CTAS);
    if (ctx->accum_f32)
        k_gather2d<KID, true><<<max(blocks, 1), G2D_THREADS, 0, ctx->stream>>>(recs, vals, tile_beg, tile_end, chunk_begin, ntiles,
                                                                      ntile_j, total_chunks, P.binq, P.in_dtype,
                                                                      G.n_images, image_k, G.npix, image,
                                                                      ctx->d_counters);
    else
        k_gather2d<KID, false><<<max(blocks, 1), G2D_THREADS, 0, ctx->stream>>>(recs, vals, tile_beg, tile_end, chunk_begin,
                                                                       ntiles, ntile_j, total_chunks, P.binq, P.in_dtype,
                                                                       G.n_images, image_k, G.npix, image,
                                                                       ctx->d_counters);
    S2G_CUDA(cudaGetLastError());
    return S2G_OK;
}

template <int KID>
KLaunch make_klaunch()
{
    KLaunch k;
    k.norm = launch_norm<KID>;
    k.gather = launch_gather<KID>;
    return k;
}

long long env_ll(const char* name, long long dflt)
{
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    return atoll(s);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// dispatch: classify -> scatter list + gather list -> scatter kernel / gather pipeline, in particle slices so the
// pair buffers stay bounded (a 1B-particle shard is walked in slices; the image accumulates across slices)
// ------------------------------------------------------------------------------------------------
namespace {
struct IsClass {
    int c;
    __host__ __device__ unsigned operator()(int v) const { return (unsigned)(v == c); }
};
struct ToU64 {
    __host__ __device__ unsigned long long operator()(unsigned v) const { return (unsigned long long)v; }
};
}  // namespace

int s2g_launch_deposit_2d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, double* image)
{
    if (P.n <= 0) return S2G_OK;
    KLaunch K;
    switch (kernel) {
    case S2G_KERNEL_CUBIC: K = make_klaunch<S2G_KERNEL_CUBIC>(); break;
    case S2G_KERNEL_QUINTIC: K = make_klaunch<S2G_KERNEL_QUINTIC>(); break;
    case S2G_KERNEL_WENDLAND_C2: K = make_klaunch<S2G_KERNEL_WENDLAND_C2>(); break;
    case S2G_KERNEL_WENDLAND_C4: K = make_klaunch<S2G_KERNEL_WENDLAND_C4>(); break;
    case S2G_KERNEL_WENDLAND_C6: K = make_klaunch<S2G_KERNEL_WENDLAND_C6>(); break;
    case S2G_KERNEL_WENDLAND_C8: K = make_klaunch<S2G_KERNEL_WENDLAND_C8>(); break;
    default: s2g_set_error("unknown kernel id %d", kernel); return S2G_EINVAL;
    }
    if (ctx->strategy == S2G_STRATEGY_SCATTER) {
        S2G_TRY(s2g_stage_wait(ctx, P.n));
        const int ph = s2g_phase_begin(ctx, PH_DEPOSIT);
        const int rc = s2g_launch_scatter_2d(ctx, P, G, kernel, nullptr, P.n, image);
        s2g_phase_end(ctx, ph);
        return rc;
    }

    const long long gather_min = env_ll("S2G_GATHER_MIN_PIXELS", 1024);
    const long long tiny_max = env_ll("S2G_TINY_MAX_PIXELS", 64);   // 0: no sub-warp bin
    const long long batch_max = env_ll("S2G_BATCH_PARTICLES", 8LL << 20);
    const long long pair_cap = env_ll("S2G_PAIR_CAP", 512LL << 20);
    const int exact_norm = (int)env_ll("S2G_EXACT_NORM", 0) || ctx->exact_norm;
    const int ntile_j = (int)((G.npix + TILE_W - 1) / TILE_W);
    const int ntile_i = (int)((G.npix + TILE_H - 1) / TILE_H);
    const int ntiles = ntile_i * ntile_j;
    cudaStream_t st = ctx->stream;

    long long p0 = 0;
    long long batch = min(batch_max, P.n);
    while (p0 < P.n) {
        long long nb = min(batch, P.n - p0);
        // overlapped staging (s2g_api.cu): a small first slice lets the deposit start while the rest is still copied;
        // every slice waits for exactly the particles it reads
        if (p0 == 0 && s2g_stage_first_slice(ctx) > 0) nb = min(nb, s2g_stage_first_slice(ctx));
        S2G_TRY(s2g_stage_wait(ctx, p0 + nb));
        void *d_cls, *d_np, *d_ps, *d_pg, *d_ls, *d_lg, *d_tmp, *d_sum;
        S2G_TRY(s2g_scratch(ctx, "g_cls", sizeof(int) * (nb + 1), &d_cls));
        S2G_TRY(s2g_scratch(ctx, "g_np", sizeof(unsigned) * (nb + 1), &d_np));
        S2G_TRY(s2g_scratch(ctx, "g_pos_s", sizeof(unsigned) * (nb + 1), &d_ps));
        S2G_TRY(s2g_scratch(ctx, "g_pos_g", sizeof(unsigned) * (nb + 1), &d_pg));
        S2G_TRY(s2g_scratch(ctx, "g_list_s", sizeof(int) * nb, &d_ls));
        S2G_TRY(s2g_scratch(ctx, "g_list_g", sizeof(int) * nb, &d_lg));
        void *d_pt, *d_lt;
        S2G_TRY(s2g_scratch(ctx, "g_pos_t", sizeof(unsigned) * (nb + 1), &d_pt));
        S2G_TRY(s2g_scratch(ctx, "g_list_t", sizeof(int) * nb, &d_lt));
        S2G_TRY(s2g_scratch(ctx, "g_sum", sizeof(unsigned long long), &d_sum));
        const int blocks = (int)((nb + 255) / 256);
        int ph = s2g_phase_begin(ctx, PH_PREP);
        S2G_CUDA(cudaMemsetAsync((int*)d_cls + nb, 0, sizeof(int), st));
        k_classify<<<blocks, 256, 0, st>>>(P, G, p0, nb, gather_min, tiny_max, ctx->strategy, (int*)d_cls,
                                           (unsigned*)d_np);
        S2G_CUDA(cudaGetLastError());

        cub::TransformInputIterator<unsigned, IsClass, const int*> it_s((const int*)d_cls, IsClass{1});
        cub::TransformInputIterator<unsigned, IsClass, const int*> it_g((const int*)d_cls, IsClass{2});
        cub::TransformInputIterator<unsigned, IsClass, const int*> it_t((const int*)d_cls, IsClass{3});
        cub::TransformInputIterator<unsigned long long, ToU64, const unsigned*> it_np((const unsigned*)d_np, ToU64{});
        size_t t1 = 0, t2 = 0, t3 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, t1, it_s, (unsigned*)d_ps, (int)(nb + 1), st);
        cub::DeviceScan::ExclusiveSum(nullptr, t2, it_g, (unsigned*)d_pg, (int)(nb + 1), st);
        cub::DeviceReduce::Sum(nullptr, t3, it_np, (unsigned long long*)d_sum, (int)nb, st);
        size_t tmp_bytes = max(t1, max(t2, t3)) + 16;
        S2G_TRY(s2g_scratch(ctx, "g_tmp", tmp_bytes, &d_tmp));
        size_t tb = tmp_bytes;
        S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, it_s, (unsigned*)d_ps, (int)(nb + 1), st));
        tb = tmp_bytes;
        S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, it_g, (unsigned*)d_pg, (int)(nb + 1), st));
        tb = tmp_bytes;
        S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, it_t, (unsigned*)d_pt, (int)(nb + 1), st));
        tb = tmp_bytes;
        S2G_CUDA(cub::DeviceReduce::Sum(d_tmp, tb, it_np, (unsigned long long*)d_sum, (int)nb, st));
        unsigned h_ns = 0, h_ng = 0, h_nt = 0;
        unsigned long long h_ub = 0;
        s2g_readback rb(ctx);
        S2G_CUDA(rb.add(&h_ns, (unsigned*)d_ps + nb, sizeof(unsigned)));
        S2G_CUDA(rb.add(&h_ng, (unsigned*)d_pg + nb, sizeof(unsigned)));
        S2G_CUDA(rb.add(&h_nt, (unsigned*)d_pt + nb, sizeof(unsigned)));
        S2G_CUDA(rb.add(&h_ub, d_sum, sizeof(unsigned long long)));
        s2g_phase_end(ctx, ph);
        ctx->launches += 4;
        S2G_CUDA(rb.sync());
        if ((long long)h_ub > pair_cap && nb > 1024) {  // too many (tile,particle) pairs: shrink the slice, redo it
            batch = std::max<long long>(1024, nb / 2);
            continue;
        }
        S2G_CHECK(h_ub < 0xfff00000ull, S2G_ENOMEM,
                  "a single slice of %lld particles spans %llu image tiles: footprints too large for this image",
                  nb, h_ub);
        const long long n_s = h_ns, n_g = h_ng, n_t = h_nt;
        k_build_lists<<<blocks, 256, 0, st>>>((const int*)d_cls, (const unsigned*)d_ps, (const unsigned*)d_pg,
                                              (const unsigned*)d_pt, p0, nb, (int*)d_ls, (int*)d_lg, (int*)d_lt);
        S2G_CUDA(cudaGetLastError());
        ctx->launches += 1;

        // ---- tiny bin (a few pixels): 8 lanes per particle
        if (n_t > 0) {
            ph = s2g_phase_begin(ctx, PH_DEPOSIT);
            S2G_TRY(s2g_launch_scatter_2d_tiny(ctx, P, G, kernel, (const int*)d_lt, n_t, image));
            s2g_phase_end(ctx, ph);
        }
        // ---- scatter bin (small footprints)
        if (n_s > 0) {
            ph = s2g_phase_begin(ctx, PH_DEPOSIT);
            S2G_TRY(s2g_launch_scatter_2d(ctx, P, G, kernel, (const int*)d_ls, n_s, image));
            s2g_phase_end(ctx, ph);
        }

        // ---- gather bin
        if (n_g > 0) {
            void *d_recs, *d_npg, *d_off, *d_rr;
            S2G_TRY(s2g_scratch(ctx, "g_reroute", sizeof(int) * n_g, &d_rr));
            S2G_TRY(s2g_scratch(ctx, "g_recs", sizeof(GRec) * n_g, &d_recs));
            S2G_TRY(s2g_scratch(ctx, "g_npg", sizeof(unsigned) * (n_g + 1), &d_npg));
            S2G_TRY(s2g_scratch(ctx, "g_off", sizeof(unsigned) * (n_g + 1), &d_off));
            S2G_CUDA(cudaMemsetAsync((unsigned*)d_npg + n_g, 0, sizeof(unsigned), st));
            ph = s2g_phase_begin(ctx, PH_NORM);
            S2G_TRY(K.norm(ctx, P, G, (const int*)d_lg, n_g, exact_norm, (GRec*)d_recs, (unsigned*)d_npg,
                           (int*)d_rr));
            s2g_phase_end(ctx, ph);
            ctx->launches += 1;
            ph = s2g_phase_begin(ctx, PH_SORT);
            size_t tb4 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb4, (const unsigned*)d_npg, (unsigned*)d_off, (int)(n_g + 1), st);
            S2G_TRY(s2g_scratch(ctx, "g_tmp", tb4 + 16, &d_tmp));
            S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb4, (const unsigned*)d_npg, (unsigned*)d_off,
                                                   (int)(n_g + 1), st));
            unsigned h_m = 0;
            unsigned long long h_rr = 0;
            S2G_CUDA(rb.add(&h_m, (unsigned*)d_off + n_g, sizeof(unsigned)));
            S2G_CUDA(rb.add(&h_rr, ctx->d_counters + CNT_PAIRS, sizeof(unsigned long long)));
            S2G_CUDA(rb.sync());
            const long long m = h_m;
            if (h_rr > 0) {  // "no pixel centre covered" particles found by pass A -> scatter kernel
                s2g_phase_end(ctx, ph);
                ph = s2g_phase_begin(ctx, PH_DEPOSIT);
                S2G_TRY(s2g_launch_scatter_2d(ctx, P, G, kernel, (const int*)d_rr, (long long)h_rr, image));
                s2g_phase_end(ctx, ph);
                ph = s2g_phase_begin(ctx, PH_SORT);
            }
            if (m > 0) {
                void *d_keys, *d_vals, *d_keys2, *d_vals2, *d_tcnt, *d_tbeg, *d_nch, *d_cbeg;
                S2G_TRY(s2g_scratch(ctx, "g_keys", sizeof(unsigned) * m, &d_keys));
                S2G_TRY(s2g_scratch(ctx, "g_vals", sizeof(unsigned) * m, &d_vals));
                S2G_TRY(s2g_scratch(ctx, "g_keys2", sizeof(unsigned) * m, &d_keys2));
                S2G_TRY(s2g_scratch(ctx, "g_vals2", sizeof(unsigned) * m, &d_vals2));
                S2G_TRY(s2g_scratch(ctx, "g_tcnt", sizeof(unsigned) * (ntiles + 1), &d_tcnt));
                S2G_TRY(s2g_scratch(ctx, "g_tbeg", sizeof(unsigned) * (ntiles + 1), &d_tbeg));
                S2G_TRY(s2g_scratch(ctx, "g_nch", sizeof(unsigned) * (ntiles + 1), &d_nch));
                S2G_TRY(s2g_scratch(ctx, "g_cbeg", sizeof(unsigned) * (ntiles + 1), &d_cbeg));
                k_expand<<<(int)((n_g + 255) / 256), 256, 0, st>>>((const GRec*)d_recs, (const unsigned*)d_off, n_g,
                                                                   ntile_j, (unsigned*)d_keys, (unsigned*)d_vals);
                S2G_CUDA(cudaGetLastError());
                int bits = 1;
                while ((1 << bits) < ntiles) ++bits;
                size_t sb = 0;
                cub::DeviceRadixSort::SortPairs(nullptr, sb, (const unsigned*)d_keys, (unsigned*)d_keys2,
                                                (const unsigned*)d_vals, (unsigned*)d_vals2, (int)m, 0, bits, st);
                S2G_TRY(s2g_scratch(ctx, "g_sort_tmp", sb + 16, &d_tmp));
                S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, (const unsigned*)d_keys, (unsigned*)d_keys2,
                                                         (const unsigned*)d_vals, (unsigned*)d_vals2, (int)m, 0, bits,
                                                         st));
                S2G_CUDA(cudaMemsetAsync(d_tcnt, 0, sizeof(unsigned) * (ntiles + 1), st));
                S2G_CUDA(cudaMemsetAsync(d_tbeg, 0, sizeof(unsigned) * (ntiles + 1), st));
                k_tile_bounds<<<(int)((m + 255) / 256), 256, 0, st>>>((const unsigned*)d_keys2, m, (unsigned*)d_tbeg,
                                                                      (unsigned*)d_tcnt);
                S2G_CUDA(cudaGetLastError());
                S2G_CUDA(cudaMemsetAsync(d_nch, 0, sizeof(unsigned) * (ntiles + 1), st));
                k_tile_chunks<<<(ntiles + 255) / 256, 256, 0, st>>>((const unsigned*)d_tbeg, (const unsigned*)d_tcnt,
                                                                    ntiles, (unsigned*)d_nch);
                S2G_CUDA(cudaGetLastError());
                size_t tb3 = 0;
                cub::DeviceScan::ExclusiveSum(nullptr, tb3, (const unsigned*)d_nch, (unsigned*)d_cbeg, ntiles + 1, st);
                S2G_TRY(s2g_scratch(ctx, "g_tmp", tb3 + 16, &d_tmp));
                size_t tbb = tb3 + 16;
                S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tbb, (const unsigned*)d_nch, (unsigned*)d_cbeg,
                                                       ntiles + 1, st));
                unsigned h_chunks = 0;
                S2G_CUDA(rb.add(&h_chunks, (unsigned*)d_cbeg + ntiles, sizeof(unsigned)));
                s2g_phase_end(ctx, ph);
                ph = -1;
                ctx->launches += 8;
                S2G_CUDA(rb.sync());
                const int phg = s2g_phase_begin(ctx, PH_DEPOSIT);
                for (int k = 0; k < G.n_images; ++k)
                    S2G_TRY(K.gather(ctx, (const GRec*)d_recs, (const unsigned*)d_vals2, (const unsigned*)d_tbeg,
                                     (const unsigned*)d_tcnt, (const unsigned*)d_cbeg, ntiles, ntile_j, h_chunks, P,
                                     G, k, image));
                s2g_phase_end(ctx, phg);
                ctx->launches += G.n_images;
                ctx->host_pairs += m;
            }
            if (ph >= 0) s2g_phase_end(ctx, ph);
        }
        p0 += nb;
    }
    return S2G_OK;
}
