// s2g_healpix.cu — HEALPix all-sky deposit (RING scheme), warp per particle.
//
// Replaces the particle loop of healpix_map (src/healpix_interpolation/main.jl:143-213):
//   contributing_pixels (constributing_pixels.jl:7-22) -> Healpix.jl vec2ang / queryDiscRing / ang2pix
//   calculate_weights   (pixel_weights.jl:87-140), weight_per_index (:34-76), contributing_area (:6-8),
//   distance_to_pixel_center (:16-22) -> Healpix.jl pix2vecRing
//   update_image!       (main.jl:25-45), particle_area_and_depth (main.jl:56-63)
// The RING arithmetic is the published HEALPix algorithm (ring_above / ring2z / get_ring_info / pix2ang_ring /
// ang2pix_ring / query_disc) that Healpix.jl ports; pixel numbers are 0-based (= Julia index - 1).
//
// Instead of materialising a pixel list per particle (the reference heap-allocates a Vector per particle) the
// warp walks the disc ring by ring: each ring contributes one contiguous (mod ring length) run of pixels, lanes
// stride over the run.  The pixel containing the particle centre is added when the disc walk did not visit it.
#include <cub/cub.cuh>
#include <cstdlib>

#include "s2g_healpix.cuh"

namespace {

// per-particle constants of the fast path
struct DiscFast {
    double ux, uy, uz, hinv, proj_h, half_ang, ang, a_scale;  // a_scale = 1/(ang_pix * (ang_pix*Dx)^2)
    double cl, sl, belt_inv_den;  // rotation by `lane` pixels of an equatorial-belt ring (all belt rings share Δφ)
    bool small;
};

// A (contributing area / (ang_pix Dx)^2) and kernel weight of the pixel whose centre is (st*cp, st*sp, ct)
template <int KID>
__device__ __forceinline__ void hp_pixel(const DiscFast& f, double st, double ez2, double cp, double sp, double& A,
                                         double& wk, bool& inside)
{
    const double ex = fma(st, cp, -f.ux), ey = fma(st, sp, -f.uy);
    const double c2 = fma(ex, ex, fma(ey, ey, ez2)) + 1e-300;  // squared chord between the unit vectors (> 0)
    const double hc = 0.5 * (c2 * hp_rsqrt(c2));                    // sin(dx/2)
    double dx;
    if (f.small) {  // asin series, |hc| <= 0.1
        const double x2 = 0.25 * c2;
        double ps = fma(x2, kAsinC[6], kAsinC[5]);
        ps = fma(ps, x2, kAsinC[4]);
        ps = fma(ps, x2, kAsinC[3]);
        ps = fma(ps, x2, kAsinC[2]);
        ps = fma(ps, x2, kAsinC[1]);
        ps = fma(ps, x2, kAsinC[0]);
        dx = 2.0 * fma(hc * x2, ps, hc);
    } else
        dx = 2.0 * asin(hc < 1.0 ? hc : 1.0);
    const double t = fma(-dx, f.hinv, 1.0);  // 1 - u
    inside = (t >= 0.0);
    const double inner = fabs(f.proj_h - (dx - f.half_ang));  // >= 0: max(0, min(ang, inner)) = min(ang, inner)
    A = (inner < f.ang ? inner : f.ang) * f.a_scale;
    const double w0 = hp_shape_t<KID>(t);
    wk = (t > 0.0) ? w0 : 0.0;
}

// walk the run [j0, j0+cnt) (mod nr) of one ring with the lanes along the ring; azimuth by exact sincospi for each
// lane's first pixel and by rotation through 32 pixel steps afterwards (re-seeded exactly every 8 steps)
template <int KID, bool PASS_B>
__device__ __forceinline__ void hp_walk_ring(const DiscFast& f, const RingTrig& rt, double c0, double s0, double c32,
                                             double s32, int sp, int nr, int j0, int cnt, int lane, int cpix,
                                             double area_norm, bool fb, double q, bool q_finite,
                                             double* __restrict__ amap, double* __restrict__ wmap, double& sw,
                                             double& sa, long long& n_in, bool& found_c)
{
    const double ez = rt.ct - f.uz, ez2 = ez * ez;
    const bool belt = (rt.inv_den == f.belt_inv_den);
    int nin = 0;
    for (int t0 = lane; t0 < cnt; t0 += 256) {  // exact azimuth every 8 steps of 32 pixels, rotation in between
        int j = j0 + t0;
        if (j >= nr) j -= nr;
        double cp, sph;
        if (belt && t0 < 32) {  // first pixel of the run rotated by `lane` pixels (belt rings share the pixel step)
            cp = fma(c0, f.cl, -s0 * f.sl);
            sph = fma(s0, f.cl, c0 * f.sl);
        } else
            sincospi(((double)(j + 1) - rt.off) * rt.inv_den, &sph, &cp);
        const int tend = min(t0 + 256, cnt);
        for (int t = t0; t < tend; t += 32) {
            double A, wk;
            bool inside;
            hp_pixel<KID>(f, rt.st, ez2, cp, sph, A, wk, inside);
            if (!PASS_B) {
                sa += A;
                sw = fma(wk, A, sw);  // wk = 0 outside
                nin += inside ? 1 : 0;
                found_c = found_c || (sp + j == cpix);
            } else {
                if (fb) wk = 1.0;
                const double pw = area_norm * wk * A;
                if (pw != 0.0 || !q_finite) {
                    red_add(amap + sp + j, q * pw);
                    red_add(wmap + sp + j, pw);
                }
            }
            const double c_new = fma(cp, c32, -sph * s32), s_new = fma(sph, c32, cp * s32);
            cp = c_new; sph = s_new;
            j += 32;
            if (j >= nr) j -= nr;
        }
    }
    n_in += nin;
}

// ---- pass B of a long run through the TMA: the warp stages the run's contributions (both maps) in shared memory,
// piece by piece (256 pixels), and lane 0 adds each contiguous piece to the maps with cp.reduce.async.bulk .add.f64
// (SASS UBLKRED) — 588 Gadd/s in 2-KiB operations against 272 Gred/s for per-lane red.global (profiles/
// r1_microbench_bulk_reduce.json).  Bulk operations need 16-byte aligned addresses and sizes: the staging index keeps
// the parity of the global pixel index, and an odd first / last pixel of a segment goes through red.global instead.
constexpr int HP_PIECE = 256;                 // pixels per staged piece
constexpr int HP_STAGE = HP_PIECE + 8;        // doubles per plane and slot (parity shifts)
struct __align__(16) HpStage {
    double v[2][2][HP_STAGE];                 // [slot][plane: 0 = q*pw, 1 = pw][pixel]
};

__device__ __forceinline__ void hp_bulk_add(double* gdst, const double* ssrc, int count)
{
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                 ::"l"(gdst), "r"(saddr), "r"(count * 8) : "memory");
}

template <int KID>
__device__ __forceinline__ void hp_walk_ring_bulk(const DiscFast& f, const RingTrig& rt, double c0, double s0,
                                                  double c32, double s32, int sp, int nr, int j0, int cnt, int lane,
                                                  double area_norm, bool fb, double q, HpStage& st, int& piece_no,
                                                  double* __restrict__ amap, double* __restrict__ wmap)
{
    const double ez = rt.ct - f.uz, ez2 = ez * ez;
    const bool belt = (rt.inv_den == f.belt_inv_den);
    for (int t0 = 0; t0 < cnt; t0 += HP_PIECE) {
        const int len = min(HP_PIECE, cnt - t0);
        // the piece covers in-ring indices js .. js+len-1 (mod nr): segment 1 = [js, e1), segment 2 = [0, len2) if it wraps
        int js = j0 + t0;
        if (js >= nr) js -= nr;
        const int len1 = min(len, nr - js), len2 = len - len1;
        const int par1 = (sp + js) & 1, par2 = sp & 1;
        const int off2 = (len1 + par1 + 1) & ~1;   // staging offset of segment 2 (even)
        const int slot = piece_no & 1;
        if (piece_no >= 2) {  // the bulk group that read this slot two pieces ago must be done with it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
        if (lane < len) {
            int j = js + lane;
            if (j >= nr) j -= nr;
            double cp, sph;
            if (belt && t0 == 0) {
                cp = fma(c0, f.cl, -s0 * f.sl);
                sph = fma(s0, f.cl, c0 * f.sl);
            } else
                sincospi(((double)(j + 1) - rt.off) * rt.inv_den, &sph, &cp);
            for (int t = lane; t < len; t += 32) {
                double A, wk;
                bool inside;
                hp_pixel<KID>(f, rt.st, ez2, cp, sph, A, wk, inside);
                if (fb) wk = 1.0;
                const double pw = area_norm * wk * A;
                // staging position with the parity of the global pixel index; unaligned ends go through red.global
                const bool seg2 = t >= len1;
                const int k = seg2 ? t - len1 : t;              // index inside the segment
                const int slen = seg2 ? len2 : len1, par = seg2 ? par2 : par1;
                const bool head = (k == 0) && (par == 1);
                const bool tail = (k == slen - 1) && (((par + slen) & 1) == 1);
                if (head || tail) {
                    red_add(amap + sp + j, q * pw);
                    red_add(wmap + sp + j, pw);
                } else {
                    const int idx = (seg2 ? off2 : 0) + par + k;
                    st.v[slot][0][idx] = q * pw;
                    st.v[slot][1][idx] = pw;
                }
                const double c_new = fma(cp, c32, -sph * s32), s_new = fma(sph, c32, cp * s32);
                cp = c_new; sph = s_new;
                j += 32;
                if (j >= nr) j -= nr;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            // aligned interior of segment 1: pixels js+par1 .. (even count)
            const int a1 = par1, n1 = (len1 - par1) & ~1;
            if (n1 > 0) {
                hp_bulk_add(amap + sp + js + a1, &st.v[slot][0][2 * a1], n1);
                hp_bulk_add(wmap + sp + js + a1, &st.v[slot][1][2 * a1], n1);
            }
            if (len2 > 0) {
                const int a2 = par2, n2 = (len2 - par2) & ~1;
                if (n2 > 0) {
                    hp_bulk_add(amap + sp + a2, &st.v[slot][0][off2 + 2 * a2], n2);
                    hp_bulk_add(wmap + sp + a2, &st.v[slot][1][off2 + 2 * a2], n2);
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++piece_no;
    }
}

// per-warp staging of the set-up of up to 32 rings (one ring per lane), so that the expensive uniform part of the
// disc walk (ring_above / atan2 / sqrt per ring) is computed ONCE per ring by one lane instead of by all 32
struct RingBatch {
    int sp[32], nr[32], j0[32], cnt[32], pre[33];
    double st[32], ct[32], off[32], inv_den[32];
    double c0[32], s0[32], c32[32], s32[32];  // azimuth of the run's first pixel; rotation by 32 pixels
};

template <int KID, bool PASS_B>
__device__ __forceinline__ void hp_process_batch(const HpGeom& g, const Disc& d, const DiscFast& f, RingBatch& rb,
                                                 long long ring0, int nb, bool setup, int lane, double area_norm,
                                                 bool fb, double q, bool q_finite, double* __restrict__ amap,
                                                 double* __restrict__ wmap, double& sw, double& sa, long long& n_in,
                                                 long long& n_tot, bool& found_c, HpStage& stg, int& piece_no,
                                                 bool use_bulk)
{
    if (setup) {
        int cnt_l = 0;
        if (lane < nb) {
            const long long ring = ring0 + lane;
            long long sp, nr, j0, cnt;
            bool sh;
            hp_ring_info(g, ring, sp, nr, sh);
            ring_run(g, d, ring, nr, sh, j0, cnt);
            const RingTrig rt = hp_ring_trig(g, ring);
            rb.sp[lane] = (int)sp; rb.nr[lane] = (int)nr; rb.j0[lane] = (int)j0; rb.cnt[lane] = (int)cnt;
            rb.st[lane] = rt.st; rb.ct[lane] = rt.ct; rb.off[lane] = rt.off; rb.inv_den[lane] = rt.inv_den;
            double sa_, ca_;
            sincospi(((double)(j0 + 1) - rt.off) * rt.inv_den, &sa_, &ca_);
            rb.c0[lane] = ca_; rb.s0[lane] = sa_;
            sincospi(32.0 * rt.inv_den, &sa_, &ca_);
            rb.c32[lane] = ca_; rb.s32[lane] = sa_;
            cnt_l = (int)cnt;
        }
        // exclusive prefix of the run lengths (flattened walk) and the batch total
        int incl = cnt_l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        rb.pre[lane] = incl - cnt_l;
        if (lane == 31) rb.pre[32] = incl;
        __syncwarp();
    }
    const int total = rb.pre[32];
    if (!PASS_B) n_tot += total;
    if (total == 0) return;
    if (total >= 24 * nb) {
        // long runs: ring by ring, lanes along the ring, azimuth by rotation
        for (int r = 0; r < nb; ++r) {
            const int cnt = rb.cnt[r];
            if (cnt == 0) continue;
            RingTrig rt;
            rt.st = rb.st[r]; rt.ct = rb.ct[r]; rt.off = rb.off[r]; rt.inv_den = rb.inv_den[r];
            if (PASS_B && use_bulk && cnt >= 64 && q_finite) {
                hp_walk_ring_bulk<KID>(f, rt, rb.c0[r], rb.s0[r], rb.c32[r], rb.s32[r], rb.sp[r], rb.nr[r], rb.j0[r], cnt,
                                       lane, area_norm, fb, q, stg, piece_no, amap, wmap);
                continue;
            }
            hp_walk_ring<KID, PASS_B>(f, rt, rb.c0[r], rb.s0[r], rb.c32[r], rb.s32[r], rb.sp[r], rb.nr[r], rb.j0[r], cnt,
                                      lane, (int)d.cpix, area_norm, fb, q, q_finite, amap, wmap, sw, sa, n_in, found_c);
        }
    } else {
        // short runs: flatten (ring, pixel) over the lanes so that small discs still fill the warp
        for (int t = lane; t < total; t += 32) {
            int lo = 0, hi = nb;  // last ring with pre[ring] <= t
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int mid = (lo + hi) >> 1;
                if (hi - lo > 1) { if (rb.pre[mid] <= t) lo = mid; else hi = mid; }
            }
            const int r = lo;
            int j = rb.j0[r] + (t - rb.pre[r]);
            if (j >= rb.nr[r]) j -= rb.nr[r];
            double sp_, cp_;
            sincospi(((double)(j + 1) - rb.off[r]) * rb.inv_den[r], &sp_, &cp_);
            const double ez = rb.ct[r] - f.uz;
            double A, wk;
            bool inside;
            hp_pixel<KID>(f, rb.st[r], ez * ez, cp_, sp_, A, wk, inside);
            const long long pix = (long long)rb.sp[r] + j;
            if (!PASS_B) {
                sa += A;
                if (inside) { sw = fma(wk, A, sw); ++n_in; }
                if (pix == d.cpix) found_c = true;
            } else {
                if (fb) wk = 1.0;
                const double pw = area_norm * wk * A;
                if (pw != 0.0 || !q_finite) {
                    red_add(amap + pix, q * pw);
                    red_add(wmap + pix, pw);
                }
            }
        }
    }
}

// COOP = false: one warp per particle (dynamic queue over all particles, skipping those flagged `heavy`).
// COOP = true : one CTA (8 warps) per particle of `heavy_list` — the ring batches of both passes are dealt round-robin
//               to the warps and the pass-A sums are combined through shared memory.  A disc of ~10^6 pixels (a
//               particle close to the observer) otherwise occupies a single warp for tens of milliseconds and the
//               whole device waits for it.
// REC = true : pass A only, one warp per particle of `heavy_list[0 .. n_list_val)` (the tile-gather's list): writes
//               recs[t] instead of depositing (pass B is k_hp_gather, s2g_hpgather.cu).  Particles in the
//               `distr_weight == 0` branch or with a non-finite normalisation get an unused record and skip[p] = 0,
//               which hands them to the ordinary scatter launch that follows.
template <int KID, bool COOP, bool REC>
__global__ void __launch_bounds__(256, 2) k_healpix(s2g_particles P, HpGeom g, int calc_mean,
                                                    const unsigned char* __restrict__ take,
                                                    const unsigned* __restrict__ order,
                                                    const unsigned char* __restrict__ heavy,
                                                    const unsigned* __restrict__ heavy_list,
                                                    const unsigned* __restrict__ n_heavy, bool use_bulk,
                                                    double* __restrict__ amap, double* __restrict__ wmap,
                                                    unsigned long long* __restrict__ counters,
                                                    long long n_list_val, HRec* __restrict__ recs,
                                                    unsigned char* __restrict__ skip_out)
{
    __shared__ long long s_pick;
    __shared__ double s_redd[8][2];
    __shared__ long long s_redl[8][2];
    constexpr int NW = COOP ? 8 : 1;          // warps per particle
    constexpr long long BR = COOP ? 4 : 32;   // rings per batch: small batches interleave the short polar and the
                                              // long equatorial rings of a big disc evenly over the 8 warps
    extern __shared__ __align__(16) unsigned char hp_smem[];
    RingBatch* s_rb = reinterpret_cast<RingBatch*>(hp_smem);
    HpStage* s_st = reinterpret_cast<HpStage*>(hp_smem + 8 * sizeof(RingBatch));
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    RingBatch& rb = s_rb[wq];
    HpStage& stg = s_st[wq];
    int piece_no = 0;
    unsigned long long touched = 0, fallback = 0, mapped = 0;
    const double belt_inv_den = 1.0 / __dmul_rn(2.0, (double)g.nside);
    double lane_c, lane_s;
    sincospi((double)lane * belt_inv_den, &lane_s, &lane_c);
    const int wsub = COOP ? wq : 0;           // this warp's slot among the warps sharing the particle
    for (;;) {
        long long p = 0, t_rec = 0;
        if (COOP) {
            __syncthreads();                   // previous particle's shared sums are consumed
            if (threadIdx.x == 0) s_pick = (long long)atomicAdd(&counters[CNT_WORK], 1ull);
            __syncthreads();
            p = s_pick;
            if (p >= (REC ? n_list_val : (long long)*n_heavy)) break;
            t_rec = p;
            p = heavy_list[p];
            if (REC && threadIdx.x == 0) { recs[t_rec].rmin = 1; recs[t_rec].rmax = 0; recs[t_rec].ntot = 0; }
        } else if (REC) {
            if (lane == 0) p = (long long)atomicAdd(&counters[CNT_WORK], 1ull);
            p = __shfl_sync(0xffffffffu, p, 0);
            if (p >= n_list_val) break;
            t_rec = p;
            p = heavy_list[p];
            if (lane == 0) { recs[t_rec].rmin = 1; recs[t_rec].rmax = 0; recs[t_rec].ntot = 0; }  // unused until proven useful
        } else {
            if (lane == 0) p = (long long)atomicAdd(&counters[CNT_WORK], 1ull);
            p = __shfl_sync(0xffffffffu, p, 0);
            if (p >= P.n) break;
            if (order) p = order[p];           // processing order: by sky region (L2 locality of the map updates)
            if (heavy && heavy[p]) continue;   // done by the cooperative launch / the tile-gather
        }
        if (take && !take[p]) continue;        // not selected by filter_sort_particles
        const double q = ld_in(P.binq, p, P.in_dtype);
        if (!calc_mean && q == 0.0) continue;  // main.jl:160-165
        Disc d;
        d.px = ld_pos(P, p, 0); d.py = ld_pos(P, p, 1); d.pz = ld_pos(P, p, 2);
        const double hs = ld_in(P.hsml, p, P.in_dtype);
        // get_norm (shared.jl:1-10): Σ pos[dim]^2 from zero, then sqrt
        d.Dx = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d.px, d.px), __dmul_rn(d.py, d.py)), __dmul_rn(d.pz, d.pz)));
        if (d.Dx < hs) continue;  // main.jl:172-174
        d.proj_h = asin(__ddiv_rn(hs, d.Dx));
        d.hinv = __ddiv_rn(1.0, d.proj_h);
        make_disc(g, d);
        d.ux = d.px / d.Dx; d.uy = d.py / d.Dx; d.uz = d.pz / d.Dx;
        d.inv_ang = 1.0 / g.ang_pix;
        {
            const double aD0 = g.ang_pix * d.Dx;
            d.inv_aD2 = 1.0 / (aD0 * aD0);
        }
        d.small = (d.proj_h + 2.0 * g.ang_pix) < 0.2;
        DiscFast f;
        f.ux = d.ux; f.uy = d.uy; f.uz = d.uz; f.hinv = d.hinv; f.proj_h = d.proj_h;
        f.half_ang = 0.5 * g.ang_pix; f.ang = g.ang_pix; f.a_scale = d.inv_ang * d.inv_aD2; f.small = d.small;
        f.cl = lane_c; f.sl = lane_s; f.belt_inv_den = belt_inv_den;
        const bool q_finite = isfinite(q);
        const long long nrings = d.ring_last - d.ring_first + 1;

        // ---- pass A (calculate_weights, pixel_weights.jl:87-140)
        double sw = 0.0, sa = 0.0;
        long long n_in = 0, n_tot = 0;
        bool found_c = false;
        __syncwarp();
        for (long long base = BR * wsub; base < nrings; base += BR * NW) {
            const int nb = (int)min(BR, nrings - base);
            hp_process_batch<KID, false>(g, d, f, rb, d.ring_first + base, nb, true, lane, 0.0, false, q, q_finite, amap,
                                         wmap, sw, sa, n_in, n_tot, found_c, stg, piece_no, false);
            if (base + BR * NW < nrings) __syncwarp();
        }
        found_c = __any_sync(0xffffffffu, found_c);
        if (COOP) found_c = __syncthreads_or(found_c ? 1 : 0) != 0;
        // the centre pixel, when the disc walk did not visit it (push! + unique!)
        double cA = 0.0, cwk = 0.0;
        bool c_inside = false;
        if (!found_c) {
            long long ring;  // ring of cpix
            if (d.cpix < g.ncap) {
                ring = (long long)((1 + (long long)floor(sqrt((double)(1 + 2 * d.cpix)))) >> 1);
                while (2 * ring * (ring - 1) > d.cpix) --ring;
                while (2 * ring * (ring + 1) <= d.cpix) ++ring;
            } else if (d.cpix < g.npix - g.ncap) {
                ring = (d.cpix - g.ncap) / g.nl4 + g.nside;
            } else {
                const long long rem = g.npix - 1 - d.cpix;  // 0-based from the end
                long long rs = (long long)((1 + (long long)floor(sqrt((double)(1 + 2 * rem)))) >> 1);
                while (2 * rs * (rs - 1) > rem) --rs;
                while (2 * rs * (rs + 1) <= rem) ++rs;
                ring = g.nl4 - rs;
            }
            long long sp, nr;
            bool sh;
            hp_ring_info(g, ring, sp, nr, sh);
            const RingTrig c_rt = hp_ring_trig(g, ring);
            double s_, c_;
            sincospi(((double)(d.cpix - sp + 1) - c_rt.off) * c_rt.inv_den, &s_, &c_);
            const double ez = c_rt.ct - f.uz;
            hp_pixel<KID>(f, c_rt.st, ez * ez, c_, s_, cA, cwk, c_inside);
            if (lane == 0 && wsub == 0) {
                sa += cA;
                if (c_inside) { sw = fma(cwk, cA, sw); ++n_in; }
            }
            if (wsub == 0) ++n_tot;
        }
        sw = warp_sum(sw);
        sa = warp_sum(sa);
        n_in = warp_sum_ll(n_in);
        if (COOP) {  // combine the warps' partial sums, every warp in the same order -> identical normalisation
            if (lane == 0) { s_redd[wq][0] = sw; s_redd[wq][1] = sa; s_redl[wq][0] = n_in; s_redl[wq][1] = n_tot; }
            __syncthreads();
            sw = 0.0; sa = 0.0; n_in = 0; n_tot = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { sw += s_redd[k][0]; sa += s_redd[k][1]; n_in += s_redl[k][0]; n_tot += s_redl[k][1]; }
        }

        // ---- normalisation (pixel_weights.jl:119-137, main.jl:32-33, :188-193)
        bool fb = false;
        double n_distr, wpp;
        if (sw == 0.0) {
            fb = true;
            n_distr = (double)n_tot;
            wpp = (sa != 0.0) ? n_distr / sa : 1.0;
            if (lane == 0 && wsub == 0) ++fallback;
        } else {
            n_distr = (double)n_in;
            wpp = n_distr / sw;
        }
        double dz = __dmul_rn(2.0, hs);
        const double area = __ddiv_rn(__ddiv_rn(ld_in(P.m, p, P.in_dtype), ld_in(P.rho, p, P.in_dtype)), dz);
        const double aD = __dmul_rn(g.ang_pix, d.Dx);
        dz = __ddiv_rn(dz, __dmul_rn(aD, aD));
        const double kernel_norm = area / n_distr;
        const double area_norm = kernel_norm * wpp * ld_in(P.w, p, P.in_dtype) * dz;
        if (REC) {
            const double an = area_norm * d.inv_aD2;
            const bool usable = !fb && q_finite && isfinite(an) && found_c && d.ring_first == d.irmin &&
                                d.ring_last == d.irmax && !d.full_sky;
            if (lane == 0 && wsub == 0) {
                if (usable) {
                    HRec r;
                    r.ux = d.ux; r.uy = d.uy; r.uz = d.uz; r.ph = d.proj_h; r.an = an; r.anq = an * q;
                    r.rmin = (int)d.irmin; r.rmax = (int)d.irmax; r.ntot = (int)n_tot; r.pad = 0;
                    recs[t_rec] = r;
                } else {
                    skip_out[p] = 0;   // the scatter launch that follows deposits (and counts) it
                    if (fb && fallback) --fallback;
                }
            }
            continue;
        }

        // ---- pass B (update_image!, main.jl:25-45); a single batch is still staged in shared memory
        {
            double d0 = 0.0, d1 = 0.0;
            long long l0 = 0, l1 = 0;
            bool b0 = false;
            for (long long base = BR * wsub; base < nrings; base += BR * NW) {
                const int nb = (int)min(BR, nrings - base);
                __syncwarp();
                hp_process_batch<KID, true>(g, d, f, rb, d.ring_first + base, nb, COOP || nrings > 32, lane, area_norm, fb, q,
                                            q_finite, amap, wmap, d0, d1, l0, l1, b0, stg, piece_no, use_bulk);
            }
        }
        if (wsub != 0) continue;  // per-particle bookkeeping and the centre pixel: once
        if (lane == 0) touched += (unsigned long long)n_tot;
        if (!found_c && lane == 0) {
            const double pw = area_norm * (fb ? 1.0 : cwk) * cA;
            if (pw != 0.0 || !q_finite) {
                red_add(amap + d.cpix, q * pw);
                red_add(wmap + d.cpix, pw);
            }
        }
        if (lane == 0) ++mapped;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // outstanding bulk reductions
    if (lane == 0) {
        if (touched) { atomicAdd(&counters[CNT_TOUCHED], touched); atomicAdd(&counters[CNT_FOOTPRINT], touched); }
        if (fallback) atomicAdd(&counters[CNT_FALLBACK], fallback);
        if (mapped) { atomicAdd(&counters[CNT_MAPPED], mapped); atomicAdd(&counters[CNT_SCATTER], mapped); }
    }
}

// single-thread pixel list in the reference's order (test hook for the bit-exact pixel-set contract)
__global__ void k_healpix_pixels(double px, double py, double pz, double radius, HpGeom g, long long* out,
                                 long long cap, long long* count)
{
    Disc d;
    d.px = px; d.py = py; d.pz = pz;
    d.Dx = 1.0;
    d.proj_h = radius;
    d.hinv = 1.0 / radius;
    make_disc(g, d);
    long long n = 0;
    bool found = false;
    for (long long ring = d.ring_first; ring <= d.ring_last; ++ring) {
        long long sp, nr, j0, cnt;
        bool sh;
        hp_ring_info(g, ring, sp, nr, sh);
        ring_run(g, d, ring, nr, sh, j0, cnt);
        // reference order inside a ring: [0..ip_hi] first, then the wrapped part (query_disc appends that way)
        if (cnt > 0 && j0 + cnt > nr) {
            for (long long j = 0; j < j0 + cnt - nr; ++j) { if (n < cap) out[n] = sp + j; ++n; if (sp + j == d.cpix) found = true; }
            for (long long j = j0; j < nr; ++j) { if (n < cap) out[n] = sp + j; ++n; if (sp + j == d.cpix) found = true; }
        } else
            for (long long t = 0; t < cnt; ++t) { if (n < cap) out[n] = sp + j0 + t; ++n; if (sp + j0 + t == d.cpix) found = true; }
    }
    if (!found) { if (n < cap) out[n] = d.cpix; ++n; }
    *count = n;
}

// sky-region key of a particle: its RING pixel at Nside 32 (12288 regions)
__global__ void __launch_bounds__(256) k_hp_order_keys(s2g_particles P, HpGeom gc, unsigned* __restrict__ keys,
                                                       unsigned* __restrict__ idx)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const double x = ld_pos(P, p, 0), y = ld_pos(P, p, 1), z = ld_pos(P, p, 2);
    const double nrm = sqrt(x * x + y * y + z * z);
    unsigned k = 0;
    if (nrm > 0.0) {
        double ph = atan2(y, x);
        if (ph < 0) ph += kTwoPi;
        k = (unsigned)hp_ang2pix_ring(gc, acos(z / nrm), ph);
    }
    keys[p] = k;
    idx[p] = (unsigned)p;
}

struct HpIsVal {
    unsigned char v;
    __host__ __device__ unsigned char operator()(unsigned char x) const { return x == v ? 1 : 0; }
};

template <int KID>
int set_smem_attr(s2g_ctx* ctx, size_t smem)
{
    static bool attr_set[64] = {};  // function attributes are per device; a process may drive several (s2g_group.cu)
    if (!attr_set[ctx->device & 63]) {
        S2G_CUDA(cudaFuncSetAttribute(k_healpix<KID, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        S2G_CUDA(cudaFuncSetAttribute(k_healpix<KID, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        S2G_CUDA(cudaFuncSetAttribute(k_healpix<KID, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        S2G_CUDA(cudaFuncSetAttribute(k_healpix<KID, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ctx->device & 63] = true;
    }
    return S2G_OK;
}

template <int KID>
int launch_healpix_k(s2g_ctx* ctx, const s2g_particles& P, long long nside, int calc_mean, const unsigned char* take,
                     double* amap, double* wmap)
{
    const HpGeom g = make_hp(nside);
    // processing order by sky region, so that the ~2400 particles in flight update a common few-MB patch of the maps
    const unsigned* order = nullptr;
    const char* e_ord = getenv("S2G_HP_ORDER");
    const char* e_blk = getenv("S2G_HP_BULK");
    // Both are OFF by default: measured on the c4s sample (2.2 M particles, Nside 2048) they cost time —
    // 640 ms (off/off) vs 716 ms (bulk) vs 681 ms (ordered) vs 758 ms (both); profiles/r1_healpix_experiments.txt.
    // The kernel is issue/latency bound there, not red-bound; sky-ordered particles also contend on the same pixels.
    const bool want_order = e_ord ? atoi(e_ord) != 0 : false;
    const bool use_bulk = e_blk ? atoi(e_blk) != 0 : false;
    if (want_order && P.n >= 65536) {
        void *d_k, *d_k2, *d_i, *d_i2, *d_tmp;
        S2G_TRY(s2g_scratch(ctx, "hp_okeys", sizeof(unsigned) * P.n, &d_k));
        S2G_TRY(s2g_scratch(ctx, "hp_okeys2", sizeof(unsigned) * P.n, &d_k2));
        S2G_TRY(s2g_scratch(ctx, "hp_oidx", sizeof(unsigned) * P.n, &d_i));
        S2G_TRY(s2g_scratch(ctx, "hp_oidx2", sizeof(unsigned) * P.n, &d_i2));
        const int ph = s2g_phase_begin(ctx, PH_SORT);
        k_hp_order_keys<<<(int)((P.n + 255) / 256), 256, 0, ctx->stream>>>(P, make_hp(32), (unsigned*)d_k, (unsigned*)d_i);
        S2G_CUDA(cudaGetLastError());
        size_t sb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sb, (const unsigned*)d_k, (unsigned*)d_k2, (const unsigned*)d_i,
                                        (unsigned*)d_i2, (int)P.n, 0, 14, ctx->stream);
        S2G_TRY(s2g_scratch(ctx, "g_sort_tmp", sb + 16, &d_tmp));
        S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, (const unsigned*)d_k, (unsigned*)d_k2, (const unsigned*)d_i,
                                                 (unsigned*)d_i2, (int)P.n, 0, 14, ctx->stream));
        s2g_phase_end(ctx, ph);
        ctx->launches += 4;
        order = (const unsigned*)d_i2;
    }
    const size_t smem = 8 * (sizeof(RingBatch) + sizeof(HpStage));
    S2G_TRY(set_smem_attr<KID>(ctx, smem));
    // Three classes (k_hp_classify, s2g_hpgather.cu):
    //   heavy  : very large discs (close to the observer) first, one CTA each — S2G_HP_COOP_RINGS (default 512 rings =
    //            an angular radius of 256 pixels, >= 2*10^5 pixels per particle; 0 switches the split off)
    //   gather : resolved discs (radius >= S2G_HP_GATHER_MIN_PIXELS pixel diameters, default 12; away from the poles;
    //            below 0.2 rad) -> tile-gather, no atomics in the inner loop (strategy AUTO / GATHER; S2G_HP_GATHER=0 off)
    //   rest   : the warp-per-particle scatter walk below
    long long coop_rings = 512;
    if (const char* e = getenv("S2G_HP_COOP_RINGS")) coop_rings = atoll(e);
    const bool coop_on = coop_rings > 0 && 0.5 * (double)coop_rings * g.ang_pix < 1.5;
    bool gather_on = ctx->strategy != S2G_STRATEGY_SCATTER;
    if (const char* e = getenv("S2G_HP_GATHER")) gather_on = gather_on && atoi(e) != 0;
    double gather_min = 12.0;
    if (const char* e = getenv("S2G_HP_GATHER_MIN_PIXELS")) gather_min = atof(e);
    if (ctx->strategy == S2G_STRATEGY_GATHER) gather_min = std::min(gather_min, 3.0);
    const unsigned char* d_skip = nullptr;
    if (coop_on || gather_on) {
        void *d_h, *d_g, *d_gh, *d_s, *d_list, *d_nh, *d_tmp;
        S2G_TRY(s2g_scratch(ctx, "hp_heavy", (size_t)P.n, &d_h));
        S2G_TRY(s2g_scratch(ctx, "hp_gath", (size_t)P.n, &d_g));
        S2G_TRY(s2g_scratch(ctx, "hp_gathh", (size_t)P.n, &d_gh));
        S2G_TRY(s2g_scratch(ctx, "hp_skip", (size_t)P.n, &d_s));
        S2G_TRY(s2g_scratch(ctx, "hp_heavy_list", sizeof(unsigned) * (size_t)P.n, &d_list));
        S2G_TRY(s2g_scratch(ctx, "hp_heavy_n", sizeof(unsigned) * 4, &d_nh));
        const double heavy_radius = coop_on ? 0.5 * (double)coop_rings * g.ang_pix : 0.0;
        int php = s2g_phase_begin(ctx, PH_PREP);
        S2G_TRY(s2g_hp_classify(ctx, P, nside, calc_mean, take, heavy_radius, gather_min * g.ang_pix, gather_on ? 1 : 0,
                                (unsigned char*)d_h, (unsigned char*)d_g, (unsigned char*)d_gh, (unsigned char*)d_s));
        cub::CountingInputIterator<unsigned> ids(0u);
        size_t tb = 0;
        cub::DeviceSelect::Flagged(nullptr, tb, ids, (const unsigned char*)d_h, (unsigned*)d_list, (unsigned*)d_nh,
                                   (int)P.n, ctx->stream);
        S2G_TRY(s2g_scratch(ctx, "hp_heavy_tmp", tb + 16, &d_tmp));
        s2g_phase_end(ctx, php);
        if (coop_on) {
            // heavy discs the gather cannot take (over a pole, non-finite quantity): one CTA each, scatter
            php = s2g_phase_begin(ctx, PH_PREP);
            S2G_CUDA(cub::DeviceSelect::Flagged(d_tmp, tb, ids, (const unsigned char*)d_h, (unsigned*)d_list,
                                                (unsigned*)d_nh, (int)P.n, ctx->stream));
            s2g_phase_end(ctx, php);
            S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
            const int phc = s2g_phase_begin(ctx, PH_DEPOSIT);
            k_healpix<KID, true, false><<<ctx->sm_count * 2, 256, smem, ctx->stream>>>(
                P, g, calc_mean, take, nullptr, nullptr, (const unsigned*)d_list, (const unsigned*)d_nh, use_bulk, amap,
                wmap, ctx->d_counters, 0, nullptr, nullptr);
            s2g_phase_end(ctx, phc);
            S2G_CUDA(cudaGetLastError());
            ctx->launches += 3;
        }
        if (gather_on) {
            // the two gather lists reuse the list buffer (everything is stream-ordered): first the heavy discs
            // (pass A by a whole CTA each), then the ordinary ones (pass A by a warp each)
            for (int pass = 0; pass < 3; ++pass) {
                // 0: discs above 0.2 rad (asin), 1: discs from 0.073 rad (gath == 2, 8 series coefficients), 2: small-angle
                // discs (gath == 1, the bulk: 5 coefficients)
                const bool heavy_pass = pass == 0;
                const int phl = s2g_phase_begin(ctx, PH_PREP);
                if (heavy_pass) {
                    S2G_CUDA(cub::DeviceSelect::Flagged(d_tmp, tb, ids, (const unsigned char*)d_gh, (unsigned*)d_list,
                                                        (unsigned*)d_nh + 1, (int)P.n, ctx->stream));
                } else {
                    cub::TransformInputIterator<unsigned char, HpIsVal, const unsigned char*> flags(
                        (const unsigned char*)d_g, HpIsVal{(unsigned char)(pass == 1 ? 2 : 1)});
                    size_t tb2 = 0;
                    cub::DeviceSelect::Flagged(nullptr, tb2, ids, flags, (unsigned*)d_list, (unsigned*)d_nh + 1, (int)P.n,
                                               ctx->stream);
                    void* d_tmp2;
                    S2G_TRY(s2g_scratch(ctx, "hp_heavy_tmp2", tb2 + 16, &d_tmp2));
                    S2G_CUDA(cub::DeviceSelect::Flagged(d_tmp2, tb2, ids, flags, (unsigned*)d_list, (unsigned*)d_nh + 1,
                                                        (int)P.n, ctx->stream));
                }
                unsigned h_ng = 0;
                S2G_CUDA(cudaMemcpyAsync(&h_ng, (unsigned*)d_nh + 1, sizeof(unsigned), cudaMemcpyDeviceToHost,
                                         ctx->stream));
                s2g_phase_end(ctx, phl);
                S2G_CUDA(cudaStreamSynchronize(ctx->stream));
                ctx->launches += 1;
                const int big = heavy_pass ? 1 : 0;   // the classification sends every disc above 0.2 rad to pass 0
                S2G_TRY(s2g_hp_gather_pipeline(ctx, P, nside, KID, calc_mean, (const unsigned*)d_list, (long long)h_ng,
                                               (unsigned char*)d_s, amap, wmap, heavy_pass ? 1 : 0, big, pass == 2 ? 5 : 8));
            }
        }
        d_skip = (const unsigned char*)d_s;
    }
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    int blocks = (int)std::min<long long>((P.n + 7) / 8, (long long)ctx->sm_count * 2);
    const int ph = s2g_phase_begin(ctx, PH_DEPOSIT);
    k_healpix<KID, false, false><<<max(blocks, 1), 256, smem, ctx->stream>>>(P, g, calc_mean, take, order, d_skip, nullptr,
                                                                             nullptr, use_bulk, amap, wmap,
                                                                             ctx->d_counters, 0, nullptr, nullptr);
    s2g_phase_end(ctx, ph);
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

template <int KID>
int launch_records_k(s2g_ctx* ctx, const s2g_particles& P, long long nside, int calc_mean, const unsigned* list,
                     long long n_list, HRec* recs, unsigned char* skip, int coop)
{
    const HpGeom g = make_hp(nside);
    const size_t smem = 8 * (sizeof(RingBatch) + sizeof(HpStage));
    S2G_TRY(set_smem_attr<KID>(ctx, smem));
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    if (coop) {   // one CTA per (very large) disc
        const int blocks = (int)std::min<long long>(n_list, (long long)ctx->sm_count * 2);
        k_healpix<KID, true, true><<<max(blocks, 1), 256, smem, ctx->stream>>>(P, g, calc_mean, nullptr, nullptr, nullptr,
                                                                               list, nullptr, false, nullptr, nullptr,
                                                                               ctx->d_counters, n_list, recs, skip);
        S2G_CUDA(cudaGetLastError());
        ctx->launches += 1;
        return S2G_OK;
    }
    const int blocks = (int)std::min<long long>((n_list + 7) / 8, (long long)ctx->sm_count * 2);
    k_healpix<KID, false, true><<<max(blocks, 1), 256, smem, ctx->stream>>>(P, g, calc_mean, nullptr, nullptr, nullptr, list,
                                                                            nullptr, false, nullptr, nullptr,
                                                                            ctx->d_counters, n_list, recs, skip);
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

}  // namespace

int s2g_launch_healpix(s2g_ctx* ctx, const s2g_particles& P, long long nside, int kernel, int calc_mean,
                       const unsigned char* take, double* amap, double* wmap)
{
    if (P.n <= 0) return S2G_OK;
    S2G_TRY(s2g_stage_wait(ctx, P.n));   // inputs may still be on their way (overlapped staging, s2g_api.cu)
    switch (kernel) {
    case S2G_KERNEL_CUBIC: return launch_healpix_k<S2G_KERNEL_CUBIC>(ctx, P, nside, calc_mean, take, amap, wmap);
    case S2G_KERNEL_QUINTIC: return launch_healpix_k<S2G_KERNEL_QUINTIC>(ctx, P, nside, calc_mean, take, amap, wmap);
    case S2G_KERNEL_WENDLAND_C2: return launch_healpix_k<S2G_KERNEL_WENDLAND_C2>(ctx, P, nside, calc_mean, take, amap, wmap);
    case S2G_KERNEL_WENDLAND_C4: return launch_healpix_k<S2G_KERNEL_WENDLAND_C4>(ctx, P, nside, calc_mean, take, amap, wmap);
    case S2G_KERNEL_WENDLAND_C6: return launch_healpix_k<S2G_KERNEL_WENDLAND_C6>(ctx, P, nside, calc_mean, take, amap, wmap);
    case S2G_KERNEL_WENDLAND_C8: return launch_healpix_k<S2G_KERNEL_WENDLAND_C8>(ctx, P, nside, calc_mean, take, amap, wmap);
    }
    s2g_set_error("unknown kernel id %d", kernel);
    return S2G_EINVAL;
}

int s2g_hp_launch_records(s2g_ctx* ctx, const s2g_particles& P, long long nside, int kernel, int calc_mean,
                          const unsigned* list, long long n_list, HRec* recs, unsigned char* skip, int coop)
{
    switch (kernel) {
#define HPR_CASE(K) case K: return launch_records_k<K>(ctx, P, nside, calc_mean, list, n_list, recs, skip, coop);
        HPR_CASE(S2G_KERNEL_CUBIC)
        HPR_CASE(S2G_KERNEL_QUINTIC)
        HPR_CASE(S2G_KERNEL_WENDLAND_C2)
        HPR_CASE(S2G_KERNEL_WENDLAND_C4)
        HPR_CASE(S2G_KERNEL_WENDLAND_C6)
        HPR_CASE(S2G_KERNEL_WENDLAND_C8)
#undef HPR_CASE
    }
    s2g_set_error("unknown kernel id %d", kernel);
    return S2G_EINVAL;
}

// ------------------------------------------------------------------------------------------------
// filter_sort_particles (src/healpix_interpolation/filter_particles.jl:17-54) on the device.
// The reference selects `sorted[sel]` with sorted = reverse(sortperm(Δx)) and sel the shell mask IN ORIGINAL ORDER:
// the particle deposited for every i with sel[i] is the one of rank i in the far-to-near order (quirk Q5).  When
// every particle is in the shell this is simply "all of them" (no sort needed); otherwise a stable radix sort of
// the radii gives the permutation and take[sorted[i]] = 1 for each selected i.  The deposit order is irrelevant.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_hp_radii(s2g_particles P, double r0, double r1,
                                                  unsigned long long* __restrict__ keys,
                                                  unsigned char* __restrict__ sel, unsigned long long* nsel)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    bool s = false;
    if (p < P.n) {
        const double x = ld_pos(P, p, 0), y = ld_pos(P, p, 1), z = ld_pos(P, p, 2);
        // Δx = @. √(Pos[1,:]^2 + Pos[2,:]^2 + Pos[3,:]^2)
        const double dx = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
        s = (r0 <= dx) && (dx <= r1);
        keys[p] = (unsigned long long)__double_as_longlong(dx);  // dx >= 0: the bit pattern orders like the value
        sel[p] = s ? 1 : 0;
    }
    const unsigned b = __ballot_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(nsel, (unsigned long long)__popc(b));
}

__global__ void __launch_bounds__(256) k_hp_iota(unsigned* __restrict__ idx, long long n)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (unsigned)i;
}

__global__ void __launch_bounds__(256) k_hp_take(const unsigned* __restrict__ asc, const unsigned char* __restrict__ sel,
                                                 long long n, unsigned char* __restrict__ take)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sel[i]) take[asc[n - 1 - i]] = 1;  // sorted = reverse(sortperm(Δx)); sorted[i] is selected
}
}  // namespace

// radii (as sortable keys) and shell mask of the particles of P; *n_selected = particles inside the shell.
// Synchronises the stream.
int s2g_hp_radii(s2g_ctx* ctx, const s2g_particles& P, double r0, double r1, unsigned long long* keys_dev,
                 unsigned char* sel_dev, long long* n_selected)
{
    *n_selected = 0;
    if (P.n <= 0) return S2G_OK;
    S2G_TRY(s2g_stage_wait(ctx, P.n));   // inputs may still be on their way (overlapped staging, s2g_api.cu)
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_PAIRS, 0, sizeof(unsigned long long), ctx->stream));
    const int blocks = (int)((P.n + 255) / 256);
    const int ph = s2g_phase_begin(ctx, PH_PREP);
    k_hp_radii<<<blocks, 256, 0, ctx->stream>>>(P, r0, r1, keys_dev, sel_dev, ctx->d_counters + CNT_PAIRS);
    S2G_CUDA(cudaGetLastError());
    unsigned long long h_nsel = 0;
    S2G_CUDA(cudaMemcpyAsync(&h_nsel, ctx->d_counters + CNT_PAIRS, sizeof(h_nsel), cudaMemcpyDeviceToHost, ctx->stream));
    s2g_phase_end(ctx, ph);
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 1;
    *n_selected = (long long)h_nsel;
    return S2G_OK;
}

// take[sorted[i]] = 1 for every i with sel[i], sorted = reverse(sortperm(Δx)) (stable ascending radix sort of the
// radii, read backwards) — the reference's `sorted[sel]` (filter_particles.jl:33-41).  keys/sel/take: n elements on
// the context's device.
int s2g_hp_take_mask(s2g_ctx* ctx, const unsigned long long* keys_dev, const unsigned char* sel_dev, long long n,
                     unsigned char* take_dev)
{
    if (n <= 0) return S2G_OK;
    S2G_CHECK(n < 2147483647LL, S2G_EINVAL, "healpix far-to-near selection: n = %lld exceeds 2^31-1", n);
    void *d_keys2, *d_idx, *d_idx2, *d_tmp;
    S2G_TRY(s2g_scratch(ctx, "hp_keys2", sizeof(unsigned long long) * n, &d_keys2));
    S2G_TRY(s2g_scratch(ctx, "hp_idx", sizeof(unsigned) * n, &d_idx));
    S2G_TRY(s2g_scratch(ctx, "hp_idx2", sizeof(unsigned) * n, &d_idx2));
    const int blocks = (int)((n + 255) / 256);
    const int ph = s2g_phase_begin(ctx, PH_SORT);
    k_hp_iota<<<blocks, 256, 0, ctx->stream>>>((unsigned*)d_idx, n);
    S2G_CUDA(cudaGetLastError());
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, keys_dev, (unsigned long long*)d_keys2, (const unsigned*)d_idx,
                                    (unsigned*)d_idx2, (int)n, 0, 64, ctx->stream);
    S2G_TRY(s2g_scratch(ctx, "g_sort_tmp", sb + 16, &d_tmp));
    S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, keys_dev, (unsigned long long*)d_keys2, (const unsigned*)d_idx,
                                             (unsigned*)d_idx2, (int)n, 0, 64, ctx->stream));
    S2G_CUDA(cudaMemsetAsync(take_dev, 0, (size_t)n, ctx->stream));
    k_hp_take<<<blocks, 256, 0, ctx->stream>>>((const unsigned*)d_idx2, sel_dev, n, take_dev);
    S2G_CUDA(cudaGetLastError());
    s2g_phase_end(ctx, ph);
    ctx->launches += 9;
    return S2G_OK;
}

int s2g_launch_healpix_filtered(s2g_ctx* ctx, const s2g_particles& P, double r0, double r1, long long nside,
                                int kernel, int calc_mean, double* amap, double* wmap, long long* n_selected)
{
    if (P.n <= 0) { if (n_selected) *n_selected = 0; return S2G_OK; }
    const long long n = P.n;
    void *d_keys, *d_sel, *d_take;
    S2G_TRY(s2g_scratch(ctx, "hp_keys", sizeof(unsigned long long) * n, &d_keys));
    S2G_TRY(s2g_scratch(ctx, "hp_sel", (size_t)n, &d_sel));
    long long nsel = 0;
    S2G_TRY(s2g_hp_radii(ctx, P, r0, r1, (unsigned long long*)d_keys, (unsigned char*)d_sel, &nsel));
    if (n_selected) *n_selected = nsel;
    const unsigned char* take = nullptr;
    if (nsel != n) {
        S2G_TRY(s2g_scratch(ctx, "hp_take", (size_t)n, &d_take));
        S2G_TRY(s2g_hp_take_mask(ctx, (const unsigned long long*)d_keys, (const unsigned char*)d_sel, n,
                                 (unsigned char*)d_take));
        take = (const unsigned char*)d_take;
    }
    return s2g_launch_healpix(ctx, P, nside, kernel, calc_mean, take, amap, wmap);
}

int s2g_launch_healpix_pixels(s2g_ctx* ctx, const double pos[3], double radius, long long nside, long long* out,
                              long long cap, long long* count)
{
    const HpGeom g = make_hp(nside);
    k_healpix_pixels<<<1, 1, 0, ctx->stream>>>(pos[0], pos[1], pos[2], radius, g, out, cap, count);
    S2G_CUDA(cudaGetLastError());
    return S2G_OK;
}
