// s2g_cic2d.cu — 2D Smac deposit, scatter strategy (warp per particle, red.global.add.f64), footprints,
// reduce_image_2D epilogue and centre/filter kernels.
//
// Replaces: cic_mapping_2D (src/cic_interpolation/cic_2D.jl:103-244), calculate_weights (:11-72),
//           reduce_image_2D (src/cic_interpolation/reduce_image.jl:8-31),
//           center_particles / filter_particles_in_image (src/cic_interpolation/filter_shift.jl:6-67).
#include <cub/cub.cuh>
#include <cstdlib>

#include "s2g_cic2d.cuh"

// ------------------------------------------------------------------------------------------------
// one warp deposits one particle: pass A (weight sums) + pass B (normalised scatter)
// lanes are laid out W wide along j (the contiguous image axis, indices.jl:6-8) and 32/W rows deep,
// W = next power of two >= footprint width (capped at 32) so that small footprints still fill the warp.
// ------------------------------------------------------------------------------------------------
// (row, column) of the flattened footprint index e (j fastest) without an integer division: float reciprocal + one
// correction step (exact for every e, nj < 2^20)
__device__ __forceinline__ void unflatten(int e, int nj, float inv_nj, int& ir, int& jc)
{
    ir = __float2int_rd(((float)e + 0.5f) * inv_nj);
    jc = e - ir * nj;
    if (jc < 0) { --ir; jc += nj; }
    else if (jc >= nj) { ++ir; jc -= nj; }
}

// The records of the (up to) 32 particles a warp has taken, one slot per lane, in shared memory: the lane that built a
// record posts it, every lane reads the slot it is told to (a broadcast when the warp works on one particle).
struct RecBoard {
    double d[7][32];
    int i[6][32];
};
__device__ __forceinline__ void board_post(RecBoard& b, int lane, const Rec2& m, long long p)
{
    b.d[0][lane] = m.x; b.d[1][lane] = m.y; b.d[2][lane] = m.h; b.d[3][lane] = m.hinv;
    b.d[4][lane] = m.area; b.d[5][lane] = m.dz; b.d[6][lane] = m.w;
    b.i[0][lane] = m.iMin; b.i[1][lane] = m.iMax; b.i[2][lane] = m.jMin; b.i[3][lane] = m.jMax;
    b.i[4][lane] = m.all_zero ? 1 : 0;
    b.i[5][lane] = (int)p;   // particle indices are below 2^31 (checked at the boundary)
}
__device__ __forceinline__ Rec2 board_read(const RecBoard& b, int src, long long& p)
{
    Rec2 r;
    r.x = b.d[0][src]; r.y = b.d[1][src]; r.h = b.d[2][src]; r.hinv = b.d[3][src];
    r.area = b.d[4][src]; r.dz = b.d[5][src]; r.w = b.d[6][src];
    r.iMin = b.i[0][src]; r.iMax = b.i[1][src]; r.jMin = b.i[2][src]; r.jMax = b.i[3][src];
    r.all_zero = b.i[4][src] != 0;
    p = (long long)b.i[5][src];
    return r;
}

constexpr int SC2_CAP = 384;   // pixels of a footprint whose pass-A weights a warp keeps in shared memory (3 KB per warp)

// Footprints of at most SC2_CAP pixels: the warp walks the FLATTENED footprint (all 32 lanes busy whatever the width)
// and pass A leaves wk dx dy of every pixel in shared memory (a lane reads back only what it wrote: no barrier), so
// that pass B is one multiply and the reds — no second square root, no second kernel evaluation.  The rare branches
// (no pixel centre covered, Inf/NaN norm) return false and take the general two-pass code below.
template <int KID>
__device__ __forceinline__ bool warp_deposit_2d_cached(const Rec2& r, const s2g_particles& P, const s2g_geom& G, long long p,
                                                       int lane, double* __restrict__ image, unsigned long long& touched,
                                                       double* __restrict__ s_g)
{
    const int ni = r.iMax - r.iMin + 1, nj = r.jMax - r.jMin + 1, npx = ni * nj;
    const double dx_lo = overlap_1d(r.x, r.h, r.iMin), dx_hi = overlap_1d(r.x, r.h, r.iMax);
    const double dy_lo = overlap_1d(r.y, r.h, r.jMin), dy_hi = overlap_1d(r.y, r.h, r.jMax);
    const float inv_nj = 1.0f / (float)nj;
    double sw = 0.0;
    int cnt = 0;
    for (int e = lane; e < npx; e += 32) {
        int ir, jc;
        unflatten(e, nj, inv_nj, ir, jc);
        const int i = r.iMin + ir, j = r.jMin + jc;
        const double xd = center_dist(r.x, (double)i), yd = center_dist(r.y, (double)j);
        const double u = u_of(__dmul_rn(xd, xd), __dmul_rn(yd, yd), r.hinv);
        double g = 0.0;
        if (u <= 1.0) {
            const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
            const double dy = (j == r.jMin) ? dy_lo : ((j == r.jMax) ? dy_hi : 1.0);
            const double wk = kernel_shape<KID>(u);
            const double dxy = dx * dy;
            sw = fma(wk, dxy, sw);
            g = wk * dxy;
            ++cnt;
        }
        s_g[e] = g;
    }
    sw = warp_sum(sw);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (sw == 0.0) return false;
    const double n_distr = (double)cnt, wpp = n_distr / sw;
    const double kernel_norm = r.area / n_distr;
    const double area_norm = kernel_norm * wpp * r.w * r.dz;
    if (!isfinite(area_norm)) return false;
    const long long npl = G.npix * G.npix;
    double* __restrict__ wplane = image + npl * G.n_images;
    const int nim = G.n_images;
    double q0 = 0.0;
    if (!r.all_zero && nim == 1) q0 = ld_in(P.binq, p, P.in_dtype);
    for (int e = lane; e < npx; e += 32) {
        const double pw = s_g[e] * area_norm;
        if (pw != 0.0) {
            int ir, jc;
            unflatten(e, nj, inv_nj, ir, jc);
            const long long idx = (long long)(r.iMin + ir) * G.npix + (r.jMin + jc);
            red_add(wplane + idx, pw);
            if (r.all_zero) {
                if (!isfinite(pw)) red_add(image + idx, 0.0 * pw);
            } else if (nim == 1) {
                red_add(image + idx, q0 * pw);
            } else {
                for (int q = 0; q < nim; ++q)
                    red_add(image + npl * q + idx, ld_in(P.binq, (long long)nim * p + q, P.in_dtype) * pw);
            }
            ++touched;
        }
    }
    return true;
}

template <int KID>
__device__ __forceinline__ void warp_deposit_2d(const Rec2& r, const s2g_particles& P, const s2g_geom& G, long long p,
                                                int lane, double* __restrict__ image, unsigned long long& touched,
                                                unsigned long long& fallback)
{
    const int ni = r.iMax - r.iMin + 1, nj = r.jMax - r.jMin + 1;
    const int lw = nj >= 32 ? 5 : (nj <= 1 ? 0 : 32 - __clz(nj - 1));
    const int W = 1 << lw, R = 32 >> lw;
    const int c0 = lane & (W - 1), r0 = lane >> lw;

    const double dx_lo = overlap_1d(r.x, r.h, r.iMin), dx_hi = overlap_1d(r.x, r.h, r.iMax);
    const double dy_lo = overlap_1d(r.y, r.h, r.jMin), dy_hi = overlap_1d(r.y, r.h, r.jMax);

    // ---- pass A: distr_weight = Σ wk·dA over pixels whose centre is inside the kernel, n_distr_pix
    double sw = 0.0;
    int cnt = 0;
    for (int jc = c0; jc < nj; jc += W) {
        const int j = r.jMin + jc;
        const double yd = center_dist(r.y, (double)j);
        const double yd2 = __dmul_rn(yd, yd);
        const double dy = (j == r.jMin) ? dy_lo : ((j == r.jMax) ? dy_hi : 1.0);
        for (int ir = r0; ir < ni; ir += R) {
            const int i = r.iMin + ir;
            const double xd = center_dist(r.x, (double)i);
            const double u = u_of(__dmul_rn(xd, xd), yd2, r.hinv);
            if (u <= 1.0) {
                const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
                sw = fma(kernel_shape<KID>(u), dx * dy, sw);
                ++cnt;
            }
        }
    }
    sw = warp_sum(sw);
    cnt = __reduce_add_sync(0xffffffffu, cnt);

    // ---- normalisation (cic_2D.jl:51-69, :187-188)
    bool fb = false;
    double n_distr, wpp;
    if (sw == 0.0) {
        fb = true;
        double da = 0.0;
        for (int jc = c0; jc < nj; jc += W) {
            const int j = r.jMin + jc;
            const double dy = (j == r.jMin) ? dy_lo : ((j == r.jMax) ? dy_hi : 1.0);
            for (int ir = r0; ir < ni; ir += R) {
                const int i = r.iMin + ir;
                const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
                da += dx * dy;
            }
        }
        da = warp_sum(da);
        n_distr = (double)ni * (double)nj;
        wpp = (da != 0.0) ? n_distr / da : 1.0;
        if (lane == 0) ++fallback;
    } else {
        n_distr = (double)cnt;
        wpp = n_distr / sw;
    }
    const double kernel_norm = r.area / n_distr;
    const double area_norm = kernel_norm * wpp * r.w * r.dz;
    const bool poison = !isfinite(area_norm);

    // ---- pass B: pix_weight = wk·A·area_norm; update_image! (cic_shared.jl:111-121)
    const long long npl = G.npix * G.npix;
    double* __restrict__ wplane = image + npl * G.n_images;
    const int nim = G.n_images;
    double q0 = 0.0;
    if (!r.all_zero && nim == 1) q0 = ld_in(P.binq, p, P.in_dtype);
    for (int jc = c0; jc < nj; jc += W) {
        const int j = r.jMin + jc;
        const double yd = center_dist(r.y, (double)j);
        const double yd2 = __dmul_rn(yd, yd);
        const double dy = (j == r.jMin) ? dy_lo : ((j == r.jMax) ? dy_hi : 1.0);
        for (int ir = r0; ir < ni; ir += R) {
            const int i = r.iMin + ir;
            const double xd = center_dist(r.x, (double)i);
            double wk;
            if (fb)
                wk = 1.0;
            else {
                const double u = u_of(__dmul_rn(xd, xd), yd2, r.hinv);
                if (!(u <= 1.0)) {
                    // wk = 0: pix_weight = 0*A*area_norm is 0 and skipped — unless area_norm is Inf/NaN (rho = 0,
                    // NaN weights): then the reference's `!iszero(pix_weight)` is true for the whole bounding box
                    if (!poison) continue;
                    wk = 0.0;
                } else
                    wk = kernel_shape<KID>(u);
            }
            const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
            const double pw = wk * (dx * dy) * area_norm;
            if (pw != 0.0) {
                const long long idx = (long long)i * G.npix + j;
                red_add(wplane + idx, pw);
                if (r.all_zero) {
                    if (!isfinite(pw)) red_add(image + idx, 0.0 * pw);  // bin_q collapsed to 0.0 (cic_2D.jl:160-162)
                } else if (nim == 1) {
                    red_add(image + idx, q0 * pw);
                } else {
                    for (int q = 0; q < nim; ++q)
                        red_add(image + npl * q + idx, ld_in(P.binq, (long long)nim * p + q, P.in_dtype) * pw);
                }
                ++touched;
            }
        }
    }
}

// particle list: either all particles [0,n) (list == nullptr) or an explicit index list (scatter bin of AUTO)
template <int KID>
__global__ void __launch_bounds__(256) k_scatter2d(s2g_particles P, s2g_geom G, const int* __restrict__ list,
                                                   long long n_list, double* __restrict__ image,
                                                   unsigned long long* __restrict__ counters, int chunk)
{
    const int lane = threadIdx.x & 31;
    __shared__ double s_cache[8 * SC2_CAP];
    __shared__ RecBoard s_board[8];
    double* __restrict__ s_g = s_cache + (threadIdx.x >> 5) * SC2_CAP;
    RecBoard& board = s_board[threadIdx.x >> 5];
    unsigned long long touched = 0, fallback = 0, mapped = 0, fpx = 0;
    // a warp takes `chunk` (<= 32) particles at a time: lane l loads particle base + l and builds its record (one
    // coalesced load per input array, the three divisions of the record once per particle and 32 particles at a time,
    // one memory latency per chunk instead of one per particle); the warp then deposits them one after the other, the
    // record handed round by shuffles.
    for (;;) {
        long long base = 0;
        if (lane == 0) base = (long long)atomicAdd(&counters[CNT_WORK], (unsigned long long)chunk);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_list) break;
        const long long tl = base + lane;
        Rec2 mine = {};
        long long pm = 0;
        bool ok = false;
        if (lane < chunk && tl < n_list) {
            pm = list ? (long long)list[tl] : tl;
            ok = make_rec2(P, G, pm, mine);
        }
        __syncwarp();   // the previous chunk's records have been read by every lane
        if (ok) {
            ++mapped;
            fpx += (unsigned long long)(mine.iMax - mine.iMin + 1) * (unsigned long long)(mine.jMax - mine.jMin + 1);
            board_post(board, lane, mine, pm);
        }
        __syncwarp();   // posts before reads
        unsigned todo = __ballot_sync(0xffffffffu, ok);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            long long p;
            const Rec2 r = board_read(board, src, p);
            const unsigned long long npx =
                (unsigned long long)(r.iMax - r.iMin + 1) * (unsigned long long)(r.jMax - r.jMin + 1);
            if (npx <= (unsigned long long)SC2_CAP && warp_deposit_2d_cached<KID>(r, P, G, p, lane, image, touched, s_g))
                continue;
            warp_deposit_2d<KID>(r, P, G, p, lane, image, touched, fallback);
        }
    }
    mapped = (unsigned long long)warp_sum_ll((long long)mapped);
    fpx = (unsigned long long)warp_sum_ll((long long)fpx);
    touched = (unsigned long long)warp_sum_ll((long long)touched);
    if (lane == 0) {
        if (touched) atomicAdd(&counters[CNT_TOUCHED], touched);
        if (fallback) atomicAdd(&counters[CNT_FALLBACK], fallback);
        if (mapped) atomicAdd(&counters[CNT_MAPPED], mapped);
        if (mapped) atomicAdd(&counters[CNT_SCATTER], mapped);
        if (fpx) atomicAdd(&counters[CNT_FOOTPRINT], fpx);
    }
}

// ------------------------------------------------------------------------------------------------
// sub-warp particles: footprints of at most SUB*8 pixels (a few pixels across) leave most of a warp idle and pay the
// per-particle set-up (record, four overlaps, three divisions of the normalisation) once per WARP.  Here SUB lanes
// share a particle (32/SUB particles per warp, each with its own footprint); a lane walks the flattened footprint
// t = sub, sub+SUB, ... (j fastest, so the lanes of a group write neighbouring pixels); the pass-A sums are reduced
// inside the group by xor-shuffles below SUB.  Same arithmetic per pixel as warp_deposit_2d.
// ------------------------------------------------------------------------------------------------
template <int KID, int SUB>
__global__ void __launch_bounds__(256) k_scatter2d_sub(s2g_particles P, s2g_geom G, const int* __restrict__ list,
                                                       long long n_list, double* __restrict__ image,
                                                       unsigned long long* __restrict__ counters, int chunk)
{
    constexpr int GROUPS = 32 / SUB;
    const int lane = threadIdx.x & 31, grp = lane / SUB, sub = lane % SUB;
    __shared__ double s_cache[8 * 256];
    __shared__ RecBoard s_board[8];
    double* __restrict__ s_g = s_cache + (threadIdx.x >> 5) * 256;
    RecBoard& board = s_board[threadIdx.x >> 5];
    unsigned long long touched = 0, fallback = 0, mapped = 0, fpx = 0;
    const long long npl = G.npix * G.npix;
    double* __restrict__ wplane = image + npl * G.n_images;
    const int nim = G.n_images;
    for (;;) {
        long long base = 0;
        if (lane == 0) base = (long long)atomicAdd(&counters[CNT_WORK], (unsigned long long)chunk);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_list) break;
        // lane l loads particle base + l and builds its record (k_scatter2d above); group g then takes the particles
        // g, g + GROUPS, ... of the chunk
        const long long tl = base + lane;
        Rec2 mine = {};
        long long pm = 0;
        bool ok = false;
        if (lane < chunk && tl < n_list) {
            pm = list ? (long long)list[tl] : tl;
            ok = make_rec2(P, G, pm, mine);
        }
        __syncwarp();   // the previous chunk's records have been read by every lane
        if (ok) board_post(board, lane, mine, pm);
        __syncwarp();   // posts before reads
        const unsigned okmask = __ballot_sync(0xffffffffu, ok);
        for (int k0 = 0; k0 < chunk; k0 += GROUPS) {
            if (((okmask >> k0) & ((1u << GROUPS) - 1u)) == 0u) continue;   // uniform: nothing to do in this round
            const int src = min(k0 + grp, 31);
            const bool act = (k0 + grp < chunk) && ((okmask >> src) & 1u);
            long long p = 0;
            Rec2 r = {};
            if (act) r = board_read(board, src, p);
            int ni = 0, nj = 1, npx = 0;
            double dx_lo = 0, dx_hi = 0, dy_lo = 0, dy_hi = 0;
            if (act) {
                ni = r.iMax - r.iMin + 1; nj = r.jMax - r.jMin + 1; npx = ni * nj;
                dx_lo = overlap_1d(r.x, r.h, r.iMin); dx_hi = overlap_1d(r.x, r.h, r.iMax);
                dy_lo = overlap_1d(r.y, r.h, r.jMin); dy_hi = overlap_1d(r.y, r.h, r.jMax);
                if (sub == 0) { ++mapped; fpx += (unsigned long long)npx; }
            }
            // ---- pass A.  A footprint of at most SUB x MAXIT = 64 pixels (the tiny class) leaves wk dx dy of the lane's
            // pixels in shared memory (slot m*32 + lane: a lane reads back only what it wrote), so that pass B neither
            // takes the square root nor evaluates the kernel a second time; (row, column) of the flattened index come
            // from a float reciprocal instead of an integer division.
            constexpr int MAXIT = 64 / SUB;
            const bool small = npx <= SUB * MAXIT;
            const float inv_nj = 1.0f / (float)nj;
            double sw = 0.0, da = 0.0;
            int cnt = 0;
            if (small) {
                for (int e = sub, m = 0; e < npx; e += SUB, ++m) {
                    int ir, jc;
                    unflatten(e, nj, inv_nj, ir, jc);
                    const int i = r.iMin + ir, j = r.jMin + jc;
                    const double xd = center_dist(r.x, (double)i), yd = center_dist(r.y, (double)j);
                    const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
                    const double dy = (j == r.jMin) ? dy_lo : ((j == r.jMax) ? dy_hi : 1.0);
                    const double u = u_of(__dmul_rn(xd, xd), __dmul_rn(yd, yd), r.hinv);
                    const double dxy = dx * dy;
                    da += dxy;
                    double g = 0.0;
                    if (u <= 1.0) {
                        const double wk = kernel_shape<KID>(u);
                        sw = fma(wk, dxy, sw);
                        g = wk * dxy;
                        ++cnt;
                    }
                    s_g[m * 32 + lane] = g;
                }
            } else {
                for (int e = sub; e < npx; e += SUB) {
                    const int ir = e / nj, jc = e - ir * nj;
                    const int i = r.iMin + ir, j = r.jMin + jc;
                    const double xd = center_dist(r.x, (double)i), yd = center_dist(r.y, (double)j);
                    const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
                    const double dy = (j == r.jMin) ? dy_lo : ((j == r.jMax) ? dy_hi : 1.0);
                    const double u = u_of(__dmul_rn(xd, xd), __dmul_rn(yd, yd), r.hinv);
                    da += dx * dy;
                    if (u <= 1.0) {
                        sw = fma(kernel_shape<KID>(u), dx * dy, sw);
                        ++cnt;
                    }
                }
            }
#pragma unroll
            for (int o = SUB / 2; o > 0; o >>= 1) {
                sw += __shfl_xor_sync(0xffffffffu, sw, o);
                da += __shfl_xor_sync(0xffffffffu, da, o);
                cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            }
            if (!act) continue;   // (after the shuffles: every lane of the warp takes part in them)
            // ---- normalisation (cic_2D.jl:51-69, :187-188)
            bool fb = false;
            double n_distr, wpp;
            if (sw == 0.0) {
                fb = true;
                n_distr = (double)ni * (double)nj;
                wpp = (da != 0.0) ? n_distr / da : 1.0;
                if (sub == 0) ++fallback;
            } else {
                n_distr = (double)cnt;
                wpp = n_distr / sw;
            }
            const double kernel_norm = r.area / n_distr;
            const double area_norm = kernel_norm * wpp * r.w * r.dz;
            const bool poison = !isfinite(area_norm);
            double q0 = 0.0;
            if (!r.all_zero && nim == 1) q0 = ld_in(P.binq, p, P.in_dtype);
            // ---- pass B
            if (small && !fb && !poison) {
                for (int e = sub, m = 0; e < npx; e += SUB, ++m) {
                    const double pw = s_g[m * 32 + lane] * area_norm;
                    if (pw != 0.0) {
                        int ir, jc;
                        unflatten(e, nj, inv_nj, ir, jc);
                        const long long idx = (long long)(r.iMin + ir) * G.npix + (r.jMin + jc);
                        red_add(wplane + idx, pw);
                        if (r.all_zero) {
                            if (!isfinite(pw)) red_add(image + idx, 0.0 * pw);
                        } else if (nim == 1) {
                            red_add(image + idx, q0 * pw);
                        } else {
                            for (int q = 0; q < nim; ++q)
                                red_add(image + npl * q + idx, ld_in(P.binq, (long long)nim * p + q, P.in_dtype) * pw);
                        }
                        ++touched;
                    }
                }
                continue;
            }
            for (int e = sub; e < npx; e += SUB) {
                const int ir = e / nj, jc = e - ir * nj;
                const int i = r.iMin + ir, j = r.jMin + jc;
                const double xd = center_dist(r.x, (double)i), yd = center_dist(r.y, (double)j);
                double wk;
                if (fb)
                    wk = 1.0;
                else {
                    const double u = u_of(__dmul_rn(xd, xd), __dmul_rn(yd, yd), r.hinv);
                    if (!(u <= 1.0)) {
                        if (!poison) continue;
                        wk = 0.0;
                    } else
                        wk = kernel_shape<KID>(u);
                }
                const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
                const double dy = (j == r.jMin) ? dy_lo : ((j == r.jMax) ? dy_hi : 1.0);
                const double pw = wk * (dx * dy) * area_norm;
                if (pw != 0.0) {
                    const long long idx = (long long)i * G.npix + j;
                    red_add(wplane + idx, pw);
                    if (r.all_zero) {
                        if (!isfinite(pw)) red_add(image + idx, 0.0 * pw);
                    } else if (nim == 1) {
                        red_add(image + idx, q0 * pw);
                    } else {
                        for (int q = 0; q < nim; ++q)
                            red_add(image + npl * q + idx, ld_in(P.binq, (long long)nim * p + q, P.in_dtype) * pw);
                    }
                    ++touched;
                }
            }
        }
    }
    touched = (unsigned long long)warp_sum_ll((long long)touched);
    fallback = (unsigned long long)warp_sum_ll((long long)fallback);
    mapped = (unsigned long long)warp_sum_ll((long long)mapped);
    fpx = (unsigned long long)warp_sum_ll((long long)fpx);
    if (lane == 0) {
        if (touched) atomicAdd(&counters[CNT_TOUCHED], touched);
        if (fallback) atomicAdd(&counters[CNT_FALLBACK], fallback);
        if (mapped) atomicAdd(&counters[CNT_MAPPED], mapped);
        if (mapped) atomicAdd(&counters[CNT_SCATTER], mapped);
        if (fpx) atomicAdd(&counters[CNT_FOOTPRINT], fpx);
    }
}

// ---- deposit order of a scatter list: by the 64x64-pixel block of the particle centre.  In input order every red row of
// a small footprint misses L2 once the image is larger than L2 (read + write of a DRAM sector per row); in block order
// the ~10^4 particles in flight update one compact region that stays in L2.  One radix sort of (key, index) pairs per
// list; S2G_2D_ORDER=0 switches it off.
__global__ void __launch_bounds__(256) k_order2d_keys(s2g_particles P, s2g_geom G, const int* __restrict__ list,
                                                      long long n_list, int nb, unsigned* __restrict__ keys,
                                                      unsigned* __restrict__ idx)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_list) return;
    const long long p = list ? (long long)list[t] : t;
    unsigned key = 0;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const double x = fma(ld_pos(P, p, d), G.len2pix, G.half_n);
        int b = (int)floor(x * (1.0 / 64.0));
        b = min(max(b, 0), nb - 1);
        key = key * (unsigned)nb + (unsigned)b;
    }
    keys[t] = key;
    idx[t] = (unsigned)p;
}

// particles a warp takes per visit of the work counter: 32 when the list keeps every warp busy for several visits,
// fewer for short lists (so that they still spread over the SMs)
static int scatter_chunk(long long n_list, int blocks)
{
    const long long warps = (long long)blocks * 8;
    if (n_list >= 4 * 32 * warps) return 32;
    if (n_list >= 4 * 8 * warps) return 8;
    return 4;
}

static int order_list_2d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, const int** list, long long n_list)
{
    const char* e_ord = getenv("S2G_2D_ORDER");
    const long long image_bytes = G.npix * G.npix * 8LL * (G.n_images + 1);
    if ((e_ord && atoi(e_ord) == 0) || n_list < 65536 || n_list >= (1LL << 31) || image_bytes < (48LL << 20)) return S2G_OK;
    void *d_k, *d_k2, *d_i, *d_i2, *d_tmp;
    S2G_TRY(s2g_scratch(ctx, "o2_keys", sizeof(unsigned) * n_list, &d_k));
    S2G_TRY(s2g_scratch(ctx, "o2_keys2", sizeof(unsigned) * n_list, &d_k2));
    S2G_TRY(s2g_scratch(ctx, "o2_idx", sizeof(unsigned) * n_list, &d_i));
    S2G_TRY(s2g_scratch(ctx, "o2_idx2", sizeof(unsigned) * n_list, &d_i2));
    const int nb = (int)((G.npix + 63) / 64);
    k_order2d_keys<<<(int)((n_list + 255) / 256), 256, 0, ctx->stream>>>(P, G, *list, n_list, nb, (unsigned*)d_k,
                                                                           (unsigned*)d_i);
    S2G_CUDA(cudaGetLastError());
    int bits = 1;
    while ((1LL << bits) < (long long)nb * nb) ++bits;
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, (const unsigned*)d_k, (unsigned*)d_k2, (const unsigned*)d_i,
                                    (unsigned*)d_i2, (int)n_list, 0, bits, ctx->stream);
    S2G_TRY(s2g_scratch(ctx, "o2_sort_tmp", sb + 16, &d_tmp));
    S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, (const unsigned*)d_k, (unsigned*)d_k2, (const unsigned*)d_i,
                                             (unsigned*)d_i2, (int)n_list, 0, bits, ctx->stream));
    ctx->launches += 4;
    *list = (const int*)d_i2;   // particle indices are below 2^31 (s2g_particles.n is checked at the boundary)
    return S2G_OK;
}

template <int KID>
static int launch_scatter2d_sub_k(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, const int* list,
                                  long long n_list, double* image)
{
    if (n_list <= 0) return S2G_OK;
    S2G_TRY(order_list_2d(ctx, P, G, &list, n_list));
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    const long long warps_needed = std::min<long long>((n_list + 15) / 16, (long long)ctx->sm_count * 8 * 8);
    int blocks = (int)std::max<long long>(1, (warps_needed + 7) / 8);
    blocks = min(blocks, ctx->sm_count * 8);
    k_scatter2d_sub<KID, 8><<<blocks, 256, 0, ctx->stream>>>(P, G, list, n_list, image, ctx->d_counters,
                                                             scatter_chunk(n_list, blocks));
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

// sub-warp launch for a list of particles whose footprints hold at most S2G_TINY_MAX_PIXELS (64) pixels
int s2g_launch_scatter_2d_tiny(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, const int* list,
                               long long n_list, double* image)
{
    switch (kernel) {
    case S2G_KERNEL_CUBIC: return launch_scatter2d_sub_k<S2G_KERNEL_CUBIC>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_QUINTIC: return launch_scatter2d_sub_k<S2G_KERNEL_QUINTIC>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C2: return launch_scatter2d_sub_k<S2G_KERNEL_WENDLAND_C2>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C4: return launch_scatter2d_sub_k<S2G_KERNEL_WENDLAND_C4>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C6: return launch_scatter2d_sub_k<S2G_KERNEL_WENDLAND_C6>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C8: return launch_scatter2d_sub_k<S2G_KERNEL_WENDLAND_C8>(ctx, P, G, list, n_list, image);
    }
    s2g_set_error("unknown kernel id %d", kernel);
    return S2G_EINVAL;
}

template <int KID>
static int launch_scatter2d_k(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, const int* list,
                              long long n_list, double* image)
{
    if (n_list <= 0) return S2G_OK;
    S2G_TRY(order_list_2d(ctx, P, G, &list, n_list));
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    const int warps_needed = (int)std::min<long long>((n_list + 3) / 4, (long long)ctx->sm_count * 8 * 8);
    int blocks = max(1, (warps_needed + 7) / 8);
    blocks = min(blocks, ctx->sm_count * 8);
    k_scatter2d<KID><<<blocks, 256, 0, ctx->stream>>>(P, G, list, n_list, image, ctx->d_counters,
                                                      scatter_chunk(n_list, blocks));
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

int s2g_launch_scatter_2d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, const int* list,
                          long long n_list, double* image)
{
    switch (kernel) {
    case S2G_KERNEL_CUBIC: return launch_scatter2d_k<S2G_KERNEL_CUBIC>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_QUINTIC: return launch_scatter2d_k<S2G_KERNEL_QUINTIC>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C2: return launch_scatter2d_k<S2G_KERNEL_WENDLAND_C2>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C4: return launch_scatter2d_k<S2G_KERNEL_WENDLAND_C4>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C6: return launch_scatter2d_k<S2G_KERNEL_WENDLAND_C6>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C8: return launch_scatter2d_k<S2G_KERNEL_WENDLAND_C8>(ctx, P, G, list, n_list, image);
    }
    s2g_set_error("unknown kernel id %d", kernel);
    return S2G_EINVAL;
}

// ------------------------------------------------------------------------------------------------
// footprints (bit-exact contract)
// ------------------------------------------------------------------------------------------------
__global__ void k_footprints2d(s2g_particles P, s2g_geom G, long long* __restrict__ out)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const double px = ld_pos(P, p, 0), py = ld_pos(P, p, 1);
    const double h = __dmul_rn(ld_in(P.hsml, p, P.in_dtype), G.len2pix);
    const double x = __dadd_rn(__dmul_rn(px, G.len2pix), G.half_n);
    const double y = __dadd_rn(__dmul_rn(py, G.len2pix), G.half_n);
    const int n1 = (int)G.npix - 1;
    out[4 * p + 0] = max(floor_to_int(__dadd_rn(x, -h)), 0);
    out[4 * p + 1] = min(floor_to_int(__dadd_rn(x, h)), n1);
    out[4 * p + 2] = max(floor_to_int(__dadd_rn(y, -h)), 0);
    out[4 * p + 3] = min(floor_to_int(__dadd_rn(y, h)), n1);
}

__global__ void k_footprints3d(s2g_particles P, s2g_geom G, long long* __restrict__ out)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const double h = __dmul_rn(ld_in(P.hsml, p, P.in_dtype), G.len2pix);
    const int n1 = (int)G.npix - 1;
    for (int d = 0; d < 3; ++d) {
        const double x = __dadd_rn(__dmul_rn(ld_pos(P, p, d), G.len2pix), G.half_n);
        out[6 * p + 2 * d + 0] = max(floor_to_int(__dadd_rn(x, -h)), 0);
        out[6 * p + 2 * d + 1] = min(floor_to_int(__dadd_rn(x, h)), n1);
    }
}

int s2g_launch_footprints(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int dims, long long* out)
{
    S2G_TRY(s2g_stage_wait(ctx, P.n));   // inputs may still be on their way (overlapped staging, s2g_api.cu)
    if (P.n <= 0) return S2G_OK;
    const int blocks = (int)((P.n + 255) / 256);
    if (dims == 2)
        k_footprints2d<<<blocks, 256, 0, ctx->stream>>>(P, G, out);
    else
        k_footprints3d<<<blocks, 256, 0, ctx->stream>>>(P, G, out);
    S2G_CUDA(cudaGetLastError());
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// reduce_image_2D: out[ix + nx*iy + nx*ny*q] = image[ix*nx + iy, q] (/ weight where reduce && weight > 0)
// = divide + transpose; 32x32 tiles through padded shared memory, both sides coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduce2d(const double* __restrict__ image, long long nx, long long ny,
                                                  int n_images, int reduce_image, double* __restrict__ out)
{
    __shared__ double tile[32][33];
    const long long npl = nx * ny;
    const long long bx = blockIdx.x * 32LL, by = blockIdx.y * 32LL;  // bx: iy block (fast axis of flat), by: ix block
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;         // 32 x 8
    const double* __restrict__ wpl = image + npl * n_images;
    for (int q = 0; q < n_images; ++q) {
        const double* __restrict__ src = image + npl * q;
        for (int r = ty; r < 32; r += 8) {
            const long long ix = by + r, iy = bx + tx;
            if (ix < nx && iy < ny) {
                const long long k = ix * nx + iy;
                double v = src[k];
                if (reduce_image) {
                    const double wv = wpl[k];
                    if (wv > 0.0) v = v / wv;
                }
                tile[r][tx] = v;
            }
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const long long iy = bx + r, ix = by + tx;
            if (ix < nx && iy < ny) out[ix + nx * iy + npl * q] = tile[tx][r];
        }
        __syncthreads();
    }
}

int s2g_launch_reduce_2d(s2g_ctx* ctx, const double* image, long long nx, long long ny, int n_images, int reduce_image,
                         double* out)
{
    dim3 grid((unsigned)((ny + 31) / 32), (unsigned)((nx + 31) / 32));
    const int ph = s2g_phase_begin(ctx, PH_EPILOGUE);
    k_reduce2d<<<grid, 256, 0, ctx->stream>>>(image, nx, ny, n_images, reduce_image, out);
    s2g_phase_end(ctx, ph);
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// center_particles + filter_particles_in_image as a standalone pass (the fused path does both on the fly)
// ------------------------------------------------------------------------------------------------
__global__ void k_center_filter(s2g_particles P, void* __restrict__ pos_out, uint8_t* __restrict__ mask)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const double x = ld_pos(P, p, 0), y = ld_pos(P, p, 1), z = ld_pos(P, p, 2);
    if (pos_out) {
        if (P.in_dtype == S2G_F64) {
            double* o = reinterpret_cast<double*>(pos_out) + 3 * p;
            o[0] = x; o[1] = y; o[2] = z;
        } else {
            float* o = reinterpret_cast<float*>(pos_out) + 3 * p;
            o[0] = (float)x; o[1] = (float)y; o[2] = (float)z;  // exact: values are Float32-representable
        }
    }
    if (mask) mask[p] = in_image(P, x, y, z) ? 1 : 0;
}

int s2g_launch_center_filter(s2g_ctx* ctx, const s2g_particles& P, void* pos_out, uint8_t* mask)
{
    S2G_TRY(s2g_stage_wait(ctx, P.n));   // inputs may still be on their way (overlapped staging, s2g_api.cu)
    if (P.n <= 0) return S2G_OK;
    const int blocks = (int)((P.n + 255) / 256);
    k_center_filter<<<blocks, 256, 0, ctx->stream>>>(P, pos_out, mask);
    S2G_CUDA(cudaGetLastError());
    return S2G_OK;
}
