// s2g_cic3d.cu — 3D Smac deposit (scatter strategy) and reduce_image_3D epilogue.
//
// Replaces: cic_mapping_3D (src/cic_interpolation/cic_3D.jl:110-209), calculate_weights (:13-78),
//           get_quantities_3D (:87-97), reduce_image_3D (src/cic_interpolation/reduce_image.jl:39-55).
#include <cub/cub.cuh>
#include <cstdlib>

#include "s2g_cic3d.cuh"

// lanes: W wide along k (contiguous axis, indices.jl:15-17), 32/W deep along j; i is walked by the whole warp in groups
// of four independent chains.  Coordinates in units of h: a cell centre is inside the kernel iff a²+b²+c² < 1
// (the reference's u = sqrt(dx²+dy²+dz²)·h⁻¹ <= 1 up to the last ulp at the rim, where w -> 0 anyway).
// S3_CAP: per-warp capacity of the shared-memory cell list.  Pass A stores the weight wk·dV and the packed box
// coordinates of every cell with a non-zero weight (warp-level compaction: ballot + prefix popcount, in the lanes'
// k-contiguous order), so that pass B neither re-evaluates the kernel (cic_3D.jl:172-188 recomputes nothing either:
// the reference keeps wk[] and V[] from calculate_weights) nor walks the empty corners of the bounding box: all 32
// lanes issue reds.  A particle with more non-zero cells than S3_CAP (h > ~6 cells) takes the two-pass path.
#ifndef S2G_3D_CAP
#define S2G_3D_CAP 1024
#endif
#ifndef S2G_3D_MINB
#define S2G_3D_MINB 2     // min CTAs/SM of k_scatter3d: 128 registers, no spills (uncapped the compiler has taken 137 -> ONE
                          // CTA per SM); -DS2G_3D_MINB=3 -DS2G_3D_CAP=640 is the measured-slower A/B variant
#endif
constexpr int S3_CAP = S2G_3D_CAP;

// Where pass B puts a cell's two contributions.  TILE: a CTA-owned shared-memory tile of T3^3 cells x 2 planes around
// the 8^3 block the particle's centre lies in (k_scatter3d_tile): shared-memory atomics instead of L2 reds, one flush of
// the tile per work item; cells of a large kernel that reach beyond the tile still go to the image directly.
constexpr int T3 = 24;          // tile edge: an 8^3 block + a halo of 8 cells on every side
constexpr int B3 = 8;           // block edge
struct Sink3 {
    double* tile;               // T3^3 x 2 doubles, (weight, quantity) interleaved
    int oi, oj, ok;             // image cell of tile cell (0,0,0)
};

template <bool TILE>
__device__ __forceinline__ void deposit_cell(const Sink3& sk, double* __restrict__ image, long long npl, long long n, int i,
                                             int j, int k, double w, double q)
{
    if (TILE) {
        const unsigned ti = (unsigned)(i - sk.oi), tj = (unsigned)(j - sk.oj), tk = (unsigned)(k - sk.ok);
        if (ti < (unsigned)T3 && tj < (unsigned)T3 && tk < (unsigned)T3) {
            double* c = sk.tile + 2 * ((ti * T3 + tj) * T3 + tk);
            atomicAdd(c, w);
            atomicAdd(c + 1, q);
            return;
        }
    }
    const long long idx = ((long long)i * n + j) * n + k;   // indices.jl:15-17
    red_add(image + npl + idx, w);
    red_add(image + idx, q);
}

template <int KID, bool TILE>
__device__ __forceinline__ void warp_deposit_3d(const Rec3& r, const s2g_geom& G, int lane,
                                                double* __restrict__ image, unsigned long long& touched,
                                                unsigned long long& fallback, double* __restrict__ s_w,
                                                unsigned* __restrict__ s_c, const Sink3& sk)
{
    const int ni = r.hi[0] - r.lo[0] + 1, nj = r.hi[1] - r.lo[1] + 1, nk = r.hi[2] - r.lo[2] + 1;
    const int lw = nk >= 32 ? 5 : (nk <= 1 ? 0 : 32 - __clz(nk - 1));
    const int W = 1 << lw, R = 32 >> lw;
    const int c0 = lane & (W - 1), r0 = lane >> lw;

    const double dx_lo = overlap_1d(r.x, r.h, r.lo[0]), dx_hi = overlap_1d(r.x, r.h, r.hi[0]);
    const double dy_lo = overlap_1d(r.y, r.h, r.lo[1]), dy_hi = overlap_1d(r.y, r.h, r.hi[1]);
    const double dz_lo = overlap_1d(r.z, r.h, r.lo[2]), dz_hi = overlap_1d(r.z, r.h, r.hi[2]);
    const double hinv = r.hinv;
    const double xb = center_dist(r.x, (double)r.lo[0]) * hinv;  // a of the first i-plane

    // ---- pass A (calculate_weights, cic_3D.jl:13-78); loops are warp-uniform (lanes without a column idle) so that
    // the list compaction can ballot
    double sw = 0.0;
    int cnt = 0;
    // list entries: TILE keeps the box coordinates (8 bits each), the red path the cell's linear offset from the box
    // corner (32 bits: ni n^2 < 2^32), so that pass B is one widening multiply-add per plane
    const unsigned nn1 = (unsigned)G.npix, nn2 = nn1 * nn1;
    bool cache = (s_w != nullptr) &&
                 (TILE ? (ni < 256 && nj < 256 && nk < 256) : ((long long)ni * G.npix * G.npix < (1LL << 32)));
    int n_list = 0;                 // uniform
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int kb = 0; kb < nk; kb += W) {
        const int kc = kb + c0;
        const int k = r.lo[2] + kc;
        const double cz = center_dist(r.z, (double)k) * hinv;
        const double dz = (k == r.lo[2]) ? dz_lo : ((k == r.hi[2]) ? dz_hi : 1.0);
        for (int jb = 0; jb < nj; jb += R) {
            const int jr = jb + r0;
            const int j = r.lo[1] + jr;
            const double by = center_dist(r.y, (double)j) * hinv;
            const double bc2 = fma(by, by, cz * cz) + 1e-300;  // > 0 even when a cell centre sits on the particle
            const bool colv = (kc < nk) && (jr < nj) && (bc2 < 1.0);
            if (!__any_sync(0xffffffffu, colv)) continue;
            const double dy = (j == r.lo[1]) ? dy_lo : ((j == r.hi[1]) ? dy_hi : 1.0);
            const double dydz = dy * dz;
            const unsigned colo = TILE ? (((unsigned)jr << 8) | (unsigned)kc) : ((unsigned)jr * nn1 + (unsigned)kc);
            double col = 0.0;
            for (int ii = 0; ii < ni; ii += 4) {
                double wk[4];
                bool in[4];
#pragma unroll
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double a = fma(-(double)(ii + q), hinv, xb);
                    const double s = fma(a, a, bc2);
                    in[q] = colv && below_one(s) && (ii + q < ni);
                    wk[q] = shape_s<KID>(s);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = r.lo[0] + ii + q;
                    const double dx = (i == r.lo[0]) ? dx_lo : ((i == r.hi[0]) ? dx_hi : 1.0);
                    const double wq = select_or_zero(in[q], wk[q]);
                    col = fma(wq, dx, col);
                    cnt += in[q] ? 1 : 0;
                    if (cache) {
                        const double gq = wq * (dx * dydz);     // the cell's wk·dV of pass B
                        const bool live = nonzero_bits(gq);
                        const unsigned m = __ballot_sync(0xffffffffu, live);
                        const int pos = n_list + __popc(m & lt_mask);
                        if (live && pos < S3_CAP) {
                            s_w[pos] = gq;
                            s_c[pos] = TILE ? (((unsigned)(ii + q) << 16) | colo) : ((unsigned)(ii + q) * nn2 + colo);
                        }
                        n_list += __popc(m);
                    }
                }
            }
            sw = fma(col, dydz, sw);
        }
    }
    cache = cache && n_list <= S3_CAP;
    sw = warp_sum(sw);
    cnt = __reduce_add_sync(0xffffffffu, cnt);

    bool fb = false;
    double n_distr, wpp;
    if (sw == 0.0) {  // cic_3D.jl:57-72
        fb = true;
        double dv = 0.0;
        for (int kc = c0; kc < nk; kc += W) {
            const int k = r.lo[2] + kc;
            const double dz = (k == r.lo[2]) ? dz_lo : ((k == r.hi[2]) ? dz_hi : 1.0);
            for (int jr = r0; jr < nj; jr += R) {
                const int j = r.lo[1] + jr;
                const double dy = (j == r.lo[1]) ? dy_lo : ((j == r.hi[1]) ? dy_hi : 1.0);
                for (int ii = 0; ii < ni; ++ii) {
                    const int i = r.lo[0] + ii;
                    const double dx = (i == r.lo[0]) ? dx_lo : ((i == r.hi[0]) ? dx_hi : 1.0);
                    dv += dx * dy * dz;
                }
            }
        }
        dv = warp_sum(dv);
        n_distr = (double)ni * (double)nj * (double)nk;
        wpp = (dv != 0.0) ? n_distr / dv : 1.0;
        if (lane == 0) ++fallback;
    } else {
        n_distr = (double)cnt;
        wpp = n_distr / sw;
    }
    const double kernel_norm = r.vol / n_distr;                       // cic_3D.jl:168
    const double volume_norm = kernel_norm * wpp * r.w * G.len2pix;   // :169
    const bool poison = !isfinite(volume_norm);

    // ---- pass B (cic_3D.jl:172-188)
    const long long n = G.npix, npl = n * n * n;
    if (fb || poison) {
        // rare: wk := 1 over the whole box (no cell centre covered), or an Inf/NaN norm that marks the whole box
        for (int kc = c0; kc < nk; kc += W) {
            const int k = r.lo[2] + kc;
            const double cz = center_dist(r.z, (double)k) * hinv;
            const double dz = (k == r.lo[2]) ? dz_lo : ((k == r.hi[2]) ? dz_hi : 1.0);
            for (int jr = r0; jr < nj; jr += R) {
                const int j = r.lo[1] + jr;
                const double by = center_dist(r.y, (double)j) * hinv;
                const double dy = (j == r.lo[1]) ? dy_lo : ((j == r.hi[1]) ? dy_hi : 1.0);
                for (int ii = 0; ii < ni; ++ii) {
                    const int i = r.lo[0] + ii;
                    const double a = fma(-(double)ii, hinv, xb);
                    const double s = fma(a, a, fma(by, by, cz * cz)) + 1e-300;
                    const double wk = fb ? 1.0 : ((s < 1.0) ? shape_s<KID>(s) : 0.0);
                    const double dx = (i == r.lo[0]) ? dx_lo : ((i == r.hi[0]) ? dx_hi : 1.0);
                    const double pw = wk * (dx * dy * dz) * volume_norm;
                    if (pw != 0.0) {
                        deposit_cell<false>(sk, image, npl, n, i, j, k, pw, r.q * pw);   // rare: straight to the image
                        ++touched;
                    }
                }
            }
        }
        return;
    }
    const double vq = volume_norm * r.q;
    const bool live_p = nonzero_bits(volume_norm);
    if (cache) {
        if (!live_p) return;
        __syncwarp();
        const long long o0 = (long long)r.lo[0] * n * n + (long long)r.lo[1] * n + r.lo[2];
        double* __restrict__ img_q = image + o0;
        double* __restrict__ img_w = image + npl + o0;
        for (int t = lane; t < n_list; t += 32) {
            const double g = s_w[t];
            const unsigned c = s_c[t];
            if (TILE)
                deposit_cell<true>(sk, image, npl, n, r.lo[0] + (int)(c >> 16), r.lo[1] + (int)((c >> 8) & 255u),
                                   r.lo[2] + (int)(c & 255u), g * volume_norm, g * vq);
            else {
                red_add(img_w + c, g * volume_norm);
                red_add(img_q + c, g * vq);
            }
        }
        if (lane == 0) touched += (unsigned long long)n_list;
        __syncwarp();   // the list is reused by the next particle
        return;
    }
    for (int kc = c0; kc < nk; kc += W) {
        const int k = r.lo[2] + kc;
        const double cz = center_dist(r.z, (double)k) * hinv;
        const double dz = (k == r.lo[2]) ? dz_lo : ((k == r.hi[2]) ? dz_hi : 1.0);
        for (int jr = r0; jr < nj; jr += R) {
            const int j = r.lo[1] + jr;
            const double by = center_dist(r.y, (double)j) * hinv;
            const double bc2 = fma(by, by, cz * cz) + 1e-300;
            if (bc2 >= 1.0 || !live_p) continue;
            const double dy = (j == r.lo[1]) ? dy_lo : ((j == r.hi[1]) ? dy_hi : 1.0);
            const double dydz = dy * dz;
            double* __restrict__ base = image + (long long)j * n + k;
            for (int ii = 0; ii < ni; ii += 4) {
                double wk[4];
                bool in[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double a = fma(-(double)(ii + q), hinv, xb);
                    const double s = fma(a, a, bc2);
                    in[q] = below_one(s) && (ii + q < ni);
                    wk[q] = shape_s<KID>(s);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (!in[q]) continue;
                    const int i = r.lo[0] + ii + q;
                    const double dx = (i == r.lo[0]) ? dx_lo : ((i == r.hi[0]) ? dx_hi : 1.0);
                    const double g = wk[q] * (dx * dydz);
                    if (nonzero_bits(g)) {
                        if (TILE)
                            deposit_cell<true>(sk, image, npl, n, i, j, k, g * volume_norm, g * vq);
                        else {
                            const long long idx = (long long)i * n * n;  // indices.jl:15-17
                            red_add(base + npl + idx, g * volume_norm);
                            red_add(base + idx, g * vq);
                        }
                        ++touched;
                    }
                }
            }
        }
    }
}

template <int KID>
__global__ void __launch_bounds__(256, S2G_3D_MINB) k_scatter3d(s2g_particles P, s2g_geom G, const unsigned* __restrict__ order,
                                                   long long n_list, double* __restrict__ image,
                                                   unsigned long long* __restrict__ counters, int use_cache)
{
    const int lane = threadIdx.x & 31;
    extern __shared__ __align__(16) unsigned char s3_smem[];
    double* s_w = nullptr;
    unsigned* s_c = nullptr;
    if (use_cache) {
        s_w = reinterpret_cast<double*>(s3_smem) + (threadIdx.x >> 5) * S3_CAP;
        s_c = reinterpret_cast<unsigned*>(s3_smem + 8 * S3_CAP * sizeof(double)) + (threadIdx.x >> 5) * S3_CAP;
    }
    unsigned long long touched = 0, fallback = 0, mapped = 0, fpx = 0;
    constexpr int CHUNK = 4;
    for (;;) {
        long long base = 0;
        if (lane == 0) base = (long long)atomicAdd(&counters[CNT_WORK], (unsigned long long)CHUNK);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_list) break;
        const long long end = min(base + CHUNK, n_list);
        for (long long t = base; t < end; ++t) {
            const long long p = order ? (long long)order[t] : t;
            Rec3 r;
            if (!make_rec3(P, G, p, r)) continue;
            if (lane == 0) {
                ++mapped;
                fpx += (unsigned long long)(r.hi[0] - r.lo[0] + 1) * (unsigned long long)(r.hi[1] - r.lo[1] + 1) *
                       (unsigned long long)(r.hi[2] - r.lo[2] + 1);
            }
            warp_deposit_3d<KID, false>(r, G, lane, image, touched, fallback, s_w, s_c, Sink3{nullptr, 0, 0, 0});
        }
    }
    touched = (unsigned long long)warp_sum_ll((long long)touched);
    if (lane == 0) {
        if (touched) atomicAdd(&counters[CNT_TOUCHED], touched);
        if (fallback) atomicAdd(&counters[CNT_FALLBACK], fallback);
        if (mapped) atomicAdd(&counters[CNT_MAPPED], mapped);
        if (mapped) atomicAdd(&counters[CNT_SCATTER], mapped);
        if (fpx) atomicAdd(&counters[CNT_FOOTPRINT], fpx);
    }
}

// ---- block-per-block scatter with a shared-memory image tile (north_star (3): "block-per-particle with shared-memory
// image tiles for large footprints, flushed with red adds").  Particles are sorted by the 8^3-cell block of their centre;
// a CTA takes a block (a chunk of <= TILE_CH of its particles), zeroes a 24^3 x 2 tile in shared memory (216 KB, one
// CTA per SM), lets its 16 warps deposit one particle each — pass A as before, pass B with shared-memory atomics — and
// flushes the tile once.  C3: ~250 particles x ~1200 cells x 2 planes per block become 27 648 reds instead of 600 000.
constexpr int TILE_THREADS = 512;
constexpr int TILE_CH = 2048;

__global__ void __launch_bounds__(256) k_tile3_keys(s2g_particles P, s2g_geom G, const unsigned* __restrict__ list,
                                                    long long n_list, int nb8, unsigned* __restrict__ keys,
                                                    unsigned* __restrict__ idx)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_list) return;
    const long long p = list ? (long long)list[t] : t;
    unsigned key = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double x = fma(ld_pos(P, p, d), G.len2pix, G.half_n);
        int b = (int)floor(x * (1.0 / B3));
        b = min(max(b, 0), nb8 - 1);
        key = key * (unsigned)nb8 + (unsigned)b;
    }
    keys[t] = key;
    idx[t] = (unsigned)p;
}

__global__ void __launch_bounds__(256) k_tile3_bounds(const unsigned* __restrict__ keys, long long m,
                                                      unsigned* __restrict__ beg, unsigned* __restrict__ end)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= m) return;
    const unsigned k = keys[t];
    if (t == 0 || keys[t - 1] != k) beg[k] = (unsigned)t;
    if (t == m - 1 || keys[t + 1] != k) end[k] = (unsigned)(t + 1);
}

__global__ void __launch_bounds__(256) k_tile3_chunks(const unsigned* __restrict__ beg, const unsigned* __restrict__ end,
                                                      int nkeys, unsigned* __restrict__ nch)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nkeys) return;
    nch[t] = (end[t] - beg[t] + TILE_CH - 1) / TILE_CH;
}

template <int KID>
__global__ void __launch_bounds__(TILE_THREADS, 1) k_scatter3d_tile(s2g_particles P, s2g_geom G,
                                                                    const unsigned* __restrict__ sorted_idx,
                                                                    const unsigned* __restrict__ beg,
                                                                    const unsigned* __restrict__ end,
                                                                    const unsigned* __restrict__ chunk_begin, int nkeys,
                                                                    int nb8, unsigned total_chunks,
                                                                    double* __restrict__ image,
                                                                    unsigned long long* __restrict__ counters)
{
    extern __shared__ __align__(16) double s_tile[];   // T3^3 x 2
    __shared__ unsigned s_work[3];
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    constexpr int NW = TILE_THREADS / 32;
    constexpr int NCELL = T3 * T3 * T3;
    unsigned long long touched = 0, fallback = 0, mapped = 0, fpx = 0;
    const long long n = G.npix, npl = n * n * n;
    for (;;) {
        if (tid == 0) {
            const unsigned w = (unsigned)atomicAdd(&counters[CNT_WORK], 1ull);
            unsigned key = 0xffffffffu, b = 0, e = 0;
            if (w < total_chunks) {
                int lo = 0, hi = nkeys;  // last key with chunk_begin[key] <= w
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (chunk_begin[mid] <= w) lo = mid; else hi = mid;
                }
                key = (unsigned)lo;
                b = beg[lo] + (w - chunk_begin[lo]) * TILE_CH;
                e = min(b + TILE_CH, end[lo]);
            }
            s_work[0] = key; s_work[1] = b; s_work[2] = e;
        }
        for (int c = tid; c < 2 * NCELL; c += TILE_THREADS) s_tile[c] = 0.0;
        __syncthreads();
        const unsigned key = s_work[0], wb = s_work[1], we = s_work[2];
        if (key == 0xffffffffu) break;
        Sink3 sk;
        sk.tile = s_tile;
        sk.ok = (int)(key % (unsigned)nb8) * B3 - (T3 - B3) / 2;
        sk.oj = (int)((key / (unsigned)nb8) % (unsigned)nb8) * B3 - (T3 - B3) / 2;
        sk.oi = (int)(key / ((unsigned)nb8 * (unsigned)nb8)) * B3 - (T3 - B3) / 2;
        for (unsigned t = wb + wq; t < we; t += NW) {
            const long long p = sorted_idx[t];
            Rec3 r;
            if (!make_rec3(P, G, p, r)) continue;
            if (lane == 0) {
                ++mapped;
                fpx += (unsigned long long)(r.hi[0] - r.lo[0] + 1) * (unsigned long long)(r.hi[1] - r.lo[1] + 1) *
                       (unsigned long long)(r.hi[2] - r.lo[2] + 1);
            }
            warp_deposit_3d<KID, true>(r, G, lane, image, touched, fallback, nullptr, nullptr, sk);
        }
        __syncthreads();
        // flush: tile cell (ti, tj, tk) -> image cell (oi+ti, oj+tj, ok+tk); tk is the contiguous axis
        for (int c = tid; c < NCELL; c += TILE_THREADS) {
            const double w = s_tile[2 * c], q = s_tile[2 * c + 1];
            if (w != 0.0 || q != 0.0) {
                const int tk = c % T3, tj = (c / T3) % T3, ti = c / (T3 * T3);
                const long long gi = sk.oi + ti, gj = sk.oj + tj, gk = sk.ok + tk;
                if (gi >= 0 && gi < n && gj >= 0 && gj < n && gk >= 0 && gk < n) {
                    const long long idx = (gi * n + gj) * n + gk;
                    red_add(image + npl + idx, w);
                    red_add(image + idx, q);
                }
            }
        }
        __syncthreads();   // tile and s_work are reused
    }
    touched = (unsigned long long)warp_sum_ll((long long)touched);
    if (lane == 0) {
        if (touched) atomicAdd(&counters[CNT_TOUCHED], touched);
        if (fallback) atomicAdd(&counters[CNT_FALLBACK], fallback);
        if (mapped) atomicAdd(&counters[CNT_MAPPED], mapped);
        if (mapped) atomicAdd(&counters[CNT_SCATTER], mapped);
        if (fpx) atomicAdd(&counters[CNT_FOOTPRINT], fpx);
    }
}

template <int KID>
static int launch_scatter3d_tile_k(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, const int* list,
                                   long long n_list, double* image)
{
    const int nb8 = (int)((G.npix + B3 - 1) / B3);
    const long long nkeys_ll = (long long)nb8 * nb8 * nb8;
    const int nkeys = (int)nkeys_ll;
    void *d_k, *d_k2, *d_i, *d_i2, *d_tmp, *d_beg, *d_end, *d_nch, *d_cb;
    S2G_TRY(s2g_scratch(ctx, "o3_keys", sizeof(unsigned) * n_list, &d_k));
    S2G_TRY(s2g_scratch(ctx, "o3_keys2", sizeof(unsigned) * n_list, &d_k2));
    S2G_TRY(s2g_scratch(ctx, "o3_idx", sizeof(unsigned) * n_list, &d_i));
    S2G_TRY(s2g_scratch(ctx, "o3_idx2", sizeof(unsigned) * n_list, &d_i2));
    S2G_TRY(s2g_scratch(ctx, "t3_beg", sizeof(unsigned) * (nkeys + 1), &d_beg));
    S2G_TRY(s2g_scratch(ctx, "t3_end", sizeof(unsigned) * (nkeys + 1), &d_end));
    S2G_TRY(s2g_scratch(ctx, "t3_nch", sizeof(unsigned) * (nkeys + 1), &d_nch));
    S2G_TRY(s2g_scratch(ctx, "t3_cb", sizeof(unsigned) * (nkeys + 1), &d_cb));
    cudaStream_t st = ctx->stream;
    const int phs = s2g_phase_begin(ctx, PH_SORT);
    k_tile3_keys<<<(int)((n_list + 255) / 256), 256, 0, st>>>(P, G, reinterpret_cast<const unsigned*>(list), n_list, nb8,
                                                              (unsigned*)d_k, (unsigned*)d_i);
    S2G_CUDA(cudaGetLastError());
    int bits = 1;
    while ((1LL << bits) < nkeys_ll) ++bits;
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, (const unsigned*)d_k, (unsigned*)d_k2, (const unsigned*)d_i,
                                    (unsigned*)d_i2, (int)n_list, 0, bits, st);
    S2G_TRY(s2g_scratch(ctx, "g_sort_tmp", sb + 16, &d_tmp));
    S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, (const unsigned*)d_k, (unsigned*)d_k2, (const unsigned*)d_i,
                                             (unsigned*)d_i2, (int)n_list, 0, bits, st));
    S2G_CUDA(cudaMemsetAsync(d_beg, 0, sizeof(unsigned) * (nkeys + 1), st));
    S2G_CUDA(cudaMemsetAsync(d_end, 0, sizeof(unsigned) * (nkeys + 1), st));
    S2G_CUDA(cudaMemsetAsync(d_nch, 0, sizeof(unsigned) * (nkeys + 1), st));
    k_tile3_bounds<<<(int)((n_list + 255) / 256), 256, 0, st>>>((const unsigned*)d_k2, n_list, (unsigned*)d_beg,
                                                                (unsigned*)d_end);
    S2G_CUDA(cudaGetLastError());
    k_tile3_chunks<<<(nkeys + 255) / 256, 256, 0, st>>>((const unsigned*)d_beg, (const unsigned*)d_end, nkeys,
                                                        (unsigned*)d_nch);
    S2G_CUDA(cudaGetLastError());
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (const unsigned*)d_nch, (unsigned*)d_cb, nkeys + 1, st);
    S2G_TRY(s2g_scratch(ctx, "g_tmp", tb + 16, &d_tmp));
    S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, (const unsigned*)d_nch, (unsigned*)d_cb, nkeys + 1, st));
    unsigned h_chunks = 0;
    s2g_readback rb(ctx);
    S2G_CUDA(rb.add(&h_chunks, (unsigned*)d_cb + nkeys, sizeof(unsigned)));
    s2g_phase_end(ctx, phs);
    S2G_CUDA(rb.sync());
    ctx->launches += 7;
    if (h_chunks == 0) return S2G_OK;
    const size_t smem = sizeof(double) * 2 * T3 * T3 * T3;
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        S2G_CUDA(cudaFuncSetAttribute(k_scatter3d_tile<KID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ctx->device & 63] = true;
    }
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), st));
    const int blocks = (int)std::min<long long>((long long)h_chunks, (long long)ctx->sm_count);
    const int ph = s2g_phase_begin(ctx, PH_DEPOSIT);
    k_scatter3d_tile<KID><<<std::max(blocks, 1), TILE_THREADS, smem, st>>>(
        P, G, (const unsigned*)d_i2, (const unsigned*)d_beg, (const unsigned*)d_end, (const unsigned*)d_cb, nkeys, nb8,
        h_chunks, image, ctx->d_counters);
    s2g_phase_end(ctx, ph);
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

// coarse spatial key of a particle: the 16^3-cell block holding its centre.  Particles are DEPOSITED in key order so
// that the ~2400 particles in flight update a compact region of the grid: with random order every red row misses L2
// (measured 474 GB of DRAM traffic for 2·10^10 reds on the c3s sample).
__global__ void __launch_bounds__(256) k_order3d_keys(s2g_particles P, s2g_geom G, const int* __restrict__ list,
                                                      long long n_list, unsigned* __restrict__ keys,
                                                      unsigned* __restrict__ idx)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_list) return;
    const long long p = list ? (long long)list[t] : t;
    const int nb = (int)((G.npix + 15) / 16);
    unsigned key = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double x = fma(ld_pos(P, p, d), G.len2pix, G.half_n);
        int b = (int)floor(x * (1.0 / 16.0));
        b = min(max(b, 0), nb - 1);
        key = key * (unsigned)nb + (unsigned)b;
    }
    keys[t] = key;
    idx[t] = (unsigned)p;
}

template <int KID>
static int launch_scatter3d_k(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, const int* list, long long n_list,
                              double* image)
{
    if (n_list <= 0) return S2G_OK;
    // shared-memory tile path: OFF by default (S2G_3D_TILE=1 switches it on).  Measured on C3 (profiles/
    // r2_3d_experiments.txt): 2470 ms per step with the default 8 Mi slices, 1437 ms with one 64 Mi slice, against
    // 1010 ms for the red.global path — a shared-memory FP64 atomicAdd costs the SM issue slots (the kernel is bound by
    // its instruction stream), while red.global is one fire-and-forget instruction whose work is done by the L2.
    const char* e_t = getenv("S2G_3D_TILE");
    const bool tile_on = e_t ? atoi(e_t) != 0 : false;
    const long long nb8 = (G.npix + B3 - 1) / B3;
    if (tile_on && n_list >= 32768 && G.npix >= T3 && nb8 * nb8 * nb8 < (1LL << 31) &&
        n_list * 64 >= nb8 * nb8 * nb8)   // on average at least 1/64 particle per block, else the flushes dominate
        return launch_scatter3d_tile_k<KID>(ctx, P, G, list, n_list, image);
    const unsigned* order = reinterpret_cast<const unsigned*>(list);
    const char* e_ord = getenv("S2G_3D_ORDER");
    // the class list of the AUTO strategy (list != nullptr) is in input order: it is put in key order like the whole set
    if ((e_ord ? atoi(e_ord) != 0 : true) && n_list >= 65536 && n_list < (1LL << 31)) {
        void *d_k, *d_k2, *d_i, *d_i2, *d_tmp;
        S2G_TRY(s2g_scratch(ctx, "o3_keys", sizeof(unsigned) * n_list, &d_k));
        S2G_TRY(s2g_scratch(ctx, "o3_keys2", sizeof(unsigned) * n_list, &d_k2));
        S2G_TRY(s2g_scratch(ctx, "o3_idx", sizeof(unsigned) * n_list, &d_i));
        S2G_TRY(s2g_scratch(ctx, "o3_idx2", sizeof(unsigned) * n_list, &d_i2));
        const int phs = s2g_phase_begin(ctx, PH_SORT);
        k_order3d_keys<<<(int)((n_list + 255) / 256), 256, 0, ctx->stream>>>(P, G, list, n_list, (unsigned*)d_k,
                                                                               (unsigned*)d_i);
        S2G_CUDA(cudaGetLastError());
        const int nb = (int)((G.npix + 15) / 16);
        int bits = 1;
        while ((1LL << bits) < (long long)nb * nb * nb) ++bits;
        size_t sb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sb, (const unsigned*)d_k, (unsigned*)d_k2, (const unsigned*)d_i,
                                        (unsigned*)d_i2, (int)n_list, 0, bits, ctx->stream);
        S2G_TRY(s2g_scratch(ctx, "g_sort_tmp", sb + 16, &d_tmp));
        S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, (const unsigned*)d_k, (unsigned*)d_k2, (const unsigned*)d_i,
                                                 (unsigned*)d_i2, (int)n_list, 0, bits, ctx->stream));
        s2g_phase_end(ctx, phs);
        ctx->launches += 4;
        order = (const unsigned*)d_i2;
    }
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    const int warps_needed = (int)std::min<long long>((n_list + 3) / 4, (long long)ctx->sm_count * 8 * 8);
    int blocks = max(1, (warps_needed + 7) / 8);
    blocks = min(blocks, ctx->sm_count * 8);
    // shared-memory cell lists (S3_CAP entries of 12 bytes per warp = 96 KB per CTA, 2 CTAs/SM); S2G_3D_CACHE=0: off
    const char* e_c = getenv("S2G_3D_CACHE");
    const int use_cache = e_c ? (atoi(e_c) != 0) : 1;
    const size_t smem = use_cache ? (size_t)8 * S3_CAP * (sizeof(double) + sizeof(unsigned)) : 0;
    static bool attr_set[64] = {};
    if (use_cache && !attr_set[ctx->device & 63]) {
        S2G_CUDA(cudaFuncSetAttribute(k_scatter3d<KID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ctx->device & 63] = true;
    }
    const int ph = s2g_phase_begin(ctx, PH_DEPOSIT);
    k_scatter3d<KID><<<blocks, 256, smem, ctx->stream>>>(P, G, order, n_list, image, ctx->d_counters, use_cache);
    s2g_phase_end(ctx, ph);
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

int s2g_launch_scatter_3d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, const int* list,
                          long long n_list, double* image)
{
    switch (kernel) {
    case S2G_KERNEL_CUBIC: return launch_scatter3d_k<S2G_KERNEL_CUBIC>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_QUINTIC: return launch_scatter3d_k<S2G_KERNEL_QUINTIC>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C2: return launch_scatter3d_k<S2G_KERNEL_WENDLAND_C2>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C4: return launch_scatter3d_k<S2G_KERNEL_WENDLAND_C4>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C6: return launch_scatter3d_k<S2G_KERNEL_WENDLAND_C6>(ctx, P, G, list, n_list, image);
    case S2G_KERNEL_WENDLAND_C8: return launch_scatter3d_k<S2G_KERNEL_WENDLAND_C8>(ctx, P, G, list, n_list, image);
    }
    s2g_set_error("unknown kernel id %d", kernel);
    return S2G_EINVAL;
}

// ------------------------------------------------------------------------------------------------
// reduce_image_3D: same memory order as the flat buffer; divide gated on the QUANTITY plane (reduce_image.jl:49-51);
// !reduce_image means the weight plane was overwritten with 1 (cic_interpolation.jl:230-232)
// ------------------------------------------------------------------------------------------------
__global__ void k_reduce3d(const double* __restrict__ image, long long ncell, int reduce_image, double* __restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ncell; e += stride) {
        double v = image[e];
        if (v > 0.0) {
            const double wv = reduce_image ? image[e + ncell] : 1.0;
            v = v / wv;
        }
        out[e] = v;
    }
}

int s2g_launch_reduce_3d(s2g_ctx* ctx, const double* image, long long npix, int reduce_image, double* out)
{
    const long long ncell = npix * npix * npix;
    const int blocks = (int)std::min<long long>((ncell + 255) / 256, (long long)ctx->sm_count * 16);
    const int ph = s2g_phase_begin(ctx, PH_EPILOGUE);
    k_reduce3d<<<max(blocks, 1), 256, 0, ctx->stream>>>(image, ncell, reduce_image, out);
    s2g_phase_end(ctx, ph);
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}
