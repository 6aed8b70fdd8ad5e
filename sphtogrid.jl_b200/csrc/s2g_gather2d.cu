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

namespace {

constexpr int TILE = 64;       // tile edge in pixels
constexpr int RPT = 16;        // rows per thread  (256 threads: 64 columns x 4 row groups x 16 rows)
constexpr int BATCH = 128;     // particle records staged in shared memory at a time
constexpr int CHUNK = 2048;    // max pairs per work item

struct __align__(16) GRec {
    double x, y;          // pixel coordinates
    double hinv;          // 1/h ; negative => "no pixel centre covered" branch (wk := 1 everywhere)
    double h2lim;         // h^2 (1+1e-14): conservative pre-test for u <= 1
    double an;            // area_norm
    double dx_lo, dx_hi, dy_lo, dy_hi;  // edge overlaps
    int iMin, iMax, jMin, jMax;
    int p;                // particle index (for the per-image quantity)
    int pad;
};

__device__ __forceinline__ int tile_count(const Rec2& r, int& ti0, int& ti1, int& tj0, int& tj1)
{
    ti0 = r.iMin / TILE; ti1 = r.iMax / TILE; tj0 = r.jMin / TILE; tj1 = r.jMax / TILE;
    return (ti1 - ti0 + 1) * (tj1 - tj0 + 1);
}

// does the kernel support (circle of radius h around (x,y)) possibly reach a pixel centre of tile (ti,tj)?
// conservative (never rejects a tile that holds a pixel with u <= 1); used identically by count and expand.
__device__ __forceinline__ bool tile_hit(double x, double y, double h, int ti, int tj)
{
    const double lo_i = ti * (double)TILE + 0.5, hi_i = lo_i + (TILE - 1);
    const double lo_j = tj * (double)TILE + 0.5, hi_j = lo_j + (TILE - 1);
    const double ddx = fmax(fmax(lo_i - x, 0.0), x - hi_i);
    const double ddy = fmax(fmax(lo_j - y, 0.0), y - hi_j);
    const double hh = h * (1.0 + 1e-9) + 1e-9;
    return ddx * ddx + ddy * ddy <= hh * hh;
}

// class: 0 = nothing to do, 1 = scatter, 2 = gather
__global__ void __launch_bounds__(256) k_classify(s2g_particles P, s2g_geom G, long long p0, long long nb,
                                                  long long gather_min_pixels, int force, int* __restrict__ cls,
                                                  unsigned* __restrict__ npairs)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nb) return;
    Rec2 r;
    int c = 0;
    unsigned np = 0;
    if (make_rec2(P, G, p0 + t, r)) {
        const long long fp = (long long)(r.iMax - r.iMin + 1) * (r.jMax - r.jMin + 1);
        const bool gather = force == S2G_STRATEGY_GATHER || (force == S2G_STRATEGY_AUTO && fp >= gather_min_pixels);
        c = gather ? 2 : 1;
        if (gather) {
            int ti0, ti1, tj0, tj1;
            tile_count(r, ti0, ti1, tj0, tj1);
            // pairs are counted after pass A (the fallback branch must keep every tile); upper bound here
            np = (unsigned)((ti1 - ti0 + 1) * (tj1 - tj0 + 1));
        }
    }
    cls[t] = c;
    npairs[t] = np;
}

// compact lists of scatter and gather particles (indices relative to the whole particle set)
__global__ void __launch_bounds__(256) k_build_lists(const int* __restrict__ cls, const unsigned* __restrict__ pos_s,
                                                     const unsigned* __restrict__ pos_g, long long p0, long long nb,
                                                     int* __restrict__ list_s, int* __restrict__ list_g)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nb) return;
    const int c = cls[t];
    if (c == 1) list_s[pos_s[t]] = (int)(p0 + t);
    if (c == 2) list_g[pos_g[t]] = (int)(p0 + t);
}

// ---- pass A for gather particles: one warp per particle, lanes along j, rows looped
template <int KID>
__global__ void __launch_bounds__(256) k_norm2d(s2g_particles P, s2g_geom G, const int* __restrict__ list,
                                                long long n_list, GRec* __restrict__ recs,
                                                unsigned* __restrict__ npairs_g,
                                                unsigned long long* __restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    unsigned long long fallback = 0, mapped = 0, fpx = 0;
    for (;;) {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(&counters[CNT_WORK], 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_list) break;
        const long long p = list[t];
        Rec2 r;
        make_rec2(P, G, p, r);  // known valid
        const int ni = r.iMax - r.iMin + 1, nj = r.jMax - r.jMin + 1;
        const double dx_lo = overlap_1d(r.x, r.h, r.iMin), dx_hi = overlap_1d(r.x, r.h, r.iMax);
        const double dy_lo = overlap_1d(r.y, r.h, r.jMin), dy_hi = overlap_1d(r.y, r.h, r.jMax);
        const double h2lim = r.h * r.h * (1.0 + 1e-14);
        double sw = 0.0;
        int cnt = 0;
        for (int jc = lane; jc < nj; jc += 32) {
            const int j = r.jMin + jc;
            const double yd = center_dist(r.y, (double)j);
            const double yd2 = __dmul_rn(yd, yd);
            const double dy = (j == r.jMin) ? dy_lo : ((j == r.jMax) ? dy_hi : 1.0);
            if (yd2 > h2lim) continue;
            for (int ir = 0; ir < ni; ++ir) {
                const int i = r.iMin + ir;
                const double xd = center_dist(r.x, (double)i);
                const double xd2 = __dmul_rn(xd, xd);
                if (__dadd_rn(xd2, yd2) > h2lim) continue;
                const double u = u_of(xd2, yd2, r.hinv);
                if (u <= 1.0) {
                    const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
                    sw = fma(kernel_shape<KID>(u), dx * dy, sw);
                    ++cnt;
                }
            }
        }
        sw = warp_sum(sw);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        bool fb = false;
        double n_distr, wpp;
        if (sw == 0.0) {
            fb = true;
            double da = 0.0;
            for (int jc = lane; jc < nj; jc += 32) {
                const int j = r.jMin + jc;
                const double dy = (j == r.jMin) ? dy_lo : ((j == r.jMax) ? dy_hi : 1.0);
                for (int ir = 0; ir < ni; ++ir) {
                    const int i = r.iMin + ir;
                    const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
                    da += dx * dy;
                }
            }
            da = warp_sum(da);
            n_distr = (double)ni * (double)nj;
            wpp = (da != 0.0) ? n_distr / da : 1.0;
            ++fallback;
        } else {
            n_distr = (double)cnt;
            wpp = n_distr / sw;
        }
        const double kernel_norm = r.area / n_distr;
        const double area_norm = kernel_norm * wpp * r.w * r.dz;
        // exact pair count (tiles the kernel support can reach; every bbox tile in the fallback branch)
        int ti0, ti1, tj0, tj1;
        tile_count(r, ti0, ti1, tj0, tj1);
        const int nti = ti1 - ti0 + 1, ntj = tj1 - tj0 + 1;
        unsigned np = 0;
        for (int q = lane; q < nti * ntj; q += 32) {
            const int ti = ti0 + q / ntj, tj = tj0 + q % ntj;
            if (fb || tile_hit(r.x, r.y, r.h, ti, tj)) ++np;
        }
        np = __reduce_add_sync(0xffffffffu, np);
        if (lane == 0) {
            GRec g;
            g.x = r.x; g.y = r.y;
            g.hinv = fb ? -r.hinv : r.hinv;
            g.h2lim = h2lim;
            g.an = area_norm;
            g.dx_lo = dx_lo; g.dx_hi = dx_hi; g.dy_lo = dy_lo; g.dy_hi = dy_hi;
            g.iMin = r.iMin; g.iMax = r.iMax; g.jMin = r.jMin; g.jMax = r.jMax;
            g.p = (int)p; g.pad = 0;
            recs[t] = g;
            npairs_g[t] = np;
            ++mapped;
            fpx += (unsigned long long)ni * (unsigned long long)nj;
        }
    }
    if (lane == 0) {
        if (fallback) atomicAdd(&counters[CNT_FALLBACK], fallback);
        if (mapped) { atomicAdd(&counters[CNT_MAPPED], mapped); atomicAdd(&counters[CNT_GATHER], mapped); }
        if (fpx) atomicAdd(&counters[CNT_FOOTPRINT], fpx);
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
    const bool fb = g.hinv < 0;
    const double h = 1.0 / fabs(g.hinv);
    const int ti0 = g.iMin / TILE, ti1 = g.iMax / TILE, tj0 = g.jMin / TILE, tj1 = g.jMax / TILE;
    unsigned o = off[t];
    for (int ti = ti0; ti <= ti1; ++ti)
        for (int tj = tj0; tj <= tj1; ++tj)
            if (fb || tile_hit(g.x, g.y, h * (1.0 + 1e-12), ti, tj)) {
                keys[o] = (unsigned)(ti * ntile_j + tj);
                vals[o] = (unsigned)t;
                ++o;
            }
}

// tile_begin[k] .. tile_begin[k+1] after an exclusive scan of per-tile counts
__global__ void __launch_bounds__(256) k_tile_hist(const unsigned* __restrict__ keys, long long m,
                                                   unsigned* __restrict__ tile_cnt)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= m) return;
    atomicAdd(&tile_cnt[keys[t]], 1u);
}

__global__ void __launch_bounds__(256) k_tile_chunks(const unsigned* __restrict__ tile_cnt, int ntiles,
                                                     unsigned* __restrict__ nchunks)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    nchunks[t] = (tile_cnt[t] + CHUNK - 1) / CHUNK;
}

// ---- the gather kernel
template <int KID>
__global__ void __launch_bounds__(256, 2) k_gather2d(const GRec* __restrict__ recs, const unsigned* __restrict__ vals,
                                                     const unsigned* __restrict__ tile_begin,   // ntiles+1
                                                     const unsigned* __restrict__ chunk_begin,  // ntiles+1
                                                     int ntiles, int ntile_j, unsigned total_chunks,
                                                     const void* __restrict__ binq, int in_dtype, int n_images,
                                                     int image_k, long long npix, double* __restrict__ image,
                                                     unsigned long long* __restrict__ counters)
{
    __shared__ GRec s_rec[BATCH];
    __shared__ double s_q[BATCH];
    __shared__ unsigned s_work[3];

    const int tid = threadIdx.x, lane = tid & 31;
    const int jl = tid & (TILE - 1);         // column inside the tile
    const int rg = tid >> 6;                 // row group 0..3
    unsigned long long touched = 0;

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
                b = tile_begin[lo] + c * CHUNK;
                e = min(b + CHUNK, tile_begin[lo + 1]);
            }
            s_work[0] = tile; s_work[1] = b; s_work[2] = e;
        }
        __syncthreads();
        const unsigned tile = s_work[0], wb = s_work[1], we = s_work[2];
        if (tile == 0xffffffffu) break;
        const int i0 = (int)(tile / ntile_j) * TILE, j0 = (int)(tile % ntile_j) * TILE;
        const int j = j0 + jl;
        const int ibase = i0 + rg * RPT;
        const double jd = (double)j;

        double acc_w[RPT], acc_q[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) { acc_w[r] = 0.0; acc_q[r] = 0.0; }

        for (unsigned b = wb; b < we; b += BATCH) {
            const int nb = (int)min((unsigned)BATCH, we - b);
            __syncthreads();  // previous batch fully consumed
            if (tid < nb) {
                const GRec g = recs[vals[b + tid]];
                s_rec[tid] = g;
                s_q[tid] = ld_in(binq, (long long)n_images * g.p + image_k, in_dtype);
            }
            __syncthreads();
            for (int e = 0; e < nb; ++e) {
                const GRec& g = s_rec[e];
                // warp-uniform row cull
                const int rlo = max(g.iMin, ibase), rhi = min(g.iMax, ibase + RPT - 1);
                if (rlo > rhi) continue;
                const bool col_in = (j >= g.jMin) && (j <= g.jMax);
                const bool fb = g.hinv < 0;
                const double yd = center_dist(g.y, jd);
                const double yd2 = __dmul_rn(yd, yd);
                const bool col_live = col_in && (fb || yd2 <= g.h2lim);
                if (!__any_sync(0xffffffffu, col_live)) continue;
                const double dy = (j == g.jMin) ? g.dy_lo : ((j == g.jMax) ? g.dy_hi : 1.0);
                const double hinv = fabs(g.hinv);
                const double dyan = dy * g.an;
                const double q = s_q[e];
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int i = ibase + r;
                    if (i < rlo || i > rhi) continue;  // uniform
                    if (!col_live) continue;
                    const double xd = center_dist(g.x, (double)i);
                    const double xd2 = __dmul_rn(xd, xd);
                    double wk;
                    if (fb)
                        wk = 1.0;
                    else {
                        if (__dadd_rn(xd2, yd2) > g.h2lim) continue;
                        const double u = u_of(xd2, yd2, hinv);
                        if (!(u <= 1.0)) continue;
                        wk = kernel_shape<KID>(u);
                    }
                    const double dx = (i == g.iMin) ? g.dx_lo : ((i == g.iMax) ? g.dx_hi : 1.0);
                    const double pw = wk * dx * dyan;
                    if (pw != 0.0) {
                        acc_w[r] += pw;
                        acc_q[r] = fma(q, pw, acc_q[r]);
                        ++touched;
                    }
                }
            }
        }
        // flush: coalesced along j
        const long long npl = npix * npix;
        if (j < npix) {
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int i = ibase + r;
                if (i < npix && (acc_w[r] != 0.0 || acc_q[r] != 0.0)) {
                    const long long idx = (long long)i * npix + j;
                    if (image_k == 0) red_add(image + npl * n_images + idx, acc_w[r]);
                    red_add(image + npl * image_k + idx, acc_q[r]);
                }
            }
        }
        __syncthreads();  // s_work reuse
    }
    if (image_k == 0) {
        touched = (unsigned long long)warp_sum_ll((long long)touched);
        if (lane == 0 && touched) atomicAdd(&counters[CNT_TOUCHED], touched);
    }
}

struct KLaunch {
    int (*norm)(s2g_ctx*, const s2g_particles&, const s2g_geom&, const int*, long long, GRec*, unsigned*);
    int (*gather)(s2g_ctx*, const GRec*, const unsigned*, const unsigned*, const unsigned*, int, int, unsigned,
                  const s2g_particles&, const s2g_geom&, int, double*);
};

template <int KID>
int launch_norm(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, const int* list, long long n_list, GRec* recs,
                unsigned* npairs_g)
{
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    const int blocks = (int)std::min<long long>((n_list + 7) / 8, (long long)ctx->sm_count * 8);
    k_norm2d<KID><<<max(blocks, 1), 256, 0, ctx->stream>>>(P, G, list, n_list, recs, npairs_g, ctx->d_counters);
    S2G_CUDA(cudaGetLastError());
    return S2G_OK;
}

template <int KID>
int launch_gather(s2g_ctx* ctx, const GRec* recs, const unsigned* vals, const unsigned* tile_begin,
                  const unsigned* chunk_begin, int ntiles, int ntile_j, unsigned total_chunks, const s2g_particles& P,
                  const s2g_geom& G, int image_k, double* image)
{
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    const int blocks = (int)std::min<long long>((long long)total_chunks, (long long)ctx->sm_count * 2);
    k_gather2d<KID><<<max(blocks, 1), 256, 0, ctx->stream>>>(recs, vals, tile_begin, chunk_begin, ntiles, ntile_j,
                                                            total_chunks, P.binq, P.in_dtype, G.n_images, image_k,
                                                            G.npix, image, ctx->d_counters);
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
    if (ctx->strategy == S2G_STRATEGY_SCATTER)
        return s2g_launch_scatter_2d(ctx, P, G, kernel, nullptr, P.n, image);

    const long long gather_min = env_ll("S2G_GATHER_MIN_PIXELS", 1024);
    const long long batch_max = env_ll("S2G_BATCH_PARTICLES", 8LL << 20);
    const long long pair_cap = env_ll("S2G_PAIR_CAP", 512LL << 20);
    const int ntile_j = (int)((G.npix + TILE - 1) / TILE);
    const int ntiles = ntile_j * ntile_j;
    cudaStream_t st = ctx->stream;

    long long p0 = 0;
    long long batch = min(batch_max, P.n);
    while (p0 < P.n) {
        const long long nb = min(batch, P.n - p0);
        void *d_cls, *d_np, *d_ps, *d_pg, *d_ls, *d_lg, *d_tmp, *d_sum;
        S2G_TRY(s2g_scratch(ctx, "g_cls", sizeof(int) * (nb + 1), &d_cls));
        S2G_TRY(s2g_scratch(ctx, "g_np", sizeof(unsigned) * (nb + 1), &d_np));
        S2G_TRY(s2g_scratch(ctx, "g_pos_s", sizeof(unsigned) * (nb + 1), &d_ps));
        S2G_TRY(s2g_scratch(ctx, "g_pos_g", sizeof(unsigned) * (nb + 1), &d_pg));
        S2G_TRY(s2g_scratch(ctx, "g_list_s", sizeof(int) * nb, &d_ls));
        S2G_TRY(s2g_scratch(ctx, "g_list_g", sizeof(int) * nb, &d_lg));
        S2G_TRY(s2g_scratch(ctx, "g_sum", sizeof(unsigned long long), &d_sum));
        const int blocks = (int)((nb + 255) / 256);
        S2G_CUDA(cudaMemsetAsync((int*)d_cls + nb, 0, sizeof(int), st));
        k_classify<<<blocks, 256, 0, st>>>(P, G, p0, nb, gather_min, ctx->strategy, (int*)d_cls, (unsigned*)d_np);
        S2G_CUDA(cudaGetLastError());

        cub::TransformInputIterator<unsigned, IsClass, const int*> it_s((const int*)d_cls, IsClass{1});
        cub::TransformInputIterator<unsigned, IsClass, const int*> it_g((const int*)d_cls, IsClass{2});
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
        S2G_CUDA(cub::DeviceReduce::Sum(d_tmp, tb, it_np, (unsigned long long*)d_sum, (int)nb, st));
        unsigned h_ns = 0, h_ng = 0;
        unsigned long long h_ub = 0;
        S2G_CUDA(cudaMemcpyAsync(&h_ns, (unsigned*)d_ps + nb, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        S2G_CUDA(cudaMemcpyAsync(&h_ng, (unsigned*)d_pg + nb, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        S2G_CUDA(cudaMemcpyAsync(&h_ub, d_sum, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        S2G_CUDA(cudaStreamSynchronize(st));
        if ((long long)h_ub > pair_cap && nb > 1024) {  // too many (tile,particle) pairs: shrink the slice, redo it
            batch = std::max<long long>(1024, nb / 2);
            continue;
        }
        S2G_CHECK(h_ub < 0xfff00000ull, S2G_ENOMEM,
                  "a single slice of %lld particles spans %llu image tiles: footprints too large for this image",
                  nb, h_ub);
        const long long n_s = h_ns, n_g = h_ng;
        k_build_lists<<<blocks, 256, 0, st>>>((const int*)d_cls, (const unsigned*)d_ps, (const unsigned*)d_pg, p0, nb,
                                              (int*)d_ls, (int*)d_lg);
        S2G_CUDA(cudaGetLastError());

        // ---- scatter bin (small footprints)
        if (n_s > 0) S2G_TRY(s2g_launch_scatter_2d(ctx, P, G, kernel, (const int*)d_ls, n_s, image));

        // ---- gather bin
        if (n_g > 0) {
            void *d_recs, *d_npg, *d_off;
            S2G_TRY(s2g_scratch(ctx, "g_recs", sizeof(GRec) * n_g, &d_recs));
            S2G_TRY(s2g_scratch(ctx, "g_npg", sizeof(unsigned) * (n_g + 1), &d_npg));
            S2G_TRY(s2g_scratch(ctx, "g_off", sizeof(unsigned) * (n_g + 1), &d_off));
            S2G_CUDA(cudaMemsetAsync((unsigned*)d_npg + n_g, 0, sizeof(unsigned), st));
            S2G_TRY(K.norm(ctx, P, G, (const int*)d_lg, n_g, (GRec*)d_recs, (unsigned*)d_npg));
            size_t tb4 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb4, (const unsigned*)d_npg, (unsigned*)d_off, (int)(n_g + 1), st);
            S2G_TRY(s2g_scratch(ctx, "g_tmp", tb4 + 16, &d_tmp));
            S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb4, (const unsigned*)d_npg, (unsigned*)d_off,
                                                   (int)(n_g + 1), st));
            unsigned h_m = 0;
            S2G_CUDA(cudaMemcpyAsync(&h_m, (unsigned*)d_off + n_g, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            S2G_CUDA(cudaStreamSynchronize(st));
            const long long m = h_m;
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
                k_tile_hist<<<(int)((m + 255) / 256), 256, 0, st>>>((const unsigned*)d_keys2, m, (unsigned*)d_tcnt);
                S2G_CUDA(cudaGetLastError());
                S2G_CUDA(cudaMemsetAsync(d_nch, 0, sizeof(unsigned) * (ntiles + 1), st));
                k_tile_chunks<<<(ntiles + 255) / 256, 256, 0, st>>>((const unsigned*)d_tcnt, ntiles, (unsigned*)d_nch);
                S2G_CUDA(cudaGetLastError());
                size_t tb3 = 0;
                cub::DeviceScan::ExclusiveSum(nullptr, tb3, (const unsigned*)d_tcnt, (unsigned*)d_tbeg, ntiles + 1, st);
                S2G_TRY(s2g_scratch(ctx, "g_tmp", tb3 + 16, &d_tmp));
                size_t tbb = tb3 + 16;
                S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tbb, (const unsigned*)d_tcnt, (unsigned*)d_tbeg,
                                                       ntiles + 1, st));
                tbb = tb3 + 16;
                S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tbb, (const unsigned*)d_nch, (unsigned*)d_cbeg,
                                                       ntiles + 1, st));
                unsigned h_chunks = 0;
                S2G_CUDA(cudaMemcpyAsync(&h_chunks, (unsigned*)d_cbeg + ntiles, sizeof(unsigned),
                                         cudaMemcpyDeviceToHost, st));
                S2G_CUDA(cudaStreamSynchronize(st));
                for (int k = 0; k < G.n_images; ++k)
                    S2G_TRY(K.gather(ctx, (const GRec*)d_recs, (const unsigned*)d_vals2, (const unsigned*)d_tbeg,
                                     (const unsigned*)d_cbeg, ntiles, ntile_j, h_chunks, P, G, k, image));
                ctx->host_pairs += m;
            }
        }
        p0 += nb;
    }
    return S2G_OK;
}
