// s2g_gather3d.cu — 3D Smac deposit, gather strategy, and the per-particle strategy dispatch.
//
// Same arithmetic as the 3D scatter path (cic_3D.jl:13-78, :110-209), organised like the 2D gather
// (s2g_gather2d.cu) so that the bulk of the work needs no global atomics:
//   k_classify3d : per particle -> class (skip / scatter / gather) + upper bound of the tiles it touches
//   k_norm3d     : pass A (Σ w·dV over the cell centres inside the kernel), one warp per particle -> GRec3
//   k_expand3d   : (tile, particle) pairs ; cub radix sort by tile ; tile ranges
//   k_gather3d   : a CTA owns an 8 x 16 x 16 tile (i x j x k); thread (j,k) owns the 8 cells of its i-column and keeps
//                  their weight and quantity sums in registers while particle records stream through shared memory;
//                  one red.add flush per work item, coalesced along k.
// Footprints of fewer than S2G_GATHER3D_MIN_CELLS cells go to the scatter kernel.
#include <cub/cub.cuh>
#include <cstdlib>

#include "s2g_cic3d.cuh"

int s2g_launch_scatter_3d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, const int* list,
                          long long n_list, double* image);

namespace {

constexpr int T_I = 8, T_J = 16, T_K = 16;  // tile extent; 256 threads = T_J x T_K columns of T_I cells
constexpr int BATCH = 128;                  // particle records staged in shared memory at a time
constexpr int CHUNK = 4096;                 // max pairs per work item

struct __align__(16) GRec3 {
    int lo[3], hi[3];       // clipped footprint; lo[0] > hi[0]: record unused (particle re-routed to the scatter kernel)
    int p, pad;
    double x, y, z, hinv;   // cell coordinates of the particle, 1/h
    double vn, vq;          // volume_norm (cic_3D.jl:169) and volume_norm * bin_q
    double dlo[3], dhi[3];  // overlap lengths of the first / last cell per axis (interior: 1)
    double h;
};

__device__ __forceinline__ bool tile_hit3(const GRec3& g, int ti, int tj, int tk)
{
    const double lo[3] = {ti * (double)T_I + 0.5, tj * (double)T_J + 0.5, tk * (double)T_K + 0.5};
    const double hi[3] = {lo[0] + (T_I - 1), lo[1] + (T_J - 1), lo[2] + (T_K - 1)};
    const double c[3] = {g.x, g.y, g.z};
    double d2 = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double a = lo[d] - c[d], b = c[d] - hi[d];
        const double dd = a > 0.0 ? a : (b > 0.0 ? b : 0.0);
        d2 += dd * dd;
    }
    const double hh = g.h * (1.0 + 1e-9) + 1e-9;
    return d2 <= hh * hh;
}

__global__ void __launch_bounds__(256) k_classify3d(s2g_particles P, s2g_geom G, long long p0, long long nb,
                                                    long long gather_min_cells, int force, int* __restrict__ cls,
                                                    unsigned* __restrict__ npairs)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nb) return;
    Rec3 r;
    int c = 0;
    unsigned np = 0;
    if (make_rec3(P, G, p0 + t, r)) {
        const long long fp = (long long)(r.hi[0] - r.lo[0] + 1) * (r.hi[1] - r.lo[1] + 1) * (r.hi[2] - r.lo[2] + 1);
        const bool gather = force == S2G_STRATEGY_GATHER || (force == S2G_STRATEGY_AUTO && fp >= gather_min_cells);
        c = gather ? 2 : 1;
        if (gather)
            np = (unsigned)((r.hi[0] / T_I - r.lo[0] / T_I + 1) * (r.hi[1] / T_J - r.lo[1] / T_J + 1) *
                            (r.hi[2] / T_K - r.lo[2] / T_K + 1));
    }
    cls[t] = c;
    npairs[t] = np;
}

__global__ void __launch_bounds__(256) k_build_lists3(const int* __restrict__ cls, const unsigned* __restrict__ pos_s,
                                                      const unsigned* __restrict__ pos_g, long long p0, long long nb,
                                                      int* __restrict__ list_s, int* __restrict__ list_g)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nb) return;
    const int c = cls[t];
    if (c == 1) list_s[pos_s[t]] = (int)(p0 + t);
    if (c == 2) list_g[pos_g[t]] = (int)(p0 + t);
}

// ---- pass A (calculate_weights, cic_3D.jl:13-78): one warp per particle, same lane layout as the scatter kernel
template <int KID>
__global__ void __launch_bounds__(256) k_norm3d(s2g_particles P, s2g_geom G, const int* __restrict__ list,
                                                long long n_list, GRec3* __restrict__ recs,
                                                unsigned* __restrict__ npairs_g, int* __restrict__ reroute,
                                                unsigned long long* __restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    unsigned long long mapped = 0, fpx = 0;
    for (;;) {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(&counters[CNT_WORK], 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_list) break;
        const long long p = list[t];
        Rec3 r;
        make_rec3(P, G, p, r);  // known valid
        const int ni = r.hi[0] - r.lo[0] + 1, nj = r.hi[1] - r.lo[1] + 1, nk = r.hi[2] - r.lo[2] + 1;
        const int lw = nk >= 32 ? 5 : (nk <= 1 ? 0 : 32 - __clz(nk - 1));
        const int W = 1 << lw, R = 32 >> lw;
        const int c0 = lane & (W - 1), r0 = lane >> lw;
        GRec3 g;
        const double ctr[3] = {r.x, r.y, r.z};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            g.lo[d] = r.lo[d]; g.hi[d] = r.hi[d];
            g.dlo[d] = overlap_1d(ctr[d], r.h, r.lo[d]);
            g.dhi[d] = overlap_1d(ctr[d], r.h, r.hi[d]);
        }
        const double hinv = r.hinv;
        const double xb = center_dist(r.x, (double)r.lo[0]) * hinv;
        double sw = 0.0;
        int cnt = 0;
        for (int kc = c0; kc < nk; kc += W) {
            const int k = r.lo[2] + kc;
            const double cz = center_dist(r.z, (double)k) * hinv;
            const double dz = (k == r.lo[2]) ? g.dlo[2] : ((k == r.hi[2]) ? g.dhi[2] : 1.0);
            for (int jr = r0; jr < nj; jr += R) {
                const int j = r.lo[1] + jr;
                const double by = center_dist(r.y, (double)j) * hinv;
                const double bc2 = fma(by, by, cz * cz) + 1e-300;
                if (bc2 >= 1.0) continue;
                const double dy = (j == r.lo[1]) ? g.dlo[1] : ((j == r.hi[1]) ? g.dhi[1] : 1.0);
                double col = 0.0;
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
                        const int i = r.lo[0] + ii + q;
                        const double dx = (i == r.lo[0]) ? g.dlo[0] : ((i == r.hi[0]) ? g.dhi[0] : 1.0);
                        col = fma(select_or_zero(in[q], wk[q]), dx, col);
                        cnt += in[q] ? 1 : 0;
                    }
                }
                sw = fma(col, dy * dz, sw);
            }
        }
        sw = warp_sum(sw);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        const double n_distr = (double)cnt;
        const double kernel_norm = r.vol / n_distr;                                 // cic_3D.jl:168
        const double vn = kernel_norm * (n_distr / sw) * r.w * G.len2pix;           // :169
        g.p = (int)p; g.pad = 0;
        g.x = r.x; g.y = r.y; g.z = r.z; g.hinv = hinv; g.h = r.h;
        g.vn = vn; g.vq = vn * r.q;
        unsigned np = 0;
        if (sw == 0.0 || !isfinite(vn)) {
            // "no cell centre covered" branch (cic_3D.jl:57-72) or an Inf/NaN normalisation -> scatter kernel
            g.lo[0] = 1; g.hi[0] = 0;
            if (lane == 0) reroute[atomicAdd(&counters[CNT_PAIRS], 1ull)] = (int)p;
        } else {
            const int t0[3] = {r.lo[0] / T_I, r.lo[1] / T_J, r.lo[2] / T_K};
            const int nt[3] = {r.hi[0] / T_I - t0[0] + 1, r.hi[1] / T_J - t0[1] + 1, r.hi[2] / T_K - t0[2] + 1};
            for (int q = lane; q < nt[0] * nt[1] * nt[2]; q += 32) {
                const int tk = q % nt[2], tj = (q / nt[2]) % nt[1], ti = q / (nt[2] * nt[1]);
                if (tile_hit3(g, t0[0] + ti, t0[1] + tj, t0[2] + tk)) ++np;
            }
            np = __reduce_add_sync(0xffffffffu, np);
            if (lane == 0) {
                ++mapped;
                fpx += (unsigned long long)ni * (unsigned long long)nj * (unsigned long long)nk;
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

__global__ void __launch_bounds__(256) k_expand3d(const GRec3* __restrict__ recs, const unsigned* __restrict__ off,
                                                  long long n_list, int ntj, int ntk, unsigned* __restrict__ keys,
                                                  unsigned* __restrict__ vals)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_list) return;
    const GRec3 g = recs[t];
    if (g.lo[0] > g.hi[0]) return;
    unsigned o = off[t];
    for (int ti = g.lo[0] / T_I; ti <= g.hi[0] / T_I; ++ti)
        for (int tj = g.lo[1] / T_J; tj <= g.hi[1] / T_J; ++tj)
            for (int tk = g.lo[2] / T_K; tk <= g.hi[2] / T_K; ++tk)
                if (tile_hit3(g, ti, tj, tk)) {
                    keys[o] = (unsigned)((ti * ntj + tj) * ntk + tk);
                    vals[o] = (unsigned)t;
                    ++o;
                }
}

__global__ void __launch_bounds__(256) k_tile_bounds3(const unsigned* __restrict__ keys, long long m,
                                                      unsigned* __restrict__ tile_beg, unsigned* __restrict__ tile_end)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= m) return;
    const unsigned k = keys[t];
    if (t == 0 || keys[t - 1] != k) tile_beg[k] = (unsigned)t;
    if (t == m - 1 || keys[t + 1] != k) tile_end[k] = (unsigned)(t + 1);
}

__global__ void __launch_bounds__(256) k_tile_chunks3(const unsigned* __restrict__ tile_beg,
                                                      const unsigned* __restrict__ tile_end, int ntiles,
                                                      unsigned* __restrict__ nchunks)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    nchunks[t] = (tile_end[t] - tile_beg[t] + CHUNK - 1) / CHUNK;
}

// ---- pass B (cic_3D.jl:172-188) without atomics
template <int KID>
__global__ void __launch_bounds__(256, 3) k_gather3d(const GRec3* __restrict__ recs, const unsigned* __restrict__ vals,
                                                     const unsigned* __restrict__ tile_beg,
                                                     const unsigned* __restrict__ tile_end,
                                                     const unsigned* __restrict__ chunk_begin, int ntiles, int ntj,
                                                     int ntk, unsigned total_chunks, long long npix,
                                                     double* __restrict__ image,
                                                     unsigned long long* __restrict__ counters)
{
    __shared__ GRec3 s_rec[BATCH];
    __shared__ unsigned s_work[3];
    const int tid = threadIdx.x, lane = tid & 31;
    const int kl = tid & (T_K - 1), jl = tid >> 4;  // a warp covers 2 j-rows x 16 k
    unsigned touched = 0;
    for (;;) {
        if (tid == 0) {
            const unsigned w = (unsigned)atomicAdd(&counters[CNT_WORK], 1ull);
            unsigned tile = 0xffffffffu, b = 0, e = 0;
            if (w < total_chunks) {
                int lo = 0, hi = ntiles;
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
        const int tk = (int)(tile % ntk), tj = (int)((tile / ntk) % ntj), ti = (int)(tile / (ntk * ntj));
        const int i0 = ti * T_I, j = tj * T_J + jl, k = tk * T_K + kl;
        const int jw0 = tj * T_J + (jl & ~1);  // first j-row of this warp (uniform)
        const double jd = (double)j, kd = (double)k, id0 = (double)i0;

        double acc_w[T_I], acc_q[T_I];
#pragma unroll
        for (int r = 0; r < T_I; ++r) { acc_w[r] = 0.0; acc_q[r] = 0.0; }

        for (unsigned b = wb; b < we; b += BATCH) {
            const int nb = (int)min((unsigned)BATCH, we - b);
            __syncthreads();
            if (tid < nb) s_rec[tid] = recs[vals[b + tid]];
            __syncthreads();
            for (int e = 0; e < nb; ++e) {
                const GRec3& g = s_rec[e];
                // warp-uniform integer culls: the warp's 2 j-rows and the tile's i range against the footprint box
                const int rlo = max(g.lo[0], i0), rhi = min(g.hi[0], i0 + T_I - 1);
                if (rlo > rhi || g.hi[1] < jw0 || g.lo[1] > jw0 + 1) continue;
                const double hinv = g.hinv;
                const double by = center_dist(g.y, jd) * hinv, cz = center_dist(g.z, kd) * hinv;
                const double bc2 = fma(by, by, fma(cz, cz, 1e-300));
                const double dy = (j == g.lo[1]) ? g.dlo[1] : ((j == g.hi[1]) ? g.dhi[1] : 1.0);
                const double dz = (k == g.lo[2]) ? g.dlo[2] : ((k == g.hi[2]) ? g.dhi[2] : 1.0);
                const double wv = dy * dz * g.vn;
                const bool live = (j >= g.lo[1]) && (j <= g.hi[1]) && (k >= g.lo[2]) && (k <= g.hi[2]) && below_one(bc2) &&
                                  nonzero_bits(wv);
                if (!__any_sync(0xffffffffu, live)) continue;
                const double wq = dy * dz * g.vq;
                const double xb = center_dist(g.x, id0) * hinv;
#pragma unroll
                for (int r4 = 0; r4 < T_I; r4 += 4) {
                    if (i0 + r4 > rhi || i0 + r4 + 3 < rlo) continue;  // uniform
                    double s[4];
                    bool in[4];
                    bool any_in = false;
                    const bool interior = (i0 + r4 > g.lo[0]) && (i0 + r4 + 3 < g.hi[0]);  // uniform
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = i0 + r4 + q;
                        const double a = fma(-(double)(r4 + q), hinv, xb);
                        s[q] = fma(a, a, bc2);
                        in[q] = live && below_one(s[q]) && (interior || (i >= g.lo[0] && i <= g.hi[0]));
                        any_in = any_in || in[q];
                    }
                    if (!__any_sync(0xffffffffu, any_in)) continue;
                    if (interior) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const double wk = select_or_zero(in[q], shape_s<KID>(s[q]));
                            acc_w[r4 + q] = fma(wk, wv, acc_w[r4 + q]);
                            acc_q[r4 + q] = fma(wk, wq, acc_q[r4 + q]);
                            touched += in[q] ? 1u : 0u;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int i = i0 + r4 + q;
                            const double dx = (i == g.lo[0]) ? g.dlo[0] : ((i == g.hi[0]) ? g.dhi[0] : 1.0);
                            const double wk = select_or_zero(in[q], shape_s<KID>(s[q]) * dx);
                            acc_w[r4 + q] = fma(wk, wv, acc_w[r4 + q]);
                            acc_q[r4 + q] = fma(wk, wq, acc_q[r4 + q]);
                            touched += (in[q] && nonzero_bits(dx)) ? 1u : 0u;
                        }
                    }
                }
            }
        }
        // flush: half-warps write 16 consecutive doubles along k
        const long long npl = npix * npix * npix;
        if (j < npix && k < npix) {
#pragma unroll
            for (int r = 0; r < T_I; ++r) {
                const int i = i0 + r;
                if (i < npix && (acc_w[r] != 0.0 || acc_q[r] != 0.0)) {
                    const long long idx = ((long long)i * npix + j) * npix + k;  // indices.jl:15-17
                    red_add(image + npl + idx, acc_w[r]);
                    red_add(image + idx, acc_q[r]);
                }
            }
        }
        if (touched > 0x7f000000u) {
            atomicAdd(&counters[CNT_TOUCHED], (unsigned long long)touched);
            touched = 0;
        }
        __syncthreads();
    }
    unsigned long long tt = (unsigned long long)warp_sum_ll((long long)touched);
    if (lane == 0 && tt) atomicAdd(&counters[CNT_TOUCHED], tt);
}

struct IsClass3 {
    int c;
    __host__ __device__ unsigned operator()(int v) const { return (unsigned)(v == c); }
};
struct ToU64_3 {
    __host__ __device__ unsigned long long operator()(unsigned v) const { return (unsigned long long)v; }
};

long long env_ll3(const char* name, long long dflt)
{
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    return atoll(s);
}

template <int KID>
int deposit3d_k(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, double* image)
{
    const int kernel = KID;
    // measured cross-over (profiles/r1_3d_strategies.txt): at ~2.7e3 cells/particle (C3) scatter wins by 8 %, at
    // ~1.4e5 cells/particle gather wins by 1.44x (pass A, not the deposit, dominates it)
    const long long gather_min = env_ll3("S2G_GATHER3D_MIN_CELLS", 16384);
    const long long batch_max = env_ll3("S2G_BATCH_PARTICLES", 8LL << 20);
    const long long pair_cap = env_ll3("S2G_PAIR_CAP", 512LL << 20);
    const int nti = (int)((G.npix + T_I - 1) / T_I), ntj = (int)((G.npix + T_J - 1) / T_J),
              ntk = (int)((G.npix + T_K - 1) / T_K);
    const long long ntiles_ll = (long long)nti * ntj * ntk;
    if (ctx->strategy == S2G_STRATEGY_SCATTER || ntiles_ll > (1LL << 30)) {
        S2G_TRY(s2g_stage_wait(ctx, P.n));
        return s2g_launch_scatter_3d(ctx, P, G, kernel, nullptr, P.n, image);  // times itself (PH_DEPOSIT)
    }
    const int ntiles = (int)ntiles_ll;
    cudaStream_t st = ctx->stream;
    long long p0 = 0;
    long long batch = std::min(batch_max, (long long)P.n);
    while (p0 < P.n) {
        long long nb = std::min(batch, (long long)P.n - p0);
        // overlapped staging (s2g_api.cu): small first slice, every slice waits for exactly the particles it reads
        if (p0 == 0 && s2g_stage_first_slice(ctx) > 0) nb = std::min(nb, s2g_stage_first_slice(ctx));
        S2G_TRY(s2g_stage_wait(ctx, p0 + nb));
        void *d_cls, *d_np, *d_ps, *d_pg, *d_ls, *d_lg, *d_tmp, *d_sum;
        S2G_TRY(s2g_scratch(ctx, "g_cls", sizeof(int) * (nb + 1), &d_cls));
        S2G_TRY(s2g_scratch(ctx, "g_np", sizeof(unsigned) * (nb + 1), &d_np));
        S2G_TRY(s2g_scratch(ctx, "g_pos_s", sizeof(unsigned) * (nb + 1), &d_ps));
        S2G_TRY(s2g_scratch(ctx, "g_pos_g", sizeof(unsigned) * (nb + 1), &d_pg));
        S2G_TRY(s2g_scratch(ctx, "g_list_s", sizeof(int) * nb, &d_ls));
        S2G_TRY(s2g_scratch(ctx, "g_list_g", sizeof(int) * nb, &d_lg));
        S2G_TRY(s2g_scratch(ctx, "g_sum", sizeof(unsigned long long), &d_sum));
        const int blocks = (int)((nb + 255) / 256);
        int ph = s2g_phase_begin(ctx, PH_PREP);
        S2G_CUDA(cudaMemsetAsync((int*)d_cls + nb, 0, sizeof(int), st));
        k_classify3d<<<blocks, 256, 0, st>>>(P, G, p0, nb, gather_min, ctx->strategy, (int*)d_cls, (unsigned*)d_np);
        S2G_CUDA(cudaGetLastError());
        cub::TransformInputIterator<unsigned, IsClass3, const int*> it_s((const int*)d_cls, IsClass3{1});
        cub::TransformInputIterator<unsigned, IsClass3, const int*> it_g((const int*)d_cls, IsClass3{2});
        cub::TransformInputIterator<unsigned long long, ToU64_3, const unsigned*> it_np((const unsigned*)d_np, ToU64_3{});
        size_t t1 = 0, t2 = 0, t3 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, t1, it_s, (unsigned*)d_ps, (int)(nb + 1), st);
        cub::DeviceScan::ExclusiveSum(nullptr, t2, it_g, (unsigned*)d_pg, (int)(nb + 1), st);
        cub::DeviceReduce::Sum(nullptr, t3, it_np, (unsigned long long*)d_sum, (int)nb, st);
        const size_t tmp_bytes = std::max(t1, std::max(t2, t3)) + 16;
        S2G_TRY(s2g_scratch(ctx, "g_tmp", tmp_bytes, &d_tmp));
        size_t tb = tmp_bytes;
        S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, it_s, (unsigned*)d_ps, (int)(nb + 1), st));
        tb = tmp_bytes;
        S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, it_g, (unsigned*)d_pg, (int)(nb + 1), st));
        tb = tmp_bytes;
        S2G_CUDA(cub::DeviceReduce::Sum(d_tmp, tb, it_np, (unsigned long long*)d_sum, (int)nb, st));
        unsigned h_ns = 0, h_ng = 0;
        unsigned long long h_ub = 0;
        s2g_readback rb(ctx);
        S2G_CUDA(rb.add(&h_ns, (unsigned*)d_ps + nb, sizeof(unsigned)));
        S2G_CUDA(rb.add(&h_ng, (unsigned*)d_pg + nb, sizeof(unsigned)));
        S2G_CUDA(rb.add(&h_ub, d_sum, sizeof(unsigned long long)));
        s2g_phase_end(ctx, ph);
        ctx->launches += 4;
        S2G_CUDA(rb.sync());
        if ((long long)h_ub > pair_cap && nb > 1024) {
            batch = std::max<long long>(1024, nb / 2);
            continue;
        }
        S2G_CHECK(h_ub < 0xfff00000ull, S2G_ENOMEM,
                  "a single slice of %lld particles spans %llu grid tiles: footprints too large for this grid", nb, h_ub);
        const long long n_s = h_ns, n_g = h_ng;
        k_build_lists3<<<blocks, 256, 0, st>>>((const int*)d_cls, (const unsigned*)d_ps, (const unsigned*)d_pg, p0, nb,
                                               (int*)d_ls, (int*)d_lg);
        S2G_CUDA(cudaGetLastError());
        ctx->launches += 1;
        if (n_s > 0) S2G_TRY(s2g_launch_scatter_3d(ctx, P, G, kernel, (const int*)d_ls, n_s, image));
        if (n_g > 0) {
            void *d_recs, *d_npg, *d_off, *d_rr;
            S2G_TRY(s2g_scratch(ctx, "g_reroute", sizeof(int) * n_g, &d_rr));
            S2G_TRY(s2g_scratch(ctx, "g_recs3", sizeof(GRec3) * n_g, &d_recs));
            S2G_TRY(s2g_scratch(ctx, "g_npg", sizeof(unsigned) * (n_g + 1), &d_npg));
            S2G_TRY(s2g_scratch(ctx, "g_off", sizeof(unsigned) * (n_g + 1), &d_off));
            S2G_CUDA(cudaMemsetAsync((unsigned*)d_npg + n_g, 0, sizeof(unsigned), st));
            S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), st));
            S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_PAIRS, 0, sizeof(unsigned long long), st));
            ph = s2g_phase_begin(ctx, PH_NORM);
            {
                const int nblk = (int)std::min<long long>((n_g + 7) / 8, (long long)ctx->sm_count * 8);
                k_norm3d<KID><<<std::max(nblk, 1), 256, 0, st>>>(P, G, (const int*)d_lg, n_g, (GRec3*)d_recs,
                                                                (unsigned*)d_npg, (int*)d_rr, ctx->d_counters);
                S2G_CUDA(cudaGetLastError());
            }
            s2g_phase_end(ctx, ph);
            ctx->launches += 1;
            ph = s2g_phase_begin(ctx, PH_SORT);
            size_t tb4 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb4, (const unsigned*)d_npg, (unsigned*)d_off, (int)(n_g + 1), st);
            S2G_TRY(s2g_scratch(ctx, "g_tmp", tb4 + 16, &d_tmp));
            S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb4, (const unsigned*)d_npg, (unsigned*)d_off, (int)(n_g + 1), st));
            unsigned h_m = 0;
            unsigned long long h_rr = 0;
            S2G_CUDA(rb.add(&h_m, (unsigned*)d_off + n_g, sizeof(unsigned)));
            S2G_CUDA(rb.add(&h_rr, ctx->d_counters + CNT_PAIRS, sizeof(unsigned long long)));
            S2G_CUDA(rb.sync());
            const long long m = h_m;
            if (h_rr > 0) {
                s2g_phase_end(ctx, ph);
                S2G_TRY(s2g_launch_scatter_3d(ctx, P, G, kernel, (const int*)d_rr, (long long)h_rr, image));
                ph = s2g_phase_begin(ctx, PH_SORT);
            }
            if (m > 0) {
                void *d_keys, *d_vals, *d_keys2, *d_vals2, *d_tend, *d_tbeg, *d_nch, *d_cbeg;
                S2G_TRY(s2g_scratch(ctx, "g_keys", sizeof(unsigned) * m, &d_keys));
                S2G_TRY(s2g_scratch(ctx, "g_vals", sizeof(unsigned) * m, &d_vals));
                S2G_TRY(s2g_scratch(ctx, "g_keys2", sizeof(unsigned) * m, &d_keys2));
                S2G_TRY(s2g_scratch(ctx, "g_vals2", sizeof(unsigned) * m, &d_vals2));
                S2G_TRY(s2g_scratch(ctx, "g_tcnt", sizeof(unsigned) * (ntiles + 1), &d_tend));
                S2G_TRY(s2g_scratch(ctx, "g_tbeg", sizeof(unsigned) * (ntiles + 1), &d_tbeg));
                S2G_TRY(s2g_scratch(ctx, "g_nch", sizeof(unsigned) * (ntiles + 1), &d_nch));
                S2G_TRY(s2g_scratch(ctx, "g_cbeg", sizeof(unsigned) * (ntiles + 1), &d_cbeg));
                k_expand3d<<<(int)((n_g + 255) / 256), 256, 0, st>>>((const GRec3*)d_recs, (const unsigned*)d_off, n_g, ntj,
                                                                     ntk, (unsigned*)d_keys, (unsigned*)d_vals);
                S2G_CUDA(cudaGetLastError());
                int bits = 1;
                while ((1LL << bits) < (long long)ntiles) ++bits;
                size_t sb = 0;
                cub::DeviceRadixSort::SortPairs(nullptr, sb, (const unsigned*)d_keys, (unsigned*)d_keys2,
                                                (const unsigned*)d_vals, (unsigned*)d_vals2, (int)m, 0, bits, st);
                S2G_TRY(s2g_scratch(ctx, "g_sort_tmp", sb + 16, &d_tmp));
                S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, (const unsigned*)d_keys, (unsigned*)d_keys2,
                                                         (const unsigned*)d_vals, (unsigned*)d_vals2, (int)m, 0, bits, st));
                S2G_CUDA(cudaMemsetAsync(d_tend, 0, sizeof(unsigned) * (ntiles + 1), st));
                S2G_CUDA(cudaMemsetAsync(d_tbeg, 0, sizeof(unsigned) * (ntiles + 1), st));
                k_tile_bounds3<<<(int)((m + 255) / 256), 256, 0, st>>>((const unsigned*)d_keys2, m, (unsigned*)d_tbeg,
                                                                       (unsigned*)d_tend);
                S2G_CUDA(cudaGetLastError());
                S2G_CUDA(cudaMemsetAsync(d_nch, 0, sizeof(unsigned) * (ntiles + 1), st));
                k_tile_chunks3<<<(ntiles + 255) / 256, 256, 0, st>>>((const unsigned*)d_tbeg, (const unsigned*)d_tend,
                                                                     ntiles, (unsigned*)d_nch);
                S2G_CUDA(cudaGetLastError());
                size_t tb3 = 0;
                cub::DeviceScan::ExclusiveSum(nullptr, tb3, (const unsigned*)d_nch, (unsigned*)d_cbeg, ntiles + 1, st);
                S2G_TRY(s2g_scratch(ctx, "g_tmp", tb3 + 16, &d_tmp));
                size_t tbb = tb3 + 16;
                S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tbb, (const unsigned*)d_nch, (unsigned*)d_cbeg, ntiles + 1, st));
                unsigned h_chunks = 0;
                S2G_CUDA(rb.add(&h_chunks, (unsigned*)d_cbeg + ntiles, sizeof(unsigned)));
                s2g_phase_end(ctx, ph);
                ph = -1;
                ctx->launches += 8;
                S2G_CUDA(rb.sync());
                S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), st));
                const int phg = s2g_phase_begin(ctx, PH_DEPOSIT);
                const int gblocks = (int)std::min<long long>((long long)h_chunks, (long long)ctx->sm_count * 3);
                k_gather3d<KID><<<std::max(gblocks, 1), 256, 0, st>>>((const GRec3*)d_recs, (const unsigned*)d_vals2,
                                                                     (const unsigned*)d_tbeg, (const unsigned*)d_tend,
                                                                     (const unsigned*)d_cbeg, ntiles, ntj, ntk, h_chunks,
                                                                     G.npix, image, ctx->d_counters);
                S2G_CUDA(cudaGetLastError());
                s2g_phase_end(ctx, phg);
                ctx->launches += 1;
                ctx->host_pairs += m;
            }
            if (ph >= 0) s2g_phase_end(ctx, ph);
        }
        p0 += nb;
    }
    return S2G_OK;
}

}  // namespace

int s2g_launch_deposit_3d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, double* image)
{
    if (P.n <= 0) return S2G_OK;
    switch (kernel) {
    case S2G_KERNEL_CUBIC: return deposit3d_k<S2G_KERNEL_CUBIC>(ctx, P, G, image);
    case S2G_KERNEL_QUINTIC: return deposit3d_k<S2G_KERNEL_QUINTIC>(ctx, P, G, image);
    case S2G_KERNEL_WENDLAND_C2: return deposit3d_k<S2G_KERNEL_WENDLAND_C2>(ctx, P, G, image);
    case S2G_KERNEL_WENDLAND_C4: return deposit3d_k<S2G_KERNEL_WENDLAND_C4>(ctx, P, G, image);
    case S2G_KERNEL_WENDLAND_C6: return deposit3d_k<S2G_KERNEL_WENDLAND_C6>(ctx, P, G, image);
    case S2G_KERNEL_WENDLAND_C8: return deposit3d_k<S2G_KERNEL_WENDLAND_C8>(ctx, P, G, image);
    }
    s2g_set_error("unknown kernel id %d", kernel);
    return S2G_EINVAL;
}
