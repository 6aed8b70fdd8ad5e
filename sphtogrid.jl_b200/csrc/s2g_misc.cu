// s2g_misc.cu — CIC/TSC stencils, finite-guarded accumulate, synthetic particle stream, roofline microbenchmarks.
#include <cub/cub.cuh>
#include <cstdlib>

#include "s2g_common.cuh"

// ------------------------------------------------------------------------------------------------
// CIC / TSC stencil deposit (semantics in DESIGN.md §stencils; the reference holds only commented-out code,
// src/tsc_interpolation/tsc_interpolation.jl:1-183).  One thread per particle, 2^d / 3^d red.add.f64 into the
// [field, weight] planes.  Purely HBM/L2-atomic bound: 3 coordinates + 1 value in, 2*order^d reds out.
// ------------------------------------------------------------------------------------------------
template <int ORDER, int DIMS>
__global__ void __launch_bounds__(256) k_stencil(const void* __restrict__ pos, const void* __restrict__ q, long long n,
                                                 int in_dtype, double len2pix, double half_n, long long npix,
                                                 int periodic, double* __restrict__ image,
                                                 const unsigned* __restrict__ order)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long p = order ? (long long)order[t] : t;
    const long long ncell = DIMS == 2 ? npix * npix : npix * npix * npix;
    double wgt[3][3];
    long long c0[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double g = __dadd_rn(__dmul_rn(ld_in(pos, 3 * p + d, in_dtype), len2pix), half_n);
        if (ORDER == 2) {
            const double s = __dadd_rn(g, -0.5);
            const double f = floor(s);
            const double fr = __dadd_rn(s, -f);
            c0[d] = (long long)f;
            wgt[d][0] = __dadd_rn(1.0, -fr);
            wgt[d][1] = fr;
            wgt[d][2] = 0.0;
        } else {
            const double f = floor(g);
            const double dd = __dadd_rn(g, -__dadd_rn(f, 0.5));
            const double a = __dadd_rn(0.5, -dd), b = __dadd_rn(0.5, dd);
            c0[d] = (long long)f - 1;
            wgt[d][0] = __dmul_rn(0.5, __dmul_rn(a, a));
            wgt[d][1] = __dadd_rn(0.75, -__dmul_rn(dd, dd));
            wgt[d][2] = __dmul_rn(0.5, __dmul_rn(b, b));
        }
    }
    const double qv = ld_in(q, p, in_dtype);
#pragma unroll
    for (int a = 0; a < ORDER; ++a) {
        long long i = c0[0] + a;
        if (i < 0 || i >= npix) {
            if (!periodic) continue;
            i = ((i % npix) + npix) % npix;
        }
#pragma unroll
        for (int b = 0; b < ORDER; ++b) {
            long long j = c0[1] + b;
            if (j < 0 || j >= npix) {
                if (!periodic) continue;
                j = ((j % npix) + npix) % npix;
            }
            const double wab = __dmul_rn(wgt[0][a], wgt[1][b]);
            if (DIMS == 2) {
                const long long idx = i * npix + j;
                red_add(image + idx, __dmul_rn(qv, wab));
                red_add(image + ncell + idx, wab);
            } else {
#pragma unroll
                for (int c = 0; c < ORDER; ++c) {
                    long long k = c0[2] + c;
                    if (k < 0 || k >= npix) {
                        if (!periodic) continue;
                        k = ((k % npix) + npix) % npix;
                    }
                    const double ww = __dmul_rn(wab, wgt[2][c]);
                    const long long idx = i * npix * npix + j * npix + k;
                    red_add(image + idx, __dmul_rn(qv, ww));
                    red_add(image + ncell + idx, ww);
                }
            }
        }
    }
}

// Deposit order: by the block (16^3 cells, 64^2 pixels) of the particle position.  A thread's reds go to its own cells —
// in input order every one of them misses L2 once the grid is larger than L2 (a DRAM sector read and written per
// 2-3-cell row); in block order the threads in flight update one compact region.  S2G_STENCIL_ORDER=0: off.
__global__ void __launch_bounds__(256) k_stencil_keys(const void* __restrict__ pos, long long n, int in_dtype, int dims,
                                                      double len2pix, double half_n, int nb, double inv_block,
                                                      unsigned* __restrict__ keys, unsigned* __restrict__ idx)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= n) return;
    unsigned key = 0;
    for (int d = 0; d < dims; ++d) {
        const double g = fma(ld_in(pos, 3 * p + d, in_dtype), len2pix, half_n);
        int b = (g == g) ? (int)floor(fmin(fmax(g * inv_block, -1.0), (double)nb)) : 0;
        b = min(max(b, 0), nb - 1);
        key = key * (unsigned)nb + (unsigned)b;
    }
    keys[p] = key;
    idx[p] = (unsigned)p;
}

int s2g_launch_stencil(s2g_ctx* ctx, int order, int dims, const void* pos, const void* q, long long n, int in_dtype,
                       double len2pix, long long npix, int periodic, double* image)
{
    if (n <= 0) return S2G_OK;
    const int blocks = (int)((n + 255) / 256);
    const double half_n = 0.5 * (double)npix;
    const unsigned* ord = nullptr;
    const char* e_ord = getenv("S2G_STENCIL_ORDER");
    const long long grid_bytes = 16LL * (dims == 2 ? npix * npix : npix * npix * npix);
    if (!(e_ord && atoi(e_ord) == 0) && n >= 65536 && n < (1LL << 31) && grid_bytes > (48LL << 20)) {
        const int bs = dims == 2 ? 64 : 16;
        const int nb = (int)((npix + bs - 1) / bs);
        void *d_k, *d_k2, *d_i, *d_i2, *d_tmp;
        S2G_TRY(s2g_scratch(ctx, "st_keys", sizeof(unsigned) * n, &d_k));
        S2G_TRY(s2g_scratch(ctx, "st_keys2", sizeof(unsigned) * n, &d_k2));
        S2G_TRY(s2g_scratch(ctx, "st_idx", sizeof(unsigned) * n, &d_i));
        S2G_TRY(s2g_scratch(ctx, "st_idx2", sizeof(unsigned) * n, &d_i2));
        k_stencil_keys<<<blocks, 256, 0, ctx->stream>>>(pos, n, in_dtype, dims, len2pix, half_n, nb, 1.0 / bs,
                                                        (unsigned*)d_k, (unsigned*)d_i);
        S2G_CUDA(cudaGetLastError());
        long long nkeys = 1;
        for (int d = 0; d < dims; ++d) nkeys *= nb;
        int bits = 1;
        while ((1LL << bits) < nkeys) ++bits;
        if (bits <= 32) {
            size_t sb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, sb, (const unsigned*)d_k, (unsigned*)d_k2, (const unsigned*)d_i,
                                            (unsigned*)d_i2, (int)n, 0, bits, ctx->stream);
            S2G_TRY(s2g_scratch(ctx, "st_sort_tmp", sb + 16, &d_tmp));
            S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, (const unsigned*)d_k, (unsigned*)d_k2,
                                                     (const unsigned*)d_i, (unsigned*)d_i2, (int)n, 0, bits, ctx->stream));
            ctx->launches += 4;
            ord = (const unsigned*)d_i2;
        }
    }
    if (order == 2 && dims == 2)
        k_stencil<2, 2><<<blocks, 256, 0, ctx->stream>>>(pos, q, n, in_dtype, len2pix, half_n, npix, periodic, image, ord);
    else if (order == 2 && dims == 3)
        k_stencil<2, 3><<<blocks, 256, 0, ctx->stream>>>(pos, q, n, in_dtype, len2pix, half_n, npix, periodic, image, ord);
    else if (order == 3 && dims == 2)
        k_stencil<3, 2><<<blocks, 256, 0, ctx->stream>>>(pos, q, n, in_dtype, len2pix, half_n, npix, periodic, image, ord);
    else
        k_stencil<3, 3><<<blocks, 256, 0, ctx->stream>>>(pos, q, n, in_dtype, len2pix, half_n, npix, periodic, image, ord);
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// sum += local where local is finite  (src/distributed_mapping/cic.jl:62-70, healpix.jl:44-52)
// ------------------------------------------------------------------------------------------------
__global__ void k_accumulate_finite(double* __restrict__ sum, const double* __restrict__ local, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += stride) {
        const double v = local[e];
        if (!isnan(v) && !isinf(v)) sum[e] += v;
    }
}

int s2g_launch_accumulate_finite(s2g_ctx* ctx, double* sum, const double* local, long long n)
{
    if (n <= 0) return S2G_OK;
    const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    k_accumulate_finite<<<blocks, 256, 0, ctx->stream>>>(sum, local, n);
    S2G_CUDA(cudaGetLastError());
    return S2G_OK;
}

// reduce_image division on a pixel SLICE (the per-rank epilogue after a reduce-scatter of the partial images):
// dims 2 (reduce_image.jl:8-31): q /= w where reduce_image and w > 0;  dims 3 (reduce_image.jl:39-55, quirk Q7): where
// q > 0, q /= (reduce_image ? w : 1).  In place, elementwise, n_images quantity planes of `stride` elements each.
__global__ void k_divide_slice(double* __restrict__ q, const double* __restrict__ w, long long n, long long stride,
                               int n_images, int dims, int reduce_image)
{
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += step) {
        const double wv = w[e];
        for (int k = 0; k < n_images; ++k) {
            double v = q[k * stride + e];
            if (dims == 2) {
                if (reduce_image && wv > 0.0) v = v / wv;
            } else if (v > 0.0)
                v = v / (reduce_image ? wv : 1.0);
            q[k * stride + e] = v;
        }
    }
}

int s2g_launch_divide_slice(s2g_ctx* ctx, int dims, double* q, const double* w, long long n, long long stride,
                            int n_images, int reduce_image)
{
    if (n <= 0) return S2G_OK;
    const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    const int ph = s2g_phase_begin(ctx, PH_EPILOGUE);
    k_divide_slice<<<blocks, 256, 0, ctx->stream>>>(q, w, n, stride, n_images, dims, reduce_image);
    s2g_phase_end(ctx, ph);
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// synthetic Gadget-like particles: Philox4x32-10, key = seed, counter = (particle id, draw)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo)
{
    const unsigned long long v = (((unsigned long long)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);  // [0,1)
}

template <typename T>
__global__ void __launch_bounds__(256) k_synth(unsigned long long seed, long long first_id, long long n,
                                               long long n_total, double box, double n_ngb, double sigma,
                                               T* __restrict__ pos, T* __restrict__ hsml, T* __restrict__ m,
                                               T* __restrict__ rho, T* __restrict__ temp)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned long long id = (unsigned long long)(first_id + t);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t r0[4], r1[4], r2[4];
    philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), 0u, 0u, k0, k1, r0);
    philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), 1u, 0u, k0, k1, r1);
    philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), 2u, 0u, k0, k1, r2);
    const double x = u53(r0[0], r0[1]) * box, y = u53(r0[2], r0[3]) * box, z = u53(r1[0], r1[1]) * box;
    // Box-Muller
    const double ua = (u53(r1[2], r1[3]) + 1.0 / 9007199254740992.0), ub = u53(r2[0], r2[1]);
    const double rad = sqrt(-2.0 * log(ua));
    const double g1 = rad * cospi(2.0 * ub), g2 = rad * sinpi(2.0 * ub);
    const double mass = 1.0 / (double)n_total;
    const double rho_bar = 1.0 / (box * box * box);
    const double rr = rho_bar * exp(sigma * g1 - 0.5 * sigma * sigma);
    const double h = cbrt(3.0 * n_ngb * mass / (4.0 * 3.14159265358979323846 * rr));
    const double tt = 1.0e4 * pow(rr / rho_bar, 2.0 / 3.0) * exp(0.5 * g2);
    pos[3 * t + 0] = (T)x; pos[3 * t + 1] = (T)y; pos[3 * t + 2] = (T)z;
    hsml[t] = (T)h; m[t] = (T)mass; rho[t] = (T)rr; temp[t] = (T)tt;
}

int s2g_launch_synth(s2g_ctx* ctx, uint64_t seed, long long first_id, long long n, long long n_total, double box,
                     double n_ngb, double sigma, int out_dtype, void* pos, void* hsml, void* m, void* rho, void* temp)
{
    if (n <= 0) return S2G_OK;
    const int blocks = (int)((n + 255) / 256);
    if (out_dtype == S2G_F64)
        k_synth<double><<<blocks, 256, 0, ctx->stream>>>(seed, first_id, n, n_total, box, n_ngb, sigma, (double*)pos,
                                                         (double*)hsml, (double*)m, (double*)rho, (double*)temp);
    else
        k_synth<float><<<blocks, 256, 0, ctx->stream>>>(seed, first_id, n, n_total, box, n_ngb, sigma, (float*)pos,
                                                        (float*)hsml, (float*)m, (float*)rho, (float*)temp);
    S2G_CUDA(cudaGetLastError());
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// microbenchmarks: the roofline denominators that MEASURED_PEAKS.json does not hold
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mb_dfma(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;
}

// every warp adds to 32 consecutive doubles of a pseudo-randomly chosen 256-byte row (the deposit's access shape)
__global__ void __launch_bounds__(256) k_mb_red_rows(double* buf, unsigned long long rows, int iters)
{
    const unsigned lane = threadIdx.x & 31;
    unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    unsigned long long s = w * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    for (int i = 0; i < iters; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const unsigned long long row = (s >> 20) % rows;
        red_add(buf + row * 32 + lane, 1.0);
    }
}

__global__ void __launch_bounds__(256) k_mb_red_random(double* buf, unsigned long long n, int iters)
{
    unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long s = t * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    for (int i = 0; i < iters; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        red_add(buf + (s >> 20) % n, 1.0);
    }
}

__global__ void __launch_bounds__(256) k_mb_copy(const double4* __restrict__ src, double4* __restrict__ dst,
                                                 unsigned long long n4)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < n4; e += stride)
        dst[e] = src[e];
}

// shared-memory FP64 atomicAdd throughput (spread addresses in a 32 KB tile)
__global__ void __launch_bounds__(256) k_mb_smem_atomic(double* out, int iters)
{
    __shared__ double tile[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tile[i] = 0.0;
    __syncthreads();
    unsigned s = threadIdx.x * 2654435761u + blockIdx.x;
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const unsigned row = (s >> 12) & 127u;
        atomicAdd(&tile[row * 32 + (threadIdx.x & 31)], 1.0);
    }
    __syncthreads();
    if (tile[threadIdx.x] == -1.0) out[0] = 1.0;
}

// TMA bulk reduction: every warp stages NB doubles in shared memory and adds them to a pseudo-randomly chosen
// NB*8-byte row of the buffer with ONE cp.reduce.async.bulk (.add.f64, SASS UBLKRED) issued by lane 0
template <int NB>
__global__ void __launch_bounds__(256) k_mb_bulk_red_rows(double* buf, unsigned long long rows, int iters)
{
    __shared__ __align__(128) double stage[8][2][NB];
    const unsigned lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
    unsigned long long s = w * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    for (int i = 0; i < iters; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const unsigned long long row = (s >> 20) % rows;
        const int slot = i & 1;
        if (i >= 2) {  // the bulk op issued two iterations ago has finished reading this slot
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
        for (int k = lane; k < NB; k += 32) stage[wq][slot][k] = 1.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const unsigned saddr = (unsigned)__cvta_generic_to_shared(&stage[wq][slot][0]);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                         ::"l"(buf + row * NB), "r"(saddr), "r"(NB * 8) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int s2g_run_microbench(s2g_ctx* ctx, int which, size_t bytes, int iters, double* rate_out)
{
    void* buf = nullptr;
    const int blocks = ctx->sm_count * 8;
    cudaEvent_t a = ctx->ev[8], b = ctx->ev[9];
    float ms = 0.f;
    if (which == 0 || which == 4) {
        S2G_TRY(s2g_scratch(ctx, "mb", 4096, &buf));
        for (int rep = 0; rep < 2; ++rep) {  // first is warm-up
            S2G_CUDA(cudaEventRecord(a, ctx->stream));
            if (which == 0)
                k_mb_dfma<<<blocks, 256, 0, ctx->stream>>>((double*)buf, iters);
            else
                k_mb_smem_atomic<<<blocks, 256, 0, ctx->stream>>>((double*)buf, iters);
            S2G_CUDA(cudaEventRecord(b, ctx->stream));
            S2G_CUDA(cudaEventSynchronize(b));
            S2G_CUDA(cudaEventElapsedTime(&ms, a, b));
        }
        const double ops = (double)blocks * 256.0 * (double)iters * (which == 0 ? 16.0 : 1.0);
        *rate_out = ops / (ms * 1e-3) * 1e-9;  // GFLOP/s or Gatomic/s
        return S2G_OK;
    }
    if (which == 1 || which == 2 || which == 5 || which == 6) {
        if (bytes < 4096) bytes = 4096;
        S2G_TRY(s2g_scratch(ctx, "mb", bytes, &buf));
        S2G_CUDA(cudaMemsetAsync(buf, 0, bytes, ctx->stream));
        const unsigned long long nd = bytes / 8;
        for (int rep = 0; rep < 2; ++rep) {
            S2G_CUDA(cudaEventRecord(a, ctx->stream));
            if (which == 1)
                k_mb_red_rows<<<blocks, 256, 0, ctx->stream>>>((double*)buf, nd / 32, iters);
            else if (which == 5)   // 256-byte bulk reductions
                k_mb_bulk_red_rows<32><<<blocks, 256, 0, ctx->stream>>>((double*)buf, nd / 32, iters);
            else if (which == 6)   // 2-KiB bulk reductions
                k_mb_bulk_red_rows<256><<<blocks, 256, 0, ctx->stream>>>((double*)buf, nd / 256, iters);
            else
                k_mb_red_random<<<blocks, 256, 0, ctx->stream>>>((double*)buf, nd, iters);
            S2G_CUDA(cudaEventRecord(b, ctx->stream));
            S2G_CUDA(cudaEventSynchronize(b));
            S2G_CUDA(cudaEventElapsedTime(&ms, a, b));
        }
        *rate_out = (double)blocks * 256.0 * (double)iters * (which == 6 ? 8.0 : 1.0) / (ms * 1e-3) * 1e-9;  // Gadd/s
        return S2G_OK;
    }
    if (which == 3) {
        if (bytes < (1u << 20)) bytes = 1u << 20;
        bytes &= ~(size_t)31;
        S2G_TRY(s2g_scratch(ctx, "mb", 2 * bytes, &buf));
        S2G_CUDA(cudaMemsetAsync(buf, 0, 2 * bytes, ctx->stream));
        float best = 1e30f;
        for (int rep = 0; rep < iters + 1; ++rep) {
            S2G_CUDA(cudaEventRecord(a, ctx->stream));
            k_mb_copy<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>((const double4*)buf,
                                                                     (double4*)((char*)buf + bytes), bytes / 32);
            S2G_CUDA(cudaEventRecord(b, ctx->stream));
            S2G_CUDA(cudaEventSynchronize(b));
            S2G_CUDA(cudaEventElapsedTime(&ms, a, b));
            if (rep > 0 && ms < best) best = ms;
        }
        *rate_out = 2.0 * (double)bytes / (best * 1e-3) * 1e-9;  // GB/s read+write
        return S2G_OK;
    }
    s2g_set_error("s2g_microbench: unknown benchmark %d", which);
    return S2G_EINVAL;
}
