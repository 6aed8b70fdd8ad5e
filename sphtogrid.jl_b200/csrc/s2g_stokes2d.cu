// s2g_stokes2d.cu — ORDERED compositing variant of the 2D Smac deposit: cic_mapping_2D called with a rotation
// measure per particle and stokes=true (src/cic_interpolation/cic_2D.jl:129-131, :201-217 and
// faraday_rotate_pixel!, src/cic_interpolation/cic_shared.jl:129-159).
//
// The reference walks the particles in the order given (far -> near after sphMapping's sort_z,
// cic_interpolation.jl:74-83) and, for every pixel of a particle's bounding box that already holds emission,
// rotates the pixel's (Q,U) = (plane 1, plane 2) by mod(RM_p * pix_weight, pi) before adding the particle's own
// contribution.  That is order dependent per pixel but independent BETWEEN pixels, so the GPU formulation is:
//
//   1. k_stokes_prep  : warp per particle, exact pass A (the same arithmetic as the scatter kernel) -> one record
//                       per particle holding the pixel-space geometry, area_norm, RM and (q1,q2)
//   2. k_stokes_expand: (tile, particle) pairs for every 16x16 tile the bounding box overlaps, written in particle
//                       order; a STABLE radix sort by tile keeps the particle order inside every tile list
//   3. k_stokes2d     : one CTA per tile, one thread per pixel; the thread keeps (Q, U, weight, touched) of its pixel
//                       in registers and replays the tile's particle list front to back.  No atomics: a tile is
//                       owned by exactly one CTA.  State is loaded from / stored to the image and a touched byte map
//                       so that long particle lists can be processed in consecutive slices.
//
// Quirks reproduced literally (faraday_rotate_pixel!, cic_shared.jl:143-155): psi = 0.5*atan(U/Q) is the
// one-argument arctangent, so a touched pixel with Q < 0 flips sign even under a zero rotation, pixels of the
// bounding box outside the kernel support are "rotated" by zero as well, Q = U = 0 gives NaN, and with a single
// mapped quantity plane 2 is the weight plane.
#include "s2g_cic2d.cuh"

#include <cub/cub.cuh>

namespace {

constexpr int ST_T = 16;  // tile edge (pixels); 256 threads = one pixel each

struct __align__(16) SRec {
    double x, y, h, hinv;
    double area_norm, rm;
    double q0, q1;
    int iMin, iMax, jMin, jMax;
    int p;        // particle index (for the planes beyond the first two)
    int flags;    // bit 0: fallback (wk := 1), bit 1: bin_q collapsed to scalar 0.0
    int pad[2];
};
static_assert(sizeof(SRec) == 96, "SRec layout");

// ---- 1. exact pass A (calculate_weights, cic_2D.jl:11-72) and the per-particle record ---------------------------
template <int KID>
__global__ void __launch_bounds__(256) k_stokes_prep(s2g_particles P, s2g_geom G, const double* __restrict__ rm,
                                                     long long p0, long long nb, SRec* __restrict__ recs,
                                                     unsigned* __restrict__ npt,
                                                     unsigned long long* __restrict__ counters)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
    unsigned long long mapped = 0, fpx = 0, fallback = 0;
    for (long long t = warp; t < nb; t += nwarps) {
        const long long p = p0 + t;
        Rec2 r;
        SRec s;
        s.iMin = 0; s.iMax = -1; s.jMin = 0; s.jMax = -1;
        s.p = (int)t; s.flags = 0; s.pad[0] = s.pad[1] = 0;
        s.x = s.y = s.h = s.hinv = s.area_norm = s.rm = s.q0 = s.q1 = 0.0;
        unsigned ntile = 0;
        if (make_rec2(P, G, p, r)) {
            const int ni = r.iMax - r.iMin + 1, nj = r.jMax - r.jMin + 1;
            const int lw = nj >= 32 ? 5 : (nj <= 1 ? 0 : 32 - __clz(nj - 1));
            const int W = 1 << lw, R = 32 >> lw;
            const int c0 = lane & (W - 1), r0 = lane >> lw;
            const double dx_lo = overlap_1d(r.x, r.h, r.iMin), dx_hi = overlap_1d(r.x, r.h, r.iMax);
            const double dy_lo = overlap_1d(r.y, r.h, r.jMin), dy_hi = overlap_1d(r.y, r.h, r.jMax);
            double sw = 0.0, da = 0.0;
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
                    const double dx = (i == r.iMin) ? dx_lo : ((i == r.iMax) ? dx_hi : 1.0);
                    da += dx * dy;
                    if (u <= 1.0) {
                        sw = fma(kernel_shape<KID>(u), dx * dy, sw);
                        ++cnt;
                    }
                }
            }
            sw = warp_sum(sw);
            da = warp_sum(da);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            double n_distr, wpp;
            if (sw == 0.0) {  // cic_2D.jl:51-66
                s.flags |= 1;
                n_distr = (double)ni * (double)nj;
                wpp = (da != 0.0) ? n_distr / da : 1.0;
                ++fallback;
            } else {
                n_distr = (double)cnt;
                wpp = n_distr / sw;
            }
            const double kernel_norm = r.area / n_distr;  // cic_2D.jl:187-188
            s.area_norm = kernel_norm * wpp * r.w * r.dz;
            s.x = r.x; s.y = r.y; s.h = r.h; s.hinv = r.hinv;
            s.iMin = r.iMin; s.iMax = r.iMax; s.jMin = r.jMin; s.jMax = r.jMax;
            s.rm = __ldg(rm + p);
            if (r.all_zero)
                s.flags |= 2;
            else {
                s.q0 = ld_in(P.binq, (long long)G.n_images * p, P.in_dtype);
                if (G.n_images > 1) s.q1 = ld_in(P.binq, (long long)G.n_images * p + 1, P.in_dtype);
            }
            ntile = (unsigned)(r.iMax / ST_T - r.iMin / ST_T + 1) * (unsigned)(r.jMax / ST_T - r.jMin / ST_T + 1);
            ++mapped;
            fpx += (unsigned long long)ni * (unsigned long long)nj;
        }
        if (lane == 0) {
            recs[t] = s;
            npt[t] = ntile;
        }
    }
    if (lane == 0) {
        if (mapped) atomicAdd(&counters[CNT_MAPPED], mapped);
        if (fpx) atomicAdd(&counters[CNT_FOOTPRINT], fpx);
        if (fallback) atomicAdd(&counters[CNT_FALLBACK], fallback);
    }
}

// ---- 2. (tile, particle) pairs in particle order ------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stokes_expand(const SRec* __restrict__ recs, const unsigned* __restrict__ off,
                                                       long long nb, int ntile_j, unsigned* __restrict__ keys,
                                                       unsigned* __restrict__ vals)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nb) return;
    const int iMin = recs[t].iMin, iMax = recs[t].iMax, jMin = recs[t].jMin, jMax = recs[t].jMax;
    if (iMin > iMax || jMin > jMax) return;
    unsigned o = off[t];
    for (int ti = iMin / ST_T; ti <= iMax / ST_T; ++ti)
        for (int tj = jMin / ST_T; tj <= jMax / ST_T; ++tj) {
            keys[o] = (unsigned)(ti * ntile_j + tj);
            vals[o] = (unsigned)t;
            ++o;
        }
}

__global__ void __launch_bounds__(256) k_stokes_bounds(const unsigned* __restrict__ keys, long long m,
                                                       unsigned* __restrict__ tile_beg, unsigned* __restrict__ tile_end)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= m) return;
    const unsigned k = keys[t];
    if (t == 0 || keys[t - 1] != k) tile_beg[k] = (unsigned)t;
    if (t == m - 1 || keys[t + 1] != k) tile_end[k] = (unsigned)(t + 1);
}

struct StU64 {
    __host__ __device__ unsigned long long operator()(unsigned v) const { return v; }
};

// Julia's mod(x::Float64, y::Float64) for y = Float64(pi): rem (exact), then moved into the sign of y
__device__ __forceinline__ double julia_mod_pi(double x)
{
    const double y = 3.141592653589793;
    const double r = fmod(x, y);
    if (r == 0.0) return copysign(r, y);
    if (r < 0.0) return __dadd_rn(r, y);
    return r;
}

// ---- 3. replay of the tile's particle list, one pixel per thread -------------------------------------------------
template <int KID>
__global__ void __launch_bounds__(256) k_stokes2d(const SRec* __restrict__ recs, const unsigned* __restrict__ vals,
                                                  const unsigned* __restrict__ tile_beg,
                                                  const unsigned* __restrict__ tile_end, int ntile_j, s2g_particles P,
                                                  s2g_geom G, long long p0, double* __restrict__ image,
                                                  unsigned char* __restrict__ touched_map,
                                                  unsigned long long* __restrict__ counters)
{
    constexpr int BATCH = 32;
    __shared__ SRec s_rec[BATCH];
    __shared__ unsigned long long s_touched;
    const int tile = blockIdx.x;
    const unsigned beg = tile_beg[tile], end = tile_end[tile];
    if (beg >= end) return;
    if (threadIdx.x == 0) s_touched = 0ull;
    const int i = (tile / ntile_j) * ST_T + (threadIdx.x >> 4);
    const int j = (tile % ntile_j) * ST_T + (threadIdx.x & 15);
    const bool inside = i < (int)G.npix && j < (int)G.npix;
    const long long npl = G.npix * G.npix;
    const long long idx = (long long)i * G.npix + j;
    const int nim = G.n_images;
    // planes 0 and 1 are what faraday_rotate_pixel! calls Q and U; the weight plane is plane nim (== 1 if nim == 1)
    double a0 = 0.0, a1 = 0.0, aw = 0.0;
    bool touched = false;
    if (inside) {
        a0 = image[idx];
        a1 = image[idx + npl];
        if (nim > 1) aw = image[idx + npl * nim];
        touched = touched_map[idx] != 0;
    }
    const double fi = (double)i, fj = (double)j;
    unsigned long long n_touch = 0;

    for (unsigned b = beg; b < end; b += BATCH) {
        const int nrec = min((unsigned)BATCH, end - b);
        __syncthreads();
        // cooperative copy of up to BATCH records (96 B each = 6 x 16 B)
        for (int k = threadIdx.x; k < nrec * 6; k += blockDim.x) {
            const int rIdx = k / 6, part = k % 6;
            reinterpret_cast<int4*>(&s_rec[rIdx])[part] =
                __ldg(reinterpret_cast<const int4*>(recs + vals[b + rIdx]) + part);
        }
        __syncthreads();
        if (!inside) continue;
        for (int k = 0; k < nrec; ++k) {
            const SRec& s = s_rec[k];
            if (i < s.iMin || i > s.iMax || j < s.jMin || j > s.jMax) continue;
            // get_x_dx (cic_shared.jl:68-76): overlap lengths and centre offsets, individually rounded
            const double dx = overlap_1d(s.x, s.h, i), dy = overlap_1d(s.y, s.h, j);
            double wk;
            if (s.flags & 1)
                wk = 1.0;
            else {
                const double xd = center_dist(s.x, fi), yd = center_dist(s.y, fj);
                const double u = u_of(__dmul_rn(xd, xd), __dmul_rn(yd, yd), s.hinv);
                wk = (u <= 1.0) ? kernel_shape<KID>(u) : 0.0;
            }
            const double pw = __dmul_rn(__dmul_rn(wk, __dmul_rn(dx, dy)), s.area_norm);  // cic_2D.jl:199
            if (touched) {  // faraday_rotate_pixel! (cic_shared.jl:129-159)
                const double ang = julia_mod_pi(__dmul_rn(s.rm, pw));
                const double ipol = __dsqrt_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)));
                const double psi = __dmul_rn(0.5, atan(__ddiv_rn(a1, a0)));
                double sn, cs;
                sincos(__dmul_rn(2.0, __dadd_rn(psi, ang)), &sn, &cs);
                a0 = __dmul_rn(ipol, cs);
                a1 = __dmul_rn(ipol, sn);
            }
            if (pw != 0.0) {  // update_image! (cic_shared.jl:111-121): weight plane first, then the quantities
                if (nim == 1) {
                    a1 = __dadd_rn(a1, pw);
                    a0 = __dadd_rn(a0, __dmul_rn(s.q0, pw));  // q0 == 0.0 when bin_q collapsed
                } else {
                    aw = __dadd_rn(aw, pw);
                    a0 = __dadd_rn(a0, __dmul_rn(s.q0, pw));
                    if (!(s.flags & 2)) {
                        a1 = __dadd_rn(a1, __dmul_rn(s.q1, pw));
                        for (int q = 2; q < nim; ++q) {  // further planes: plain accumulation, this thread owns the pixel
                            const double v = ld_in(P.binq, (long long)nim * (p0 + s.p) + q, P.in_dtype);
                            image[idx + npl * q] = __dadd_rn(image[idx + npl * q], __dmul_rn(v, pw));
                        }
                    }
                }
                touched = true;
                ++n_touch;
            }
        }
    }
    if (inside) {
        image[idx] = a0;
        image[idx + npl] = a1;
        if (nim > 1) image[idx + npl * nim] = aw;
        touched_map[idx] = touched ? 1 : 0;
    }
    n_touch = (unsigned long long)warp_sum_ll((long long)n_touch);
    if ((threadIdx.x & 31) == 0 && n_touch) atomicAdd(&s_touched, n_touch);
    __syncthreads();
    if (threadIdx.x == 0 && s_touched) atomicAdd(&counters[CNT_TOUCHED], s_touched);
}

template <int KID>
int run_stokes(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, const double* rm, double* image)
{
    cudaStream_t st = ctx->stream;
    const int ntile_j = (int)((G.npix + ST_T - 1) / ST_T);
    const int ntiles = ntile_j * ntile_j;
    const long long npl = G.npix * G.npix;
    void *d_touch, *d_tbeg, *d_tend, *d_tmp;
    S2G_TRY(s2g_scratch(ctx, "st_touched", (size_t)npl, &d_touch));
    S2G_TRY(s2g_scratch(ctx, "st_tbeg", sizeof(unsigned) * (ntiles + 1), &d_tbeg));
    S2G_TRY(s2g_scratch(ctx, "st_tend", sizeof(unsigned) * (ntiles + 1), &d_tend));
    S2G_CUDA(cudaMemsetAsync(d_touch, 0, (size_t)npl, st));

    long long batch = 1ll << 22;
    if (const char* e = getenv("S2G_STOKES_BATCH")) batch = std::max<long long>(1024, atoll(e));
    long long pair_cap = 1ll << 28;
    if (const char* e = getenv("S2G_STOKES_PAIR_CAP")) pair_cap = std::max<long long>(1 << 16, atoll(e));

    long long p0 = 0;
    while (p0 < P.n) {
        const long long nb = std::min(batch, P.n - p0);
        void *d_recs, *d_npt, *d_off;
        S2G_TRY(s2g_scratch(ctx, "st_recs", sizeof(SRec) * nb, &d_recs));
        S2G_TRY(s2g_scratch(ctx, "st_npt", sizeof(unsigned) * (nb + 1), &d_npt));
        S2G_TRY(s2g_scratch(ctx, "st_off", sizeof(unsigned) * (nb + 1), &d_off));
        int ph = s2g_phase_begin(ctx, PH_NORM);
        S2G_CUDA(cudaMemsetAsync((unsigned*)d_npt + nb, 0, sizeof(unsigned), st));
        const int blocks = (int)std::min<long long>((nb + 7) / 8, (long long)ctx->sm_count * 16);
        // counters are cumulative; the footprint/mapped/fallback counts of a slice that is redone are subtracted
        // by running the prep on a scratch counter block first would cost a launch; instead the slice is only
        // accepted or shrunk BEFORE anything else is added (see below) and the counters are snapshotted on the host.
        unsigned long long h_before[CNT_N];
        S2G_CUDA(cudaMemcpyAsync(h_before, ctx->d_counters, sizeof(h_before), cudaMemcpyDeviceToHost, st));
        k_stokes_prep<KID><<<std::max(blocks, 1), 256, 0, st>>>(P, G, rm, p0, nb, (SRec*)d_recs, (unsigned*)d_npt,
                                                               ctx->d_counters);
        S2G_CUDA(cudaGetLastError());
        s2g_phase_end(ctx, ph);
        ph = s2g_phase_begin(ctx, PH_SORT);
        // 64-bit total first (a slice may span more than 2^32 tile visits), then the 32-bit offsets
        cub::TransformInputIterator<unsigned long long, StU64, const unsigned*> it((const unsigned*)d_npt, StU64{});
        void* d_sum;
        S2G_TRY(s2g_scratch(ctx, "st_sum", sizeof(unsigned long long), &d_sum));
        size_t tb = 0, tb2 = 0;
        cub::DeviceReduce::Sum(nullptr, tb, it, (unsigned long long*)d_sum, (int)nb, st);
        cub::DeviceScan::ExclusiveSum(nullptr, tb2, (const unsigned*)d_npt, (unsigned*)d_off, (int)(nb + 1), st);
        S2G_TRY(s2g_scratch(ctx, "st_tmp", std::max(tb, tb2) + 16, &d_tmp));
        S2G_CUDA(cub::DeviceReduce::Sum(d_tmp, tb, it, (unsigned long long*)d_sum, (int)nb, st));
        unsigned long long h_m = 0;
        S2G_CUDA(cudaMemcpyAsync(&h_m, d_sum, sizeof(h_m), cudaMemcpyDeviceToHost, st));
        S2G_CUDA(cudaStreamSynchronize(st));
        ctx->launches += 2;
        if ((long long)h_m > pair_cap && nb > 1024) {  // too many pairs for one slice: undo its counters, halve it
            S2G_CUDA(cudaMemcpyAsync(ctx->d_counters, h_before, sizeof(h_before), cudaMemcpyHostToDevice, st));
            S2G_CUDA(cudaStreamSynchronize(st));
            s2g_phase_end(ctx, ph);
            batch = std::max<long long>(1024, nb / 2);
            continue;
        }
        S2G_CHECK(h_m < 0xfff00000ull, S2G_ENOMEM,
                  "a slice of %lld particles spans %llu image tiles: footprints too large for this image", nb, h_m);
        const long long m = (long long)h_m;
        if (m > 0) {
            S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb2, (const unsigned*)d_npt, (unsigned*)d_off, (int)(nb + 1),
                                                   st));
            void *d_keys, *d_vals, *d_keys2, *d_vals2, *d_stmp;
            S2G_TRY(s2g_scratch(ctx, "st_keys", sizeof(unsigned) * m, &d_keys));
            S2G_TRY(s2g_scratch(ctx, "st_vals", sizeof(unsigned) * m, &d_vals));
            S2G_TRY(s2g_scratch(ctx, "st_keys2", sizeof(unsigned) * m, &d_keys2));
            S2G_TRY(s2g_scratch(ctx, "st_vals2", sizeof(unsigned) * m, &d_vals2));
            k_stokes_expand<<<(int)((nb + 255) / 256), 256, 0, st>>>((const SRec*)d_recs, (const unsigned*)d_off, nb,
                                                                     ntile_j, (unsigned*)d_keys, (unsigned*)d_vals);
            S2G_CUDA(cudaGetLastError());
            int bits = 1;
            while ((1 << bits) < ntiles) ++bits;
            size_t sb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, sb, (const unsigned*)d_keys, (unsigned*)d_keys2,
                                            (const unsigned*)d_vals, (unsigned*)d_vals2, (int)m, 0, bits, st);
            S2G_TRY(s2g_scratch(ctx, "st_sort_tmp", sb + 16, &d_stmp));
            // LSD radix sort: stable, so every tile list stays in particle (= compositing) order
            S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_stmp, sb, (const unsigned*)d_keys, (unsigned*)d_keys2,
                                                     (const unsigned*)d_vals, (unsigned*)d_vals2, (int)m, 0, bits, st));
            S2G_CUDA(cudaMemsetAsync(d_tbeg, 0, sizeof(unsigned) * (ntiles + 1), st));
            S2G_CUDA(cudaMemsetAsync(d_tend, 0, sizeof(unsigned) * (ntiles + 1), st));
            k_stokes_bounds<<<(int)((m + 255) / 256), 256, 0, st>>>((const unsigned*)d_keys2, m, (unsigned*)d_tbeg,
                                                                    (unsigned*)d_tend);
            S2G_CUDA(cudaGetLastError());
            s2g_phase_end(ctx, ph);
            ph = s2g_phase_begin(ctx, PH_DEPOSIT);
            k_stokes2d<KID><<<ntiles, 256, 0, st>>>((const SRec*)d_recs, (const unsigned*)d_vals2,
                                                    (const unsigned*)d_tbeg, (const unsigned*)d_tend, ntile_j, P, G, p0,
                                                    image, (unsigned char*)d_touch, ctx->d_counters);
            S2G_CUDA(cudaGetLastError());
            ctx->launches += 5;
            ctx->host_pairs += m;
        }
        s2g_phase_end(ctx, ph);
        p0 += nb;
    }
    return S2G_OK;
}

}  // namespace

int s2g_launch_stokes_2d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, const double* rm_dev,
                         double* image)
{
    S2G_TRY(s2g_stage_wait(ctx, P.n));   // inputs may still be on their way (overlapped staging, s2g_api.cu)
    if (P.n <= 0) return S2G_OK;
    switch (kernel) {
    case S2G_KERNEL_CUBIC: return run_stokes<S2G_KERNEL_CUBIC>(ctx, P, G, rm_dev, image);
    case S2G_KERNEL_QUINTIC: return run_stokes<S2G_KERNEL_QUINTIC>(ctx, P, G, rm_dev, image);
    case S2G_KERNEL_WENDLAND_C2: return run_stokes<S2G_KERNEL_WENDLAND_C2>(ctx, P, G, rm_dev, image);
    case S2G_KERNEL_WENDLAND_C4: return run_stokes<S2G_KERNEL_WENDLAND_C4>(ctx, P, G, rm_dev, image);
    case S2G_KERNEL_WENDLAND_C6: return run_stokes<S2G_KERNEL_WENDLAND_C6>(ctx, P, G, rm_dev, image);
    case S2G_KERNEL_WENDLAND_C8: return run_stokes<S2G_KERNEL_WENDLAND_C8>(ctx, P, G, rm_dev, image);
    }
    s2g_set_error("unknown kernel id %d", kernel);
    return S2G_EINVAL;
}
