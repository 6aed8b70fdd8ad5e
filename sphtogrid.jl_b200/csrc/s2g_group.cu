// s2g_group.cu — one process, several GPUs: the `parallel=true` branch of sphMapping
// (src/cic_interpolation/cic_interpolation.jl:171-215, 236-271) and the per-worker split of healpix_map driven INSIDE
// the library, for callers that have no process group of their own (a single Julia process with 8 visible GPUs).
//
//   phase A  one host thread per device: the particles of slice r of domain_decomposition
//            (src/parallel/domain_decomp.jl:7-17) are staged on device r and deposited into a full private flat image
//            (`@spawnat batch[i] cic_mapping_2D(x[:,batch[i]], ...)`, :185-196).
//   phase B  `image = sum(fetch.(futures))` (:199, :256) + reduce_image (:212, :234) as ONE kernel per device over
//            peer memory: device r owns a slice of the pixels, loads that slice of EVERY device's image directly over
//            NVLink (P2P loads; rank order 0,1,2.. so the sum is deterministic), applies the divide / transposition
//            epilogue and copies its slice of the result to the caller's host buffer.  No collective library is
//            involved.  Without peer access (or with S2G_GROUP_NO_P2P=1) the peer images are first copied with
//            cudaMemcpyPeerAsync and the same kernel reads the copies.
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "s2g_common.cuh"

#define S2G_MAX_GROUP 16

struct s2g_group {
    std::vector<s2g_ctx*> ctx;
    std::vector<int> devices;
    int p2p = 1;  // 1: every pair of distinct devices has peer access enabled -> direct loads in phase B
};

// per-device image base pointers and plane strides (in doubles) handed to the phase-B kernels
struct s2g_peer_srcs {
    const double* ptr[S2G_MAX_GROUP];
    long long stride[S2G_MAX_GROUP];
    int n;
};

// ------------------------------------------------------------------------------------------------
// domain_decomposition (parallel/domain_decomp.jl:7-17): size = floor(N/W); slice i = [i*size, (i+1)*size), the last
// slice takes the remainder.  0-based starts.
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_domain_decomposition(int64_t n, int32_t n_parts, int64_t* starts_out, int64_t* counts_out)
{
    S2G_CHECK(n >= 0 && n_parts >= 1 && starts_out && counts_out, S2G_EINVAL, "%s: bad arguments", __func__);
    const int64_t size = n / n_parts;
    for (int i = 0; i < n_parts; ++i) {
        starts_out[i] = (int64_t)i * size;
        counts_out[i] = i == n_parts - 1 ? n - (int64_t)i * size : size;
    }
    return S2G_OK;
}

// pixel slice [lo, hi) of device r: balanced split of `total` units
static inline void pixel_slice(long long total, int r, int ndev, long long& lo, long long& hi)
{
    lo = total * r / ndev;
    hi = total * (r + 1) / ndev;
}

// ------------------------------------------------------------------------------------------------
// phase-B kernels.  Peer images were completed before the launch (host barrier after phase A), nobody writes them
// while these kernels run; plain loads (not ld.global.nc) because the sources are peer-device memory.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double peer_sum(const s2g_peer_srcs& S, int plane, long long c)
{
    double acc = S.ptr[0][plane * S.stride[0] + c];
    for (int r = 1; r < S.n; ++r) acc = __dadd_rn(acc, S.ptr[r][plane * S.stride[r] + c]);
    return acc;
}

// flat images / HEALPix maps: out[pl*(c1-c0) + (c-c0)] = Σ_r image_r[pl][c]
__global__ void __launch_bounds__(256) k_group_sum(s2g_peer_srcs S, int planes, long long c0, long long c1,
                                                   double* __restrict__ out)
{
    const long long width = c1 - c0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int pl = 0; pl < planes; ++pl)
        for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < width; e += stride)
            out[pl * width + e] = peer_sum(S, pl, c0 + e);
}

// reduce_image_3D (reduce_image.jl:39-55) on the summed image: division gated on the QUANTITY plane (Q7);
// reduce_image == 0 -> the weight plane counts as 1 (cic_interpolation.jl:230-232)
__global__ void __launch_bounds__(256) k_group_reduce3d(s2g_peer_srcs S, long long c0, long long c1, int reduce_image,
                                                        double* __restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = c0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; e < c1; e += stride) {
        double v = peer_sum(S, 0, e);
        if (v > 0.0) {
            const double wv = reduce_image ? peer_sum(S, 1, e) : 1.0;
            v = v / wv;
        }
        out[e - c0] = v;
    }
}

// reduce_image_2D (reduce_image.jl:8-31) on the summed image for the columns iy in [y0,y1) of the flat image
// (k = ix*nx + iy): out[q][iy-y0][ix] = Σ image[q][k] (/ Σ weight[k] where reduce && weight > 0), i.e. the slice
// [nx*y0, nx*y1) of every plane of Julia's Array(nx,ny,n_images).  32x32 tiles through shared memory: loads run along
// iy (contiguous in the peers' images), stores along ix (contiguous in the result).
__global__ void __launch_bounds__(256) k_group_reduce2d(s2g_peer_srcs S, long long nx, int n_images, int reduce_image,
                                                        long long y0, long long y1, double* __restrict__ out)
{
    __shared__ double tile[32][33];
    const long long wy = y1 - y0;
    const long long bx = blockIdx.x * 32LL, by = blockIdx.y * 32LL;  // bx: local iy block, by: ix block
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;         // 32 x 8
    double wv[4] = {0.0, 0.0, 0.0, 0.0};
    if (reduce_image) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const long long ix = by + ty + 8 * g, iyl = bx + tx;
            if (ix < nx && iyl < wy) wv[g] = peer_sum(S, n_images, ix * nx + y0 + iyl);
        }
    }
    for (int q = 0; q < n_images; ++q) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int r = ty + 8 * g;
            const long long ix = by + r, iyl = bx + tx;
            if (ix < nx && iyl < wy) {
                double v = peer_sum(S, q, ix * nx + y0 + iyl);
                if (reduce_image && wv[g] > 0.0) v = v / wv[g];
                tile[r][tx] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int r = ty + 8 * g;
            const long long iyl = bx + r, ix = by + tx;
            if (ix < nx && iyl < wy) out[(long long)q * nx * wy + iyl * nx + ix] = tile[tx][r];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct rank_result {
    int rc = S2G_OK;
    std::string err;
};

// runs f(r) for r in [0, n) on one host thread per rank (inline when n == 1); the first failure is returned and its
// message becomes the calling thread's s2g_last_error()
template <class F>
static int run_ranks(int n, F f)
{
    std::vector<rank_result> res((size_t)n);
    auto body = [&](int r) {
        res[(size_t)r].rc = f(r);
        if (res[(size_t)r].rc != S2G_OK) res[(size_t)r].err = s2g_last_error();
    };
    if (n == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        th.reserve((size_t)n);
        for (int r = 0; r < n; ++r) th.emplace_back(body, r);
        for (auto& t : th) t.join();
    }
    for (int r = 0; r < n; ++r)
        if (res[(size_t)r].rc != S2G_OK) {
            s2g_set_error("device %d of the group: %s", r, res[(size_t)r].err.c_str());
            return res[(size_t)r].rc;
        }
    return S2G_OK;
}

extern "C" int s2g_group_init(const int32_t* devices, int32_t n_devices, s2g_group** out)
{
    S2G_CHECK(out != nullptr, S2G_EINVAL, "s2g_group_init: out is NULL");
    *out = nullptr;
    S2G_CHECK(n_devices >= 1 && n_devices <= S2G_MAX_GROUP, S2G_EINVAL, "s2g_group_init: n_devices must be in [1, %d]",
              S2G_MAX_GROUP);
    S2G_CHECK(devices != nullptr, S2G_EINVAL, "s2g_group_init: devices is NULL");
    s2g_group* grp = new (std::nothrow) s2g_group();
    S2G_CHECK(grp != nullptr, S2G_ENOMEM, "s2g_group_init: out of host memory");
    for (int i = 0; i < n_devices; ++i) {
        s2g_ctx* c = nullptr;
        const int rc = s2g_init(devices[i], &c);
        if (rc != S2G_OK) {
            for (auto* k : grp->ctx) s2g_shutdown(k);
            delete grp;
            return rc;
        }
        grp->ctx.push_back(c);
        grp->devices.push_back(devices[i]);
    }
    // peer access between every ordered pair of distinct devices; a device may be listed more than once (two contexts
    // on one GPU see each other's memory anyway)
    const char* no_p2p = getenv("S2G_GROUP_NO_P2P");
    grp->p2p = (no_p2p && atoi(no_p2p) != 0) ? 0 : 1;
    for (int a = 0; a < n_devices && grp->p2p; ++a)
        for (int b = 0; b < n_devices && grp->p2p; ++b) {
            if (devices[a] == devices[b]) continue;
            int ok = 0;
            if (cudaDeviceCanAccessPeer(&ok, devices[a], devices[b]) != cudaSuccess || !ok) {
                cudaGetLastError();
                grp->p2p = 0;
                break;
            }
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) grp->p2p = 0;
            cudaGetLastError();
        }
    *out = grp;
    return S2G_OK;
}

extern "C" int s2g_group_shutdown(s2g_group* grp)
{
    if (!grp) return S2G_OK;
    for (auto* c : grp->ctx) s2g_shutdown(c);
    delete grp;
    return S2G_OK;
}

extern "C" int s2g_group_size(const s2g_group* grp) { return grp ? (int)grp->ctx.size() : 0; }

extern "C" int s2g_group_peer_access(const s2g_group* grp) { return grp ? grp->p2p : 0; }

extern "C" int s2g_group_context(s2g_group* grp, int32_t rank, s2g_ctx** out)
{
    S2G_CHECK(grp && out, S2G_EINVAL, "%s: NULL argument", __func__);
    S2G_CHECK(rank >= 0 && rank < (int)grp->ctx.size(), S2G_EINVAL, "%s: rank %d out of range [0,%d)", __func__, rank,
              (int)grp->ctx.size());
    *out = grp->ctx[(size_t)rank];
    return S2G_OK;
}

static float gev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

// after phase A of rank r: counters + phase timers (synchronises the stream = "image r is complete")
static int finish_phase_a(s2g_ctx* c)
{
    S2G_TRY(s2g_stats_collect(c));
    c->stats.ms_h2d = gev_ms(c->ev[0], c->ev[1]);
    c->stats.ms_compute = gev_ms(c->ev[1], c->ev[2]);
    return S2G_OK;
}

// the image table device r reads in phase B: direct peer pointers, or copies of the peer images on device r
static int peer_table(s2g_group* grp, int r, const std::vector<double*>& img, size_t image_doubles, long long plane_stride,
                      s2g_peer_srcs& S)
{
    const int ndev = (int)grp->ctx.size();
    s2g_ctx* c = grp->ctx[(size_t)r];
    S.n = ndev;
    char* stage = nullptr;
    if (!grp->p2p && ndev > 1) {
        void* p;
        S2G_TRY(s2g_scratch(c, "peer_stage", sizeof(double) * image_doubles * (size_t)(ndev - 1), &p));
        stage = (char*)p;
    }
    int slot = 0;
    for (int p = 0; p < ndev; ++p) {
        S.stride[p] = plane_stride;
        if (p == r || grp->p2p) {
            S.ptr[p] = img[(size_t)p];
        } else {
            double* dst = (double*)(stage + sizeof(double) * image_doubles * (size_t)slot++);
            S2G_CUDA(cudaMemcpyPeerAsync(dst, grp->devices[(size_t)r], img[(size_t)p], grp->devices[(size_t)p],
                                         sizeof(double) * image_doubles, c->stream));
            S.ptr[p] = dst;
        }
    }
    return S2G_OK;
}

static int grid_for(s2g_ctx* c, long long work)
{
    long long b = (work + 255) / 256;
    const long long cap = (long long)c->sm_count * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------------------
// sphMapping(...; parallel=true) on the devices of the group
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_group_sphmap(s2g_group* grp, int32_t dims, const void* pos, const void* hsml, const void* m,
                                const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                                int32_t in_dtype, const double shift[3], int32_t periodic, double boxsize,
                                const double halfsize[3], double len2pix, int64_t npix, int32_t kernel,
                                int32_t calc_mean, int32_t reduce_image, int32_t return_both_maps,
                                void* pos_recentred_out, double* out, s2g_stats* stats)
{
    S2G_CHECK(grp != nullptr, S2G_EINVAL, "%s: group is NULL", __func__);
    S2G_CHECK(out != nullptr, S2G_EINVAL, "%s: out is NULL", __func__);
    S2G_CHECK(n >= 0, S2G_EINVAL, "%s: n < 0", __func__);
    S2G_CHECK(in_dtype == S2G_F32 || in_dtype == S2G_F64, S2G_EINVAL, "%s: in_dtype must be 0 (f32) or 1 (f64)", __func__);
    S2G_CHECK(dims == 2 || dims == 3, S2G_EINVAL, "%s: dims must be 2 or 3", __func__);
    S2G_CHECK(n_images >= 1 && n_images <= 64, S2G_EINVAL, "%s: n_images out of range", __func__);
    const int ndev = (int)grp->ctx.size();
    const size_t es = in_dtype == S2G_F64 ? 8 : 4;
    std::vector<int64_t> start((size_t)ndev), count((size_t)ndev);
    S2G_TRY(s2g_domain_decomposition(n, ndev, start.data(), count.data()));
    std::vector<double*> img((size_t)ndev, nullptr);
    const char* fn = __func__;

    // phase A: every device deposits its slice into a full private image
    S2G_TRY(run_ranks(ndev, [&](int r) -> int {
        const size_t s = (size_t)start[(size_t)r];
        auto off = [&](const void* p, size_t per) { return p ? (const void*)((const char*)p + s * per * es) : p; };
        void* pr = pos_recentred_out ? (void*)((char*)pos_recentred_out + s * 3 * es) : nullptr;
        s2g_ctx* c = grp->ctx[(size_t)r];
        S2G_TRY(s2g_sphmap_stage_deposit(fn, c, dims, off(pos, 3), off(hsml, 1), off(m, 1), off(rho, 1),
                                         off(binq, (size_t)n_images), off(w, 1), count[(size_t)r], n_images, in_dtype,
                                         nullptr, nullptr, shift, periodic, boxsize, halfsize, len2pix, npix, kernel,
                                         calc_mean, pr, &img[(size_t)r]));
        return finish_phase_a(c);
    }));

    // phase B: peer-memory sum + reduce_image epilogue, one pixel slice per device
    const long long ncell = dims == 2 ? (long long)(npix * npix) : (long long)(npix * npix * npix);
    const int planes = dims == 2 ? n_images + 1 : 2;
    const bool both = dims == 2 && return_both_maps;
    S2G_TRY(run_ranks(ndev, [&](int r) -> int {
        s2g_ctx* c = grp->ctx[(size_t)r];
        S2G_CUDA(cudaSetDevice(c->device));
        S2G_CUDA(cudaEventRecord(c->ev[5], c->stream));
        s2g_peer_srcs S;
        S2G_TRY(peer_table(grp, r, img, (size_t)ncell * (size_t)planes, ncell, S));
        void* dslab;
        if (both) {
            long long c0, c1;
            pixel_slice(ncell, r, ndev, c0, c1);
            const long long width = c1 - c0;
            S2G_TRY(s2g_scratch(c, "reduced", sizeof(double) * (size_t)(width > 0 ? width : 1) * (size_t)planes, &dslab));
            if (width > 0) {
                k_group_sum<<<grid_for(c, width), 256, 0, c->stream>>>(S, planes, c0, c1, (double*)dslab);
                S2G_CUDA(cudaGetLastError());
                S2G_CUDA(cudaEventRecord(c->ev[6], c->stream));
                for (int pl = 0; pl < planes; ++pl)
                    S2G_CUDA(cudaMemcpyAsync(out + (size_t)pl * (size_t)ncell + (size_t)c0,
                                             (double*)dslab + (size_t)pl * (size_t)width, sizeof(double) * (size_t)width,
                                             cudaMemcpyDeviceToHost, c->stream));
            } else {
                S2G_CUDA(cudaEventRecord(c->ev[6], c->stream));
            }
        } else if (dims == 2) {
            long long y0, y1;
            pixel_slice(npix, r, ndev, y0, y1);
            const long long wy = y1 - y0;
            S2G_TRY(s2g_scratch(c, "reduced", sizeof(double) * (size_t)(wy > 0 ? wy : 1) * (size_t)npix * (size_t)n_images,
                                &dslab));
            if (wy > 0) {
                dim3 grid((unsigned)((wy + 31) / 32), (unsigned)((npix + 31) / 32));
                k_group_reduce2d<<<grid, 256, 0, c->stream>>>(S, npix, n_images, reduce_image, y0, y1, (double*)dslab);
                S2G_CUDA(cudaGetLastError());
                S2G_CUDA(cudaEventRecord(c->ev[6], c->stream));
                for (int q = 0; q < n_images; ++q)
                    S2G_CUDA(cudaMemcpyAsync(out + (size_t)q * (size_t)ncell + (size_t)(npix * y0),
                                             (double*)dslab + (size_t)q * (size_t)(npix * wy),
                                             sizeof(double) * (size_t)(npix * wy), cudaMemcpyDeviceToHost, c->stream));
            } else {
                S2G_CUDA(cudaEventRecord(c->ev[6], c->stream));
            }
        } else {
            long long c0, c1;
            pixel_slice(ncell, r, ndev, c0, c1);
            const long long width = c1 - c0;
            S2G_TRY(s2g_scratch(c, "reduced", sizeof(double) * (size_t)(width > 0 ? width : 1), &dslab));
            if (width > 0) {
                k_group_reduce3d<<<grid_for(c, width), 256, 0, c->stream>>>(S, c0, c1, reduce_image, (double*)dslab);
                S2G_CUDA(cudaGetLastError());
                S2G_CUDA(cudaEventRecord(c->ev[6], c->stream));
                S2G_CUDA(cudaMemcpyAsync(out + (size_t)c0, dslab, sizeof(double) * (size_t)width, cudaMemcpyDeviceToHost,
                                         c->stream));
            } else {
                S2G_CUDA(cudaEventRecord(c->ev[6], c->stream));
            }
        }
        S2G_CUDA(cudaEventRecord(c->ev[7], c->stream));
        S2G_CUDA(cudaStreamSynchronize(c->stream));
        c->launches += 1;
        c->stats.n_launches = c->launches;
        c->stats.ms_epilogue = gev_ms(c->ev[5], c->ev[6]);
        c->stats.ms_d2h = gev_ms(c->ev[6], c->ev[7]);
        c->stats.ms_total = c->stats.ms_h2d + c->stats.ms_compute + c->stats.ms_epilogue + c->stats.ms_d2h;
        return S2G_OK;
    }));
    if (stats)
        for (int r = 0; r < ndev; ++r) stats[r] = grp->ctx[(size_t)r]->stats;
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// healpix_map (src/healpix_interpolation/main.jl:92-227) on the devices of the group.  The result is the one of the
// single call on the full arrays: filter_sort_particles deposits `sorted[sel]` with the shell mask in ORIGINAL order
// (filter_particles.jl:33-41), which depends on the radii of ALL particles, so when some particle is outside the shell
// the radii of all slices are brought to device 0 (peer copies), the selection is made there once and the take mask
// is handed back slice by slice.  stats[r].n_in = particles of slice r inside the shell.
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_group_healpix_map(s2g_group* grp, const void* pos, const void* hsml, const void* m, const void* rho,
                                     const void* binq, const void* w, int64_t n, const double center[3],
                                     const double radius_limits[2], int64_t nside, int32_t kernel, int32_t calc_mean,
                                     void* pos_recentred_out, double* map_out, double* wmap_out, s2g_stats* stats)
{
    S2G_CHECK(grp != nullptr, S2G_EINVAL, "%s: group is NULL", __func__);
    S2G_CHECK(map_out && wmap_out && center && radius_limits, S2G_EINVAL, "%s: NULL argument", __func__);
    S2G_CHECK(n >= 0 && n < 2147483647LL, S2G_EINVAL, "%s: n out of range [0, 2^31-1)", __func__);
    S2G_CHECK(n == 0 || (pos && hsml && m && rho && binq && w), S2G_EINVAL, "%s: NULL particle array", __func__);
    S2G_CHECK(nside >= 1 && nside <= 8192 && (nside & (nside - 1)) == 0, S2G_EINVAL,
              "%s: nside must be a power of two in [1, 8192]", __func__);
    S2G_CHECK(kernel >= S2G_KERNEL_CUBIC && kernel <= S2G_KERNEL_WENDLAND_C8, S2G_EINVAL, "%s: unknown kernel id %d",
              __func__, kernel);
    const int ndev = (int)grp->ctx.size();
    std::vector<int64_t> start((size_t)ndev), count((size_t)ndev);
    S2G_TRY(s2g_domain_decomposition(n, ndev, start.data(), count.data()));
    const long long npix = 12LL * nside * nside;
    std::vector<double*> maps((size_t)ndev, nullptr);
    std::vector<long long> nsel((size_t)ndev, 0);
    std::vector<s2g_particles> parts((size_t)ndev);
    std::vector<unsigned long long*> keys((size_t)ndev, nullptr);
    std::vector<unsigned char*> sel((size_t)ndev, nullptr), take((size_t)ndev, nullptr);

    // phase A1: stage the slice, zero the maps, radii + shell mask of the slice
    S2G_TRY(run_ranks(ndev, [&](int r) -> int {
        s2g_ctx* c = grp->ctx[(size_t)r];
        S2G_CUDA(cudaSetDevice(c->device));
        const size_t s = (size_t)start[(size_t)r];
        const int64_t cnt = count[(size_t)r];
        auto off = [&](const void* p, size_t per) { return p ? (const void*)((const char*)p + s * per * 8) : p; };
        S2G_TRY(s2g_stats_begin(c, cnt));
        void *dmaps, *dk, *ds, *dt;
        S2G_TRY(s2g_scratch(c, "image", sizeof(double) * (size_t)npix * 2, &dmaps));
        const size_t nn = (size_t)(cnt > 0 ? cnt : 1);
        S2G_TRY(s2g_scratch(c, "hp_keys", sizeof(unsigned long long) * nn, &dk));
        S2G_TRY(s2g_scratch(c, "hp_sel", nn, &ds));
        S2G_TRY(s2g_scratch(c, "hp_take", nn, &dt));
        maps[(size_t)r] = (double*)dmaps;
        keys[(size_t)r] = (unsigned long long*)dk;
        sel[(size_t)r] = (unsigned char*)ds;
        take[(size_t)r] = (unsigned char*)dt;
        S2G_CUDA(cudaEventRecord(c->ev[0], c->stream));
        s2g_particles& P = parts[(size_t)r];
        S2G_TRY(s2g_stage_particles(c, off(pos, 3), off(hsml, 1), off(m, 1), off(rho, 1), off(binq, 1), off(w, 1), cnt, 1,
                                    S2G_F64, P));
        S2G_CUDA(cudaMemsetAsync(dmaps, 0, sizeof(double) * (size_t)npix * 2, c->stream));
        S2G_CUDA(cudaEventRecord(c->ev[1], c->stream));
        P.fuse_center = 1;  // Pos .-= center (Float64), no periodic wrap, no box filter
        P.periodic = 0;
        for (int d = 0; d < 3; ++d) { P.shift[d] = center[d]; P.halfsize[d] = 0.0; }
        return s2g_hp_radii(c, P, radius_limits[0], radius_limits[1], keys[(size_t)r], sel[(size_t)r], &nsel[(size_t)r]);
    }));
    long long nsel_total = 0;
    for (int r = 0; r < ndev; ++r) nsel_total += nsel[(size_t)r];

    // the `sorted[sel]` selection over ALL particles, on device 0
    const bool select = nsel_total != n;
    if (select) {
        s2g_ctx* c0 = grp->ctx[0];
        S2G_CUDA(cudaSetDevice(c0->device));
        void *gk, *gs, *gt;
        S2G_TRY(s2g_scratch(c0, "hp_gkeys", sizeof(unsigned long long) * (size_t)n, &gk));
        S2G_TRY(s2g_scratch(c0, "hp_gsel", (size_t)n, &gs));
        S2G_TRY(s2g_scratch(c0, "hp_gtake", (size_t)n, &gt));
        for (int r = 0; r < ndev; ++r) {
            if (count[(size_t)r] == 0) continue;
            S2G_CUDA(cudaMemcpyPeerAsync((unsigned long long*)gk + start[(size_t)r], grp->devices[0], keys[(size_t)r],
                                         grp->devices[(size_t)r], sizeof(unsigned long long) * (size_t)count[(size_t)r],
                                         c0->stream));
            S2G_CUDA(cudaMemcpyPeerAsync((unsigned char*)gs + start[(size_t)r], grp->devices[0], sel[(size_t)r],
                                         grp->devices[(size_t)r], (size_t)count[(size_t)r], c0->stream));
        }
        S2G_TRY(s2g_hp_take_mask(c0, (const unsigned long long*)gk, (const unsigned char*)gs, n, (unsigned char*)gt));
        for (int r = 0; r < ndev; ++r) {
            if (count[(size_t)r] == 0) continue;
            S2G_CUDA(cudaMemcpyPeerAsync(take[(size_t)r], grp->devices[(size_t)r], (unsigned char*)gt + start[(size_t)r],
                                         grp->devices[0], (size_t)count[(size_t)r], c0->stream));
        }
        S2G_CUDA(cudaStreamSynchronize(c0->stream));
    }

    // phase A2: particle loop of every slice into its private pair of maps
    S2G_TRY(run_ranks(ndev, [&](int r) -> int {
        s2g_ctx* c = grp->ctx[(size_t)r];
        S2G_CUDA(cudaSetDevice(c->device));
        const size_t s = (size_t)start[(size_t)r];
        const int64_t cnt = count[(size_t)r];
        const s2g_particles& P = parts[(size_t)r];
        S2G_TRY(s2g_launch_healpix(c, P, nside, kernel, calc_mean, select ? take[(size_t)r] : nullptr, maps[(size_t)r],
                                   maps[(size_t)r] + npix));
        S2G_CUDA(cudaEventRecord(c->ev[2], c->stream));
        if (pos_recentred_out && cnt > 0) {  // the reference recentres the caller's Pos in place (filter_particles.jl:20)
            void* dpo;
            S2G_TRY(s2g_scratch(c, "pos_out", 3 * (size_t)cnt * sizeof(double), &dpo));
            S2G_TRY(s2g_launch_center_filter(c, P, dpo, nullptr));
            S2G_CUDA(cudaMemcpyAsync((char*)pos_recentred_out + s * 3 * 8, dpo, 3 * (size_t)cnt * sizeof(double),
                                     cudaMemcpyDeviceToHost, c->stream));
        }
        return finish_phase_a(c);
    }));

    // phase B: the two maps summed over peer memory, one pixel slice per device
    S2G_TRY(run_ranks(ndev, [&](int r) -> int {
        s2g_ctx* c = grp->ctx[(size_t)r];
        S2G_CUDA(cudaSetDevice(c->device));
        S2G_CUDA(cudaEventRecord(c->ev[5], c->stream));
        s2g_peer_srcs S;
        S2G_TRY(peer_table(grp, r, maps, (size_t)npix * 2, npix, S));
        long long c0, c1;
        pixel_slice(npix, r, ndev, c0, c1);
        const long long width = c1 - c0;
        void* dslab;
        S2G_TRY(s2g_scratch(c, "reduced", sizeof(double) * (size_t)(width > 0 ? width : 1) * 2, &dslab));
        if (width > 0) {
            k_group_sum<<<grid_for(c, width), 256, 0, c->stream>>>(S, 2, c0, c1, (double*)dslab);
            S2G_CUDA(cudaGetLastError());
        }
        S2G_CUDA(cudaEventRecord(c->ev[6], c->stream));
        if (width > 0) {
            S2G_CUDA(cudaMemcpyAsync(map_out + (size_t)c0, dslab, sizeof(double) * (size_t)width, cudaMemcpyDeviceToHost,
                                     c->stream));
            S2G_CUDA(cudaMemcpyAsync(wmap_out + (size_t)c0, (double*)dslab + (size_t)width, sizeof(double) * (size_t)width,
                                     cudaMemcpyDeviceToHost, c->stream));
        }
        S2G_CUDA(cudaEventRecord(c->ev[7], c->stream));
        S2G_CUDA(cudaStreamSynchronize(c->stream));
        c->launches += 1;
        c->stats.n_launches = c->launches;
        c->stats.n_in = nsel[(size_t)r];
        c->stats.ms_epilogue = gev_ms(c->ev[5], c->ev[6]);
        c->stats.ms_d2h = gev_ms(c->ev[6], c->ev[7]);
        c->stats.ms_total = c->stats.ms_h2d + c->stats.ms_compute + c->stats.ms_epilogue + c->stats.ms_d2h;
        return S2G_OK;
    }));
    if (stats)
        for (int r = 0; r < ndev; ++r) stats[r] = grp->ctx[(size_t)r]->stats;
    return S2G_OK;
}
