// s2g_common.cuh — shared declarations of libsphtogrid_cuda.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>

#include "sphtogrid_cuda.h"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
void s2g_set_error(const char* fmt, ...);

#define S2G_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            s2g_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (_e == cudaErrorMemoryAllocation) ? S2G_ENOMEM : S2G_ECUDA;                      \
        }                                                                                           \
    } while (0)

#define S2G_CHECK(cond, code, ...)                                                                  \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            s2g_set_error(__VA_ARGS__);                                                             \
            return (code);                                                                          \
        }                                                                                           \
    } while (0)

#define S2G_TRY(expr)                                                                               \
    do {                                                                                            \
        int _rc = (expr);                                                                           \
        if (_rc != S2G_OK) return _rc;                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------
// context: one device, one stream, a grow-only workspace arena (no cudaMalloc on the steady-state path)
// ------------------------------------------------------------------------------------------------
struct s2g_buffer {
    void* ptr = nullptr;
    size_t bytes = 0;
};

struct s2g_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int strategy = S2G_STRATEGY_AUTO;
    int exact_norm = 0;  // 1: always sum pass A numerically (never use the closed-form kernel integral)
    int accum_f32 = 0;   // 1: FP32-accumulate mode of the 2D tile-gather kernel (bar 1e-5 instead of 1e-10)
    s2g_stats stats{};
    long long host_pairs = 0;  // (tile,particle) pairs of the gather path, counted on the host
    std::map<std::string, s2g_buffer> pool;  // named scratch buffers, grow-only
    cudaEvent_t ev[10]{};
    // device-side counters (footprint pixels, touched, fallback, mapped, pairs ...)
    unsigned long long* d_counters = nullptr;  // 16 x u64
    unsigned long long* h_counters = nullptr;  // pinned mirror
    unsigned char* h_small = nullptr;          // 256 pinned bytes for the small device->host read-backs (s2g_readback)
    // per-phase device timers: (phase, start, stop) event pairs recorded on the stream, summed in s2g_get_stats
    struct timer { int phase; cudaEvent_t a, b; };
    std::vector<timer> timers;
    size_t timers_used = 0;
    int launches = 0;  // kernels of this library launched since stats_begin (cub kernels counted as one each)
    struct s2g_stager* stager = nullptr;  // overlapped host->device staging of the current call (s2g_api.cu)
};

// Host->device staging that overlaps the deposit: a helper thread copies the caller's (pageable) arrays chunk by chunk
// on its own stream and records an event per chunk; the deposit calls s2g_stage_wait(ctx, upto) before the first
// kernel that reads particles [0, upto) — a no-op when the call's inputs are already on the device.
int s2g_stage_wait(s2g_ctx* ctx, long long upto);
// first slice of a sliced deposit while a staging thread runs (small, so that the deposit starts early); 0 = no limit
long long s2g_stage_first_slice(const s2g_ctx* ctx);

// Small device->host read-backs (list lengths, pair counts) through PINNED memory.  A cudaMemcpyAsync into a pageable
// stack variable takes the driver's pageable-staging path, which serialises with the pageable host->device copies of
// the overlapped staging thread (s2g_api.cu) — measured: ~60 ms of stalls per C2 map.  Usage: add() ... then sync().
struct s2g_readback {
    s2g_ctx* ctx;
    int n = 0;
    size_t used = 0;
    struct { void* dst; size_t off, bytes; } item[8];
    explicit s2g_readback(s2g_ctx* c) : ctx(c) {}
    cudaError_t add(void* dst, const void* dev_src, size_t bytes)
    {
        item[n].dst = dst; item[n].off = used; item[n].bytes = bytes;
        const cudaError_t e = cudaMemcpyAsync(ctx->h_small + used, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        used += (bytes + 7) & ~(size_t)7;
        ++n;
        return e;
    }
    cudaError_t sync()
    {
        const cudaError_t e = cudaStreamSynchronize(ctx->stream);
        for (int i = 0; i < n; ++i) memcpy(item[i].dst, ctx->h_small + item[i].off, item[i].bytes);
        n = 0; used = 0;
        return e;
    }
};

enum { PH_PREP = 0, PH_SORT = 1, PH_NORM = 2, PH_DEPOSIT = 3, PH_EPILOGUE = 4, PH_N = 5 };
// records the start of a phase; returns a handle for s2g_phase_end
int s2g_phase_begin(s2g_ctx* ctx, int phase);
void s2g_phase_end(s2g_ctx* ctx, int handle);

// returns a device scratch buffer of at least `bytes` (contents undefined); keeps it for reuse
int s2g_scratch(s2g_ctx* ctx, const char* name, size_t bytes, void** out);

enum {
    CNT_MAPPED = 0,
    CNT_FOOTPRINT = 1,
    CNT_TOUCHED = 2,
    CNT_FALLBACK = 3,
    CNT_PAIRS = 4,
    CNT_SCATTER = 5,
    CNT_GATHER = 6,
    CNT_WORK = 7,  // dynamic work-queue cursor
    CNT_AUX = 8,   // length of a device-built work list (k_norm2d_fast -> k_norm2d)
    CNT_N = 16
};

// ------------------------------------------------------------------------------------------------
// geometry handed to the kernels
// ------------------------------------------------------------------------------------------------
struct s2g_geom {
    double len2pix;     // par.len2pix
    double inv_l3;      // 1/(len2pix*len2pix*len2pix)   (cic_2D.jl:86)
    double l3;          // (len2pix*len2pix)*len2pix      (cic_3D.jl:93)
    double half_n;      // 0.5*Npixels                    (cic_shared.jl:94-96)
    long long npix;     // Npixels[1] == [2] == [3]
    int n_images;
    int calc_mean;
};

// particle inputs as handed over the ABI (device pointers)
struct s2g_particles {
    const void* pos;   // 3 x N interleaved
    const void* hsml;
    const void* m;
    const void* rho;
    const void* binq;  // n_images x N
    const void* w;
    long long n;
    int in_dtype;      // S2G_F32 / S2G_F64
    // optional fused centre+filter (sphMapping path); when active, positions are shifted in the INPUT precision
    int fuse_center;   // 0/1
    int periodic;
    double shift[3];
    double boxsize;
    double halfsize[3];
    // optional projection applied BEFORE the recentring (map_it, cic_interpolation.jl:331-345):
    //   proj == 1: axis permutation, new component d = old component perm[d]   (rotate_to_xz/yz_plane!)
    //   proj == 2: 3x3 matrix, new = rot * old in Float64                      (rotate_3D, rot row-major)
    int proj;
    int perm[3];
    double rot[9];
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ double ld_as_f64(const void* p, long long i)
{
    return (double)__ldg(reinterpret_cast<const T*>(p) + i);
}

__device__ __forceinline__ double ld_in(const void* p, long long i, int dtype)
{
    return dtype == S2G_F64 ? __ldg(reinterpret_cast<const double*>(p) + i)
                            : (double)__ldg(reinterpret_cast<const float*>(p) + i);
}

// position component with the optional fused recentring of center_particles (filter_shift.jl:6-32):
// the subtraction is evaluated in Float64 and ROUNDED BACK to the storage type (Float32 stays Float32).
__device__ __forceinline__ double ld_pos(const s2g_particles& P, long long p, int d)
{
    if (P.proj == 2) {
        // rotate_3D (rotate_particles.jl:7-13): rot * x is a Float64 matrix whatever the storage type of x was, so
        // the recentring that follows happens in Float64.  Row times column, left to right, no contraction.
        double x0, x1, x2;
        if (P.in_dtype == S2G_F64) {
            const double* q = reinterpret_cast<const double*>(P.pos) + 3 * p;
            x0 = __ldg(q); x1 = __ldg(q + 1); x2 = __ldg(q + 2);
        } else {
            const float* q = reinterpret_cast<const float*>(P.pos) + 3 * p;
            x0 = (double)__ldg(q); x1 = (double)__ldg(q + 1); x2 = (double)__ldg(q + 2);
        }
        double v = __dadd_rn(__dadd_rn(__dmul_rn(P.rot[3 * d], x0), __dmul_rn(P.rot[3 * d + 1], x1)),
                             __dmul_rn(P.rot[3 * d + 2], x2));
        if (P.fuse_center) {
            v = __dadd_rn(v, -P.shift[d]);
            if (P.periodic) {
                double hb = P.boxsize / 2;
                if (fabs(v) > hb) v = v > 0 ? __dadd_rn(v, -hb) : __dadd_rn(v, hb);
            }
        }
        return v;
    }
    const int c = P.proj == 1 ? P.perm[d] : d;
    if (P.in_dtype == S2G_F64) {
        double v = __ldg(reinterpret_cast<const double*>(P.pos) + 3 * p + c);
        if (P.fuse_center) {
            v = __dadd_rn(v, -P.shift[d]);
            if (P.periodic) {
                double hb = P.boxsize / 2;
                if (fabs(v) > hb) v = v > 0 ? __dadd_rn(v, -hb) : __dadd_rn(v, hb);
            }
        }
        return v;
    } else {
        float f = __ldg(reinterpret_cast<const float*>(P.pos) + 3 * p + c);
        if (P.fuse_center) {
            f = __double2float_rn(__dadd_rn((double)f, -P.shift[d]));
            if (P.periodic) {
                double hb = P.boxsize / 2;
                if (fabs((double)f) > hb)
                    f = f > 0 ? __double2float_rn(__dadd_rn((double)f, -hb)) : __double2float_rn(__dadd_rn((double)f, hb));
            }
        }
        return (double)f;
    }
}

// inclusive box filter on the recentred centre (filter_shift.jl:50-58); centre of the recentred box is 0
__device__ __forceinline__ bool in_image(const s2g_particles& P, double x, double y, double z)
{
    // corner = 0 -/+ halfsize ; !(lo <= v <= hi) is also false for NaN -> rejected, like the reference
    return (-P.halfsize[0] <= x && x <= P.halfsize[0]) && (-P.halfsize[1] <= y && y <= P.halfsize[1]) &&
           (-P.halfsize[2] <= z && z <= P.halfsize[2]);
}

// SPHKernels.jl shape functions w(u) for 0 <= u <= 1 (the norm*h_inv^dim prefactor cancels in every deposit:
// kernel_norm*weight_per_pix = area/Σ(wk·dA)), evaluated in FP64.
template <int KID>
__device__ __forceinline__ double kernel_shape(double u)
{
    if (!(u < 1.0)) return 0.0;
    const double t = 1.0 - u;
    if (KID == S2G_KERNEL_CUBIC) {
        if (u < 0.5) return fma(6.0 * (u - 1.0), u * u, 1.0);
        return 2.0 * (t * t * t);
    } else if (KID == S2G_KERNEL_QUINTIC) {
        double b = fmax(2.0 / 3.0 - u, 0.0), c = fmax(1.0 / 3.0 - u, 0.0);
        double a2 = t * t, b2 = b * b, c2 = c * c;
        return fma(15.0 * c, c2 * c2, fma(-6.0 * b, b2 * b2, a2 * a2 * t));
    } else if (KID == S2G_KERNEL_WENDLAND_C2) {
        double t2 = t * t;
        return (t2 * t2) * fma(4.0, u, 1.0);
    } else if (KID == S2G_KERNEL_WENDLAND_C4) {
        double t2 = t * t;
        return (t2 * t2 * t2) * fma(fma(35.0 / 3.0, u, 6.0), u, 1.0);
    } else if (KID == S2G_KERNEL_WENDLAND_C6) {
        double t2 = t * t, t4 = t2 * t2;
        return (t4 * t4) * fma(fma(fma(32.0, u, 25.0), u, 8.0), u, 1.0);
    } else {  // WENDLAND_C8
        double t2 = t * t, t4 = t2 * t2;
        return (t4 * t4 * t2) * fma(fma(fma(fma(429.0, u, 450.0), u, 210.0), u, 50.0), u, 5.0);
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fire-and-forget FP64 add: compiles to RED.E.ADD.F64 (no return value -> no round trip)
__device__ __forceinline__ void red_add(double* addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------
// internal launchers (defined in the per-path .cu files)
// ------------------------------------------------------------------------------------------------
int s2g_launch_deposit_2d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, double* image_dev);
int s2g_launch_stokes_2d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, const double* rm_dev,
                         double* image_dev);
int s2g_launch_deposit_3d(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int kernel, double* image_dev);
int s2g_launch_footprints(s2g_ctx* ctx, const s2g_particles& P, const s2g_geom& G, int dims, long long* bounds_dev);
int s2g_launch_reduce_2d(s2g_ctx* ctx, const double* image_dev, long long nx, long long ny, int n_images,
                         int reduce_image, double* out_dev);
int s2g_launch_reduce_3d(s2g_ctx* ctx, const double* image_dev, long long npix, int reduce_image, double* out_dev);
int s2g_launch_center_filter(s2g_ctx* ctx, const s2g_particles& P, void* pos_out_dev, uint8_t* mask_dev);
int s2g_launch_healpix(s2g_ctx* ctx, const s2g_particles& P, long long nside, int kernel, int calc_mean,
                       const unsigned char* take, double* map_dev, double* wmap_dev);
int s2g_launch_healpix_filtered(s2g_ctx* ctx, const s2g_particles& P, double r0, double r1, long long nside,
                                int kernel, int calc_mean, double* map_dev, double* wmap_dev, long long* n_selected);
int s2g_hp_radii(s2g_ctx* ctx, const s2g_particles& P, double r0, double r1, unsigned long long* keys_dev,
                 unsigned char* sel_dev, long long* n_selected);
int s2g_hp_take_mask(s2g_ctx* ctx, const unsigned long long* keys_dev, const unsigned char* sel_dev, long long n,
                     unsigned char* take_dev);
int s2g_launch_healpix_pixels(s2g_ctx* ctx, const double pos[3], double radius, long long nside, long long* out_dev,
                              long long cap, long long* count_dev);
int s2g_launch_stencil(s2g_ctx* ctx, int order, int dims, const void* pos, const void* q, long long n, int in_dtype,
                       double len2pix, long long npix, int periodic, double* image_dev);
int s2g_launch_accumulate_finite(s2g_ctx* ctx, double* sum_dev, const double* local_dev, long long n);
int s2g_launch_divide_slice(s2g_ctx* ctx, int dims, double* q, const double* w, long long n, long long stride,
                            int n_images, int reduce_image);
int s2g_launch_synth(s2g_ctx* ctx, uint64_t seed, long long first_id, long long n, long long n_total, double box,
                     double n_ngb, double sigma, int out_dtype, void* pos, void* hsml, void* m, void* rho, void* temp);
int s2g_run_microbench(s2g_ctx* ctx, int which, size_t bytes, int iters, double* rate_out);

// first halves of the host entry points s2g_sphmap / s2g_healpix_map (s2g_api.cu), shared with the device group
int s2g_sphmap_stage_deposit(const char* fn, s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml,
                             const void* m, const void* rho, const void* binq, const void* w, int64_t n,
                             int32_t n_images, int32_t in_dtype, const int32_t* perm, const double* rot,
                             const double shift[3], int32_t periodic, double boxsize, const double halfsize[3],
                             double len2pix, int64_t npix, int32_t kernel, int32_t calc_mean, void* pos_recentred_out,
                             double** image_dev_out);
int s2g_healpix_stage_deposit(const char* fn, s2g_ctx* ctx, const void* pos, const void* hsml, const void* m,
                              const void* rho, const void* binq, const void* w, int64_t n, const double center[3],
                              const double radius_limits[2], int64_t nside, int32_t kernel, int32_t calc_mean,
                              void* pos_recentred_out, double** maps_dev_out, long long* n_selected);
// copies the device counters and phase timers into ctx->stats (synchronises the context stream)
int s2g_stats_collect(s2g_ctx* ctx);
int s2g_stats_begin(s2g_ctx* ctx, long long n_in);
// H2D staging of the six particle arrays into the context's scratch buffers
int s2g_stage_particles(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                        const void* binq, const void* w, int64_t n, int n_images, int in_dtype, s2g_particles& P);
