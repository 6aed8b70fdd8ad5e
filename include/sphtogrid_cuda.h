/*
 * sphtogrid_cuda.h — C ABI of libsphtogrid_cuda.so
 *
 * B200 (sm_100a) implementation of the particle-deposition hot path of
 * SPHtoGrid.jl v0.5.3.  The reference is pure Julia and has no FFI of its own;
 * each entry point below replaces the body of one internal Julia function that
 * the public API (sphMapping / healpix_map) calls, and is what the Julia glue
 * (sphtogrid.jl_b200/julia/SPHtoGridCUDA.jl) binds with
 * `ccall((:s2g_xxx, "libsphtogrid_cuda"), Cint, (...), ...)`.
 * Reference locations are given as file:line relative to the reference root.
 *
 * Conventions
 *  - plain C types only; every pointer is caller-owned; the library never keeps
 *    a caller pointer after the call returns and never writes to an input.
 *  - host entry points take HOST pointers (Julia Arrays under GC.@preserve);
 *    `_dev` variants take DEVICE pointers on the context's device and enqueue on
 *    the context's stream (they return after the work is enqueued unless stated;
 *    call s2g_sync() or use the returned stats, which force a sync).
 *  - `pos` is exactly the memory of a Julia Matrix{T}(3,N): xyz interleaved.
 *  - `binq` is Julia Matrix{T}(n_images,N) memory (or Vector{T}(N), n_images=1).
 *  - flat images are plane-separated: plane q at image + q*n_pixels (Julia
 *    column-major Matrix(n_pixels, n_images+1)), weight plane last; the flat
 *    pixel index is the reference's calculate_index (src/shared/indices.jl:6-17)
 *    minus one: 2D i*nx + j, 3D i*nx*ny + j*ny + k.
 *  - HEALPix maps are RING ordered, element p (0-based) = Julia pixels[p+1].
 *  - every function returns S2G_OK (0) or a negative s2g_status; the message is
 *    available from s2g_last_error() (thread-local).  Nothing throws.
 *  - a context is bound to ONE device and is not re-entrant.  Multi-GPU runs either
 *    use one process (and one context) per GPU, the partial images being combined
 *    by the host layer with NCCL (sphtogrid.jl_b200/distributed.py), or ONE process
 *    with a device group (s2g_group_*, below): the library shards the particles,
 *    runs one host thread per device and sums the partial images over peer memory
 *    (NVLink).  Both replace `@spawnat` + `sum(fetch.(futures))` of
 *    src/cic_interpolation/cic_interpolation.jl:185-199, 244-256.
 *  - there is NO CPU fallback: without a usable CUDA device every compute entry
 *    point fails with S2G_ECUDA.
 */
#ifndef SPHTOGRID_CUDA_H
#define SPHTOGRID_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2G_API __attribute__((visibility("default")))

typedef enum {
    S2G_OK = 0,
    S2G_EINVAL = -1,       /* bad argument */
    S2G_ECUDA = -2,        /* CUDA runtime error / no device */
    S2G_ENOMEM = -3,       /* device or host allocation failed */
    S2G_EUNSUPPORTED = -4, /* reserved */
    S2G_EINTERNAL = -5
} s2g_status;

/* SPHKernels.jl kernel types (call site src/cic_interpolation/cic_shared.jl:24) */
typedef enum {
    S2G_KERNEL_CUBIC = 0,
    S2G_KERNEL_QUINTIC = 1,
    S2G_KERNEL_WENDLAND_C2 = 2,
    S2G_KERNEL_WENDLAND_C4 = 3,
    S2G_KERNEL_WENDLAND_C6 = 4,
    S2G_KERNEL_WENDLAND_C8 = 5
} s2g_kernel;

typedef enum { S2G_F32 = 0, S2G_F64 = 1 } s2g_dtype;

/* deposit strategy (default AUTO: per-particle choice by footprint size) */
typedef enum {
    S2G_STRATEGY_AUTO = 0,
    S2G_STRATEGY_SCATTER = 1, /* warp-per-particle, red.global.add.f64           */
    S2G_STRATEGY_GATHER = 2   /* tile-owning CTAs, register accumulators, no atomics */
} s2g_strategy;

/* arithmetic of the per-pixel inner loop (north_star: FP64 mode to 1e-10, optional FP32-accumulate mode to 1e-5) */
typedef enum {
    S2G_ACCUM_F64 = 0, /* default: everything in FP64                                                       */
    S2G_ACCUM_F32 = 1  /* 2D tile-gather kernel: per-pixel kernel evaluation and per-batch partial sums in
                          FP32, folded into FP64 accumulators every 256 particles; footprints, pass A
                          (normalisation), the scatter kernel, 3D, HEALPix and the stencils stay FP64           */
} s2g_accumulate_mode;

typedef struct s2g_ctx s2g_ctx;

/* counters and device-side timings (CUDA events on the context stream) of the last call */
typedef struct {
    int64_t n_in;             /* particles handed in                                   */
    int64_t n_mapped;         /* particles with a non-empty footprint that were mapped */
    int64_t footprint_pixels; /* sum over mapped particles of the bounding-box pixels  */
    int64_t touched_pixels;   /* pixel updates with pix_weight != 0                    */
    int64_t n_fallback;       /* particles in the "no pixel centre covered" branch     */
    int64_t n_pairs;          /* (particle,tile) pairs of the gather path              */
    int64_t n_scatter;        /* particles deposited by the scatter kernel             */
    int64_t n_gather;         /* particles deposited by the gather kernel              */
    int64_t n_launches;       /* kernels launched by the library for this call          */
    /* device times [ms] from CUDA events on the context stream:
     * h2d / compute / d2h / total bracket the host entry points; prep (classify, scans, lists), sort (pair
     * expansion + radix sort + tile ranges), norm (pass A kernel), deposit (scatter + gather kernels),
     * epilogue (reduce_image) are per-phase sums inside `compute`. */
    double ms_h2d, ms_compute, ms_d2h, ms_total, ms_prep, ms_sort, ms_norm, ms_deposit, ms_epilogue;
} s2g_stats;

/* ---- context ------------------------------------------------------------------------------- */
S2G_API int s2g_device_count(void);
S2G_API int s2g_init(int device, s2g_ctx** out);
S2G_API int s2g_shutdown(s2g_ctx* ctx);
S2G_API int s2g_sync(s2g_ctx* ctx);
S2G_API const char* s2g_last_error(void);
S2G_API const char* s2g_version(void);
/* use an externally owned stream (e.g. torch's current stream, passed as cudaStream_t) */
S2G_API int s2g_set_stream(s2g_ctx* ctx, void* cuda_stream);
S2G_API int s2g_set_strategy(s2g_ctx* ctx, int strategy /* s2g_strategy */);
/* Pass A of the 2D deposit (calculate_weights, cic_2D.jl:11-72) sums w(u)*dA over the footprint.  For footprints
 * that are not clipped by the image and resolved by >= 20..64 pixels per kernel radius (kernel dependent) that sum
 * equals h^2 * ∫w(u) 2πu du to better than 5e-12 relative (tools/analytic_norm_study.py) and the closed form is
 * used by default.  on != 0 forces the numerical sum for every particle. */
S2G_API int s2g_set_exact_norm(s2g_ctx* ctx, int on);
/* optional FP32-accumulate mode (s2g_accumulate_mode); maps then agree with the FP64 result to 1e-5 per pixel
 * (plus 1e-9 of the plane maximum for pixels fed only by kernel-rim contributions). */
S2G_API int s2g_set_accumulate_mode(s2g_ctx* ctx, int mode /* s2g_accumulate_mode */);
S2G_API int s2g_get_stats(s2g_ctx* ctx, s2g_stats* out);
/* pinned host buffers for the end-to-end path */
S2G_API int s2g_host_alloc(void** out, uint64_t bytes);
S2G_API int s2g_host_free(void* p);
/* device buffers (for callers without their own CUDA allocator) */
S2G_API int s2g_dev_alloc(s2g_ctx* ctx, void** out, uint64_t bytes);
S2G_API int s2g_dev_free(s2g_ctx* ctx, void* p);
S2G_API int s2g_memcpy_h2d(s2g_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes);
S2G_API int s2g_memcpy_d2h(s2g_ctx* ctx, void* dst_host, const void* src_dev, uint64_t bytes);
S2G_API int s2g_memset_dev(s2g_ctx* ctx, void* dst_dev, int value, uint64_t bytes);

/* ---- Smac 2D deposit: replaces cic_mapping_2D (src/cic_interpolation/cic_2D.jl:103-244)
 *      incl. calculate_weights (:11-72), get_quantities_2D (:80-91) and the cic_shared.jl primitives.
 *      image_out: nx*ny x (n_images+1) doubles, overwritten (host variant) .
 *      accumulate != 0 (dev variant): add into image_dev instead of zero-filling it first. */
S2G_API int s2g_deposit_2d(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                           const void* binq, const void* w, int64_t n, int32_t n_images, int32_t in_dtype,
                           double len2pix, int64_t nx, int64_t ny, int32_t kernel, int32_t calc_mean,
                           double* image_out, s2g_stats* stats_or_null);
S2G_API int s2g_deposit_2d_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                               const void* binq, const void* w, int64_t n, int32_t n_images, int32_t in_dtype,
                               double len2pix, int64_t nx, int64_t ny, int32_t kernel, int32_t calc_mean,
                               int32_t accumulate, double* image_dev);

/* ---- Smac 2D deposit with a rotation measure per particle: replaces cic_mapping_2D(Pos,HSML,M,Rho,Bin_Q,Weights,RM;
 *      param,kernel,calc_mean,stokes) for RM !== nothing (cic_2D.jl:103-111, branch :129-131 / :201-217, and
 *      faraday_rotate_pixel! cic_shared.jl:129-159).  Particles are composited strictly in the order given (the
 *      caller passes them far -> near, cic_interpolation.jl:74-83); planes 1/2 of the image are Stokes Q/U.
 *      rm: n doubles (the reference's faraday_rotate_pixel! only accepts Float64).  stokes == 0 reproduces the
 *      reference too: RM is then read but nothing rotates (which is also all sphMapping ever asks for, because it
 *      does not forward `stokes`, cic_interpolation.jl:152-155). */
S2G_API int s2g_deposit_2d_rm(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                              const void* binq, const void* w, const double* rm, int64_t n, int32_t n_images,
                              int32_t in_dtype, double len2pix, int64_t nx, int64_t ny, int32_t kernel,
                              int32_t calc_mean, int32_t stokes, double* image_out, s2g_stats* stats_or_null);
S2G_API int s2g_deposit_2d_rm_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                  const void* binq, const void* w, const double* rm, int64_t n, int32_t n_images,
                                  int32_t in_dtype, double len2pix, int64_t nx, int64_t ny, int32_t kernel,
                                  int32_t calc_mean, int32_t stokes, double* image_dev);

/* ---- Smac 3D deposit: replaces cic_mapping_3D (src/cic_interpolation/cic_3D.jl:110-209).
 *      image: n^3 x 2 doubles (quantity plane, weight plane). */
S2G_API int s2g_deposit_3d(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                           const void* binq, const void* w, int64_t n, int32_t in_dtype, double len2pix, int64_t npix,
                           int32_t kernel, int32_t calc_mean, double* image_out, s2g_stats* stats_or_null);
S2G_API int s2g_deposit_3d_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                               const void* binq, const void* w, int64_t n, int32_t in_dtype, double len2pix,
                               int64_t npix, int32_t kernel, int32_t calc_mean, int32_t accumulate, double* image_dev);

/* ---- footprints only (bit-exact contract): pix_index_min_max (src/cic_interpolation/cic_shared.jl:46-52)
 *      after get_xyz (:85-100).  bounds_out: int64[2*dims*n] = {iMin,iMax,jMin,jMax[,kMin,kMax]} per particle. */
S2G_API int s2g_footprints(s2g_ctx* ctx, const void* pos, const void* hsml, int64_t n, int32_t in_dtype,
                           double len2pix, int64_t npix, int32_t dims, int64_t* bounds_out);

/* ---- reduce_image_2D / reduce_image_3D (src/cic_interpolation/reduce_image.jl:8-31, :39-55).
 *      2D: out is Julia Array{Float64,3}(nx,ny,n_images) memory: out[ix + nx*iy + nx*ny*q].
 *      3D: out is Array{Float64,3}(nz,ny,nx) memory (= flat order); when reduce_image == 0 the weight plane is
 *          taken as 1 (cic_interpolation.jl:230-232); division gated on the quantity plane > 0 (reduce_image.jl:49). */
S2G_API int s2g_reduce_image_2d(s2g_ctx* ctx, const double* image, int64_t nx, int64_t ny, int32_t n_images,
                                int32_t reduce_image, double* out);
S2G_API int s2g_reduce_image_2d_dev(s2g_ctx* ctx, const double* image_dev, int64_t nx, int64_t ny, int32_t n_images,
                                    int32_t reduce_image, double* out_dev);
S2G_API int s2g_reduce_image_3d(s2g_ctx* ctx, const double* image, int64_t npix, int32_t reduce_image, double* out);
S2G_API int s2g_reduce_image_3d_dev(s2g_ctx* ctx, const double* image_dev, int64_t npix, int32_t reduce_image,
                                    double* out_dev);

/* ---- centre + filter: center_particles (src/cic_interpolation/filter_shift.jl:6-32, arithmetic in the
 *      precision of pos, periodic wrap by boxsize/2 as the reference does) and filter_particles_in_image
 *      (:40-58).  pos_out may be NULL or == a separate buffer of 3*n elements (the Julia glue copies it back to
 *      reproduce the reference's in-place mutation); mask_out: uint8[n]. center/halfsize are those of the
 *      RECENTRED parameters for the filter (center = 0) and of the original ones for the shift. */
S2G_API int s2g_center_filter(s2g_ctx* ctx, const void* pos, int64_t n, int32_t in_dtype, const double shift[3],
                              int32_t periodic, double boxsize, const double filter_center[3],
                              const double filter_halfsize[3], void* pos_out, uint8_t* mask_out);

/* ---- fused sphMapping body (src/cic_interpolation/cic_interpolation.jl:35-273, parallel=false/true):
 *      centre (input precision) -> filter -> deposit -> reduce_image, all on the device.
 *      dims 2: out = Array(nx,ny,n_images) memory, or the flat both-maps buffer when return_both_maps != 0.
 *      dims 3: out = Array(n,n,n) memory (calc_mean is not forwarded in 3D, as in the reference :219-221).
 *      pos_recentred_out (optional, host): receives the recentred positions (the reference mutates Pos). */
S2G_API int s2g_sphmap(s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m, const void* rho,
                       const void* binq, const void* w, int64_t n, int32_t n_images, int32_t in_dtype,
                       const double shift[3], int32_t periodic, double boxsize, const double halfsize[3],
                       double len2pix, int64_t npix, int32_t kernel, int32_t calc_mean, int32_t reduce_image,
                       int32_t return_both_maps, void* pos_recentred_out, double* out, s2g_stats* stats_or_null);
/* device-resident variant used by the benchmark and the multi-GPU driver: leaves the FLAT image on the device
 * (so that partial images can be NCCL-reduced before reduce_image); positions are not written back. */
S2G_API int s2g_sphmap_dev(s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m,
                           const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                           int32_t in_dtype, const double shift[3], int32_t periodic, double boxsize,
                           const double halfsize[3], double len2pix, int64_t npix, int32_t kernel, int32_t calc_mean,
                           int32_t accumulate, double* image_dev);
/* ---- map_it's projection pre-step (cic_interpolation.jl:331-345) fused into the position load of the same call:
 *      perm (3 ints, or NULL): new component d = old component perm[d]; {0,2,1} = rotate_to_xz_plane!,
 *           {1,2,0} = rotate_to_yz_plane! (src/shared/rotate_particles.jl:35-73); exact, stays in the input precision.
 *      rot  (9 doubles row-major, or NULL): new = rot * old in Float64 = rotate_3D (rotate_particles.jl:7-13).
 *      shift/halfsize/len2pix describe the map in the ROTATED frame (rotate_to_xz_plane(par),
 *      src/shared/rotate_parameters.jl:27-59).  The caller's positions are only read (map_it works on a copy). */
S2G_API int s2g_sphmap_projected(s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m,
                                 const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                                 int32_t in_dtype, const int32_t* perm_or_null, const double* rot_or_null,
                                 const double shift[3], int32_t periodic, double boxsize, const double halfsize[3],
                                 double len2pix, int64_t npix, int32_t kernel, int32_t calc_mean,
                                 int32_t reduce_image, int32_t return_both_maps, void* pos_recentred_out, double* out,
                                 s2g_stats* stats_or_null);
S2G_API int s2g_sphmap_projected_dev(s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m,
                                     const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                                     int32_t in_dtype, const int32_t* perm_or_null, const double* rot_or_null,
                                     const double shift[3], int32_t periodic, double boxsize,
                                     const double halfsize[3], double len2pix, int64_t npix, int32_t kernel,
                                     int32_t calc_mean, int32_t accumulate, double* image_dev);

/* ---- HEALPix particle loop (src/healpix_interpolation/main.jl:143-213, pixel_weights.jl, constributing_pixels.jl).
 *      pos relative to the observer, already filtered (filter_sort_particles stays host logic).
 *      map_out / wmap_out: 12*nside^2 doubles each. */
S2G_API int s2g_healpix_deposit(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                const void* binq, const void* w, int64_t n, int32_t in_dtype, int64_t nside,
                                int32_t kernel, int32_t calc_mean, double* map_out, double* wmap_out,
                                s2g_stats* stats_or_null);
S2G_API int s2g_healpix_deposit_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                    const void* binq, const void* w, int64_t n, int32_t in_dtype, int64_t nside,
                                    int32_t kernel, int32_t calc_mean, int32_t accumulate, double* map_dev,
                                    double* wmap_dev);
/* ---- fused healpix_map body (src/healpix_interpolation/main.jl:92-227): `Pos .-= center`, shell filter and the
 *      far-to-near selection of filter_sort_particles (filter_particles.jl:17-54, including its `sorted[mask]`
 *      semantics) on the device, then the particle loop.  Float64 inputs only, like the reference (its methods are
 *      `where T` over homogeneous Float64 arrays).  pos_recentred_out (optional): the recentred positions, for the
 *      glue to reproduce the in-place mutation of Pos.  stats->n_in returns the number of particles in the shell.
 *      The `calc_mean=false` BoundsError of the reference (filter_particles.jl:28-30) is raised by the host glue. */
S2G_API int s2g_healpix_map(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                            const void* binq, const void* w, int64_t n, const double center[3],
                            const double radius_limits[2], int64_t nside, int32_t kernel, int32_t calc_mean,
                            void* pos_recentred_out, double* map_out, double* wmap_out, s2g_stats* stats_or_null);
S2G_API int s2g_healpix_map_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                const void* binq, const void* w, int64_t n, const double center[3],
                                const double radius_limits[2], int64_t nside, int32_t kernel, int32_t calc_mean,
                                int32_t accumulate, double* map_dev, double* wmap_dev, int64_t* n_selected);
/* pixel list of one particle (bit-exact contract vs the oracle): contributing_pixels (constributing_pixels.jl:7-22) */
S2G_API int s2g_healpix_pixels(s2g_ctx* ctx, const double pos[3], double radius, int64_t nside, int64_t* out,
                               int64_t cap, int64_t* count_out);

/* ---- CIC / TSC stencils (semantics: DESIGN.md; reference code is commented out, tsc_interpolation.jl:1-183).
 *      order 2 = CIC, 3 = TSC; image: n^dims x 2 doubles (field, weight). */
S2G_API int s2g_stencil_deposit(s2g_ctx* ctx, int32_t order, int32_t dims, const void* pos, const void* q, int64_t n,
                                int32_t in_dtype, double len2pix, int64_t npix, int32_t periodic, double* image_out,
                                s2g_stats* stats_or_null);
S2G_API int s2g_stencil_deposit_dev(s2g_ctx* ctx, int32_t order, int32_t dims, const void* pos, const void* q,
                                    int64_t n, int32_t in_dtype, double len2pix, int64_t npix, int32_t periodic,
                                    int32_t accumulate, double* image_dev);

/* ---- finite-guarded accumulation of partial maps (src/distributed_mapping/cic.jl:62-70, healpix.jl:44-52) */
S2G_API int s2g_accumulate_finite_dev(s2g_ctx* ctx, double* sum_dev, const double* local_dev, int64_t n);

/* ---- per-rank epilogue of the multi-process exchange (one process per GPU): after a reduce-scatter of the partial
 *      flat images (`image = sum(fetch.(futures))`, src/cic_interpolation/cic_interpolation.jl:199, :256) every rank
 *      holds the SUMMED planes of its own pixel slice and applies the reduce_image division to it
 *      (src/cic_interpolation/reduce_image.jl:8-31 for dims 2: q /= w where reduce_image and w > 0; :39-55 for dims 3:
 *      where q > 0, q /= (reduce_image ? w : 1)); the slices are then gathered and, in 2D, transposed by
 *      s2g_reduce_image_2d_dev(..., reduce_image = 0, ...).  q_slice: n_images planes of plane_stride elements. */
S2G_API int s2g_divide_slice_dev(s2g_ctx* ctx, int32_t dims, double* q_slice_dev, const double* w_slice_dev, int64_t n,
                                 int64_t plane_stride, int32_t n_images, int32_t reduce_image);

/* ---- device group: `parallel=true` of sphMapping (src/cic_interpolation/cic_interpolation.jl:171-215, 236-271) for a
 *      caller that is ONE process with several visible GPUs (a Julia session without Distributed workers).
 *      s2g_domain_decomposition  = domain_decomposition (src/parallel/domain_decomp.jl:7-17), 0-based starts; needs no
 *                                  device.  Slice r of it is what device r of a group deposits.
 *      s2g_group_init            one context per listed device (a device may be listed more than once); enables peer
 *                                access between all pairs.  s2g_group_peer_access() == 1: the partial images are summed
 *                                by direct peer loads; 0 (no P2P, or S2G_GROUP_NO_P2P=1): through peer copies.
 *      s2g_group_context         the context of rank r, e.g. for s2g_set_strategy / s2g_set_exact_norm.
 *      s2g_group_sphmap          = s2g_sphmap on the whole arrays: every device centres, filters and deposits its slice
 *                                into a private flat image (`@spawnat batch[i] cic_mapping_2D(...)`, :185-196), then
 *                                device r sums pixel slice r of ALL images in rank order (`sum(fetch.(futures))`, :199)
 *                                fused with the reduce_image epilogue (:212, :234) and writes its slice of `out`.
 *      s2g_group_healpix_map     = s2g_healpix_map on the whole arrays, incl. the `sorted[sel]` selection over ALL
 *                                particles (filter_particles.jl:33-41) — made once on device 0 when any particle is
 *                                outside the shell.
 *      stats_or_null: array of s2g_group_size() entries, one per device (ms_epilogue = peer sum + reduce_image). */
typedef struct s2g_group s2g_group;
S2G_API int s2g_domain_decomposition(int64_t n, int32_t n_parts, int64_t* starts_out, int64_t* counts_out);
S2G_API int s2g_group_init(const int32_t* devices, int32_t n_devices, s2g_group** out);
S2G_API int s2g_group_shutdown(s2g_group* grp);
S2G_API int s2g_group_size(const s2g_group* grp);
S2G_API int s2g_group_peer_access(const s2g_group* grp);
S2G_API int s2g_group_context(s2g_group* grp, int32_t rank, s2g_ctx** out);
S2G_API int s2g_group_sphmap(s2g_group* grp, int32_t dims, const void* pos, const void* hsml, const void* m,
                             const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                             int32_t in_dtype, const double shift[3], int32_t periodic, double boxsize,
                             const double halfsize[3], double len2pix, int64_t npix, int32_t kernel, int32_t calc_mean,
                             int32_t reduce_image, int32_t return_both_maps, void* pos_recentred_out, double* out,
                             s2g_stats* stats_or_null);
S2G_API int s2g_group_healpix_map(s2g_group* grp, const void* pos, const void* hsml, const void* m, const void* rho,
                                  const void* binq, const void* w, int64_t n, const double center[3],
                                  const double radius_limits[2], int64_t nside, int32_t kernel, int32_t calc_mean,
                                  void* pos_recentred_out, double* map_out, double* wmap_out, s2g_stats* stats_or_null);

/* ---- synthetic "Gadget-like" particle stream, generated on the device, counter-based
 *      (Philox4x32-10 keyed by seed, counter = global particle id) so any sharding sees the same particles.
 *      Writes particles [first_id, first_id+n) as SoA doubles (pos is 3xN interleaved). */
S2G_API int s2g_synth_particles_dev(s2g_ctx* ctx, uint64_t seed, int64_t first_id, int64_t n, int64_t n_total,
                                    double box, double n_ngb, double sigma_ln_rho, int32_t out_dtype, void* pos,
                                    void* hsml, void* m, void* rho, void* temp);

/* ---- roofline denominators measured live (bench.py): returns the achieved rate of a microbenchmark.
 *      which: 0 = FP64 DFMA [GFLOP/s], 1 = coalesced red.global.add.f64 [Gred/s] (32 consecutive doubles/warp,
 *      footprint `bytes`), 2 = random-address red.f64 [Gred/s], 3 = HBM copy [GB/s], 4 = shared-memory f64 atomics,
 *      5 / 6 = cp.reduce.async.bulk .add.f64 with 256-byte / 2-KiB operations [Gadd/s]. */
S2G_API int s2g_microbench(s2g_ctx* ctx, int32_t which, uint64_t bytes, int32_t iters, double* rate_out);

#ifdef __cplusplus
}
#endif
#endif /* SPHTOGRID_CUDA_H */
