/*
 * s2g_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, FP64 restatement of the particle-deposition hot path of
 * SPHtoGrid.jl v0.5.3.  Each function cites the reference file:line it
 * follows (paths relative to the reference checkout).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (libsphtogrid_cuda.so) never does.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off, so that no FMA is
 * contracted and the arithmetic is the reference's operation-by-operation).
 *
 * PARITY PINNING
 *   - Julia is not installed here, so the reference itself cannot be run.
 *   - Pinned against the reference's own self-contained known-answer tests
 *     (test/runtests.jl:38-99 params/filter/shift, :139-166 index bijection,
 *      :324-342 3D mass conservation, :718-738 2D fallback mass conservation)
 *     in tests/test_oracle_kat.py.
 *   - The third-party arithmetic the path calls is NOT vendored in the
 *     reference checkout and is restated from its published algorithm:
 *       SPHKernels.jl  ([compat] "2", Project.toml:53): kernel shape functions
 *       Healpix.jl     ([compat] "4", Project.toml:46): RING pixelisation
 *         (ang2pix, pix2vec, non-inclusive query_disc; = HEALPix C++
 *          healpix_base.cc algorithms, which Healpix.jl ports)
 *     For those two the status is "parity unpinned" (the only reference tests
 *     that would pin them need snapshots that are downloaded at test time).
 *     ang2pix_ring / pix2ang_ring / pix2vec_ring are additionally checked against
 *     the known answers printed in healpy's docstrings (Nside 16), i.e. against
 *     the standard HEALPix library (tests/test_oracle_healpix.py).
 *   - CIC/TSC stencils: the reference holds no live code (tsc_interpolation.jl
 *     is fully commented out) -> semantics defined here, "parity unpinned".
 *
 * Index conventions: everything here is 0-based; Julia's 1-based flat index
 * idx_julia = idx_c + 1, HEALPix pixel p_julia = p_c + 1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define S2GO_API __attribute__((visibility("default")))

enum {
    S2GO_CUBIC = 0,
    S2GO_QUINTIC = 1,
    S2GO_WC2 = 2,
    S2GO_WC4 = 3,
    S2GO_WC6 = 4,
    S2GO_WC8 = 5
};

/* ------------------------------------------------------------------------- */
/* SPHKernels.jl v2 (third-party, restated): W(kernel, u, h_inv)              */
/*   = norm(dim) * h_inv^dim * w(u);   call site cic_shared.jl:24             */
/* ------------------------------------------------------------------------- */
static double s2go_kernel_norm(int kid, int dim)
{
    const double pi = 3.14159265358979323846;
    switch (kid) {
    case S2GO_CUBIC:   return dim == 1 ? 4.0 / 3.0 : dim == 2 ? 40.0 / (7.0 * pi) : 8.0 / pi;
    case S2GO_QUINTIC: return dim == 1 ? 243.0 / 40.0 : dim == 2 ? 15309.0 / (478.0 * pi) : 2187.0 / (40.0 * pi);
    case S2GO_WC2:     return dim == 1 ? 5.0 / 4.0 : dim == 2 ? 7.0 / pi : 21.0 / (2.0 * pi);
    case S2GO_WC4:     return dim == 1 ? 3.0 / 2.0 : dim == 2 ? 9.0 / pi : 495.0 / (32.0 * pi);
    case S2GO_WC6:     return dim == 1 ? 55.0 / 32.0 : dim == 2 ? 78.0 / (7.0 * pi) : 1365.0 / (64.0 * pi);
    case S2GO_WC8:     return dim == 1 ? 1.0 : dim == 2 ? 8.0 / (3.0 * pi) : 357.0 / (64.0 * pi);
    default:           return 1.0;
    }
}

/* h_inv^dim as Julia's power_by_squaring evaluates it for dim = 1,2,3 */
static inline double s2go_pow_dim(double h_inv, int dim)
{
    if (dim == 1) return h_inv;
    if (dim == 2) return h_inv * h_inv;
    return (h_inv * h_inv) * h_inv;
}

static inline double s2go_pos(double x) { return x > 0.0 ? x : 0.0; }

/* shape function w(u), u in [0,1]; 0 for u >= 1 */
S2GO_API double s2go_kernel_shape(int kid, double u)
{
    if (!(u < 1.0)) return 0.0;
    double t = 1.0 - u;
    switch (kid) {
    case S2GO_CUBIC:
        if (u < 0.5) return 1.0 + 6.0 * (u - 1.0) * (u * u);
        return 2.0 * (t * t * t);
    case S2GO_QUINTIC: {
        double a = t, b = s2go_pos(2.0 / 3.0 - u), c = s2go_pos(1.0 / 3.0 - u);
        double a5 = (a * a) * (a * a) * a, b5 = (b * b) * (b * b) * b, c5 = (c * c) * (c * c) * c;
        return a5 - 6.0 * b5 + 15.0 * c5;
    }
    case S2GO_WC2: {
        double t2 = t * t;
        return (t2 * t2) * (1.0 + 4.0 * u);
    }
    case S2GO_WC4: {
        double t2 = t * t;
        return (t2 * t2 * t2) * (1.0 + 6.0 * u + (35.0 / 3.0) * (u * u));
    }
    case S2GO_WC6: {
        double t2 = t * t, t4 = t2 * t2;
        return (t4 * t4) * (1.0 + 8.0 * u + 25.0 * (u * u) + 32.0 * (u * u * u));
    }
    case S2GO_WC8: {
        double t2 = t * t, t4 = t2 * t2, u2 = u * u;
        return (t4 * t4 * t2) * (5.0 + 50.0 * u + 210.0 * u2 + 450.0 * (u2 * u) + 429.0 * (u2 * u2));
    }
    default:
        return 0.0;
    }
}

S2GO_API double s2go_kernel_value(int kid, int dim, double u, double h_inv)
{
    double n = s2go_kernel_norm(kid, dim) * s2go_pow_dim(h_inv, dim);
    return s2go_kernel_shape(kid, u) * n;
}

/* ------------------------------------------------------------------------- */
/* src/shared/indices.jl:6-17  (0-based result = Julia result - 1)            */
/* ------------------------------------------------------------------------- */
S2GO_API int64_t s2go_calculate_index_2d(int64_t i, int64_t j, int64_t x_pixels)
{
    return i * x_pixels + j;
}
S2GO_API int64_t s2go_calculate_index_3d(int64_t i, int64_t j, int64_t k, int64_t x_pixels, int64_t y_pixels)
{
    return i * x_pixels * y_pixels + j * y_pixels + k; /* sic: j*y_pixels (indices.jl:16) */
}

/* src/shared/distances.jl:6-17 */
static inline double get_d_hsml_2d(double dx, double dy, double hinv) { return sqrt(dx * dx + dy * dy) * hinv; }
static inline double get_d_hsml_3d(double dx, double dy, double dz, double hinv)
{
    return sqrt(dx * dx + dy * dy + dz * dz) * hinv;
}

/* src/cic_interpolation/cic_shared.jl:46-52 */
static inline void pix_index_min_max(double x, double hsml, int64_t n_pixels, int64_t* imin, int64_t* imax)
{
    int64_t lo = (int64_t)floor(x - hsml);
    int64_t hi = (int64_t)floor(x + hsml);
    *imin = lo > 0 ? lo : 0;
    *imax = hi < n_pixels - 1 ? hi : n_pixels - 1;
}

/* cic_shared.jl:60-76 */
static inline double get_dxyz(double x, double hsml, int64_t i)
{
    double a = x + hsml, b = (double)(i + 1);
    double c = x - hsml, d = (double)i;
    return (a < b ? a : b) - (c > d ? c : d);
}
static inline void get_x_dx(double x, double hsml, int64_t i, double* x_dist, double* dx)
{
    *dx = get_dxyz(x, hsml, i);
    *x_dist = x - (double)i - 0.5;
}

/* ------------------------------------------------------------------------- */
/* src/shared/parameters.jl:44-125                                            */
/* in : lims (use_lims=1) or center+sizes; pixelSideLength<0 / Npixels==0 mean */
/*      "not given" exactly like the -1.0 / 0 defaults of the reference        */
/* out: par[0..1]=x_lim par[2..3]=y_lim par[4..5]=z_lim par[6..8]=center       */
/*      par[9..11]=halfsize par[12]=len2pix par[13]=pixelSideLength            */
/*      *npix_out, returns 0 ok, 1 / 2 = the two reference error() cases       */
/* ------------------------------------------------------------------------- */
S2GO_API int s2go_mapping_parameters(const double x_lim_in[2], const double y_lim_in[2], const double z_lim_in[2],
                                     const double center_in[3], double x_size, double y_size, double z_size,
                                     double pixelSideLength, int64_t Npixels, double par[14], int64_t* npix_out)
{
    double x_lim[2] = {x_lim_in[0], x_lim_in[1]}, y_lim[2] = {y_lim_in[0], y_lim_in[1]};
    double z_lim[2] = {z_lim_in[0], z_lim_in[1]}, center[3] = {center_in[0], center_in[1], center_in[2]};
    int xl_def = (x_lim[0] == -1.0 && x_lim[1] == -1.0), yl_def = (y_lim[0] == -1.0 && y_lim[1] == -1.0);
    int zl_def = (z_lim[0] == -1.0 && z_lim[1] == -1.0);
    int c_def = (center[0] == -1.0 && center[1] == -1.0 && center[2] == -1.0);
    if (xl_def && yl_def && zl_def) { /* parameters.jl:58-69 */
        if (!c_def && (x_size != -1.0 && (y_size != -1.0 && z_size != -1.0))) {
            x_lim[0] = center[0] - 0.5 * x_size; x_lim[1] = center[0] + 0.5 * x_size;
            y_lim[0] = center[1] - 0.5 * y_size; y_lim[1] = center[1] + 0.5 * y_size;
            z_lim[0] = center[2] - 0.5 * z_size; z_lim[1] = center[2] + 0.5 * z_size;
        } else
            return 1;
    }
    if (x_size == -1.0) x_size = x_lim[1] - x_lim[0]; /* :72-80 */
    if (y_size == -1.0) y_size = y_lim[1] - y_lim[0];
    if (z_size == -1.0) z_size = z_lim[1] - z_lim[0];
    if (c_def) { /* :82-86 */
        center[0] = x_lim[0] + 0.5 * x_size;
        center[1] = y_lim[0] + 0.5 * y_size;
        center[2] = z_lim[0] + 0.5 * z_size;
    }
    double max_size = x_size > y_size ? x_size : y_size; /* :89 */
    if ((pixelSideLength == -1.0) && (Npixels != 0)) {   /* :91-99 */
        pixelSideLength = max_size / (double)Npixels;
    } else if ((pixelSideLength != -1.0) && (Npixels == 0)) {
        Npixels = (int64_t)floor(max_size / pixelSideLength);
        pixelSideLength = max_size / (double)Npixels;
    } else
        return 2;
    par[0] = x_lim[0]; par[1] = x_lim[1]; par[2] = y_lim[0]; par[3] = y_lim[1]; par[4] = z_lim[0]; par[5] = z_lim[1];
    par[6] = center[0]; par[7] = center[1]; par[8] = center[2];
    par[9] = 0.5 * x_size; par[10] = 0.5 * y_size; par[11] = 0.5 * z_size; /* :113 */
    par[12] = 1.0 / pixelSideLength;                                      /* :115 */
    par[13] = pixelSideLength;
    *npix_out = Npixels;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* src/cic_interpolation/filter_shift.jl:6-32  center_particles               */
/* in place, in the precision of Pos (Q1,Q2), periodic wrap by boxsize/2 (Q3) */
/* ------------------------------------------------------------------------- */
S2GO_API void s2go_center_particles_f64(double* pos, int64_t n, const double cen[3], int periodic, double boxsize)
{
    for (int64_t i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            double v = pos[3 * i + d] - cen[d];
            if (periodic) {
                if (fabs(v) > boxsize / 2) v = v > 0 ? v - boxsize / 2 : v + boxsize / 2;
            }
            pos[3 * i + d] = v;
        }
}
S2GO_API void s2go_center_particles_f32(float* pos, int64_t n, const double cen[3], int periodic, double boxsize)
{
    for (int64_t i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) {
            /* Float32 - Float64 promotes, the store rounds back to Float32 */
            float v = (float)((double)pos[3 * i + d] - cen[d]);
            if (periodic) {
                if (fabs((double)v) > boxsize / 2)
                    v = v > 0 ? (float)((double)v - boxsize / 2) : (float)((double)v + boxsize / 2);
            }
            pos[3 * i + d] = v;
        }
}

/* filter_shift.jl:40-58  (mask only; sort_z / Q5 is host logic, see s2go_sorted_mask_select) */
S2GO_API void s2go_filter_particles_f64(const double* pos, int64_t n, const double center[3], const double halfsize[3],
                                        uint8_t* mask)
{
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = center[d] - halfsize[d]; hi[d] = center[d] + halfsize[d]; }
    for (int64_t i = 0; i < n; i++) {
        int in = 1;
        for (int d = 0; d < 3; d++)
            if (!(lo[d] <= pos[3 * i + d] && pos[3 * i + d] <= hi[d])) in = 0;
        mask[i] = (uint8_t)in;
    }
}
S2GO_API void s2go_filter_particles_f32(const float* pos, int64_t n, const double center[3], const double halfsize[3],
                                        uint8_t* mask)
{
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = center[d] - halfsize[d]; hi[d] = center[d] + halfsize[d]; }
    for (int64_t i = 0; i < n; i++) {
        int in = 1;
        for (int d = 0; d < 3; d++) {
            double v = (double)pos[3 * i + d];
            if (!(lo[d] <= v && v <= hi[d])) in = 0;
        }
        mask[i] = (uint8_t)in;
    }
}

/* src/parallel/domain_decomp.jl:7-17  -> 0-based half-open [start,end) */
S2GO_API void s2go_domain_decomposition(int64_t n, int64_t n_workers, int64_t* start, int64_t* end)
{
    int64_t size = (int64_t)floor((double)n / (double)n_workers);
    for (int64_t i = 1; i <= n_workers - 1; i++) { start[i - 1] = (i - 1) * size; end[i - 1] = i * size; }
    start[n_workers - 1] = (n_workers - 1) * size;
    end[n_workers - 1] = n;
}

/* ------------------------------------------------------------------------- */
/* 2D deposit: src/cic_interpolation/cic_2D.jl:11-72 (calculate_weights),     */
/*             :80-91 (get_quantities_2D), :103-244 (cic_mapping_2D)          */
/* image: npix*npix rows x (n_images+1) planes, plane-separated (column       */
/* major), weight plane last; must be zero-filled by the caller (accumulates) */
/* binq : n_images x n column-major (or length n when n_images == 1)          */
/* fp   : optional int64[4*n] {iMin,iMax,jMin,jMax} per particle              */
/* ------------------------------------------------------------------------- */
typedef struct {
    int64_t n_mapped, footprint_pixels, touched_pixels, n_fallback;
} s2go_stats;

/* faraday_rotate_pixel! (cic_shared.jl:129-159).  image[idx,1] = Stokes Q, image[idx,2] = Stokes U (plane 2 is the
 * weight plane when only one quantity is mapped -- the reference does not check).  mod(x, pi) is Julia's
 * floating-point mod: r = rem(x, y) (exact, fmod); r == 0 -> copysign(r, y); sign(r) != sign(y) -> r + y.
 * psi = 0.5*atan(U/Q) is the ONE-argument arctangent (the quadrant of (Q,U) is lost: a pixel with Q < 0 comes back
 * with both signs flipped even for a zero rotation; Q = U = 0 gives NaN).  Reproduced as is. */
static inline double julia_mod_pi(double x)
{
    const double y = 3.141592653589793; /* Float64(pi) */
    double r = fmod(x, y);
    if (r == 0.0) return copysign(r, y);
    if ((r > 0.0) != (y > 0.0)) return r + y;
    return r;
}
static inline void faraday_rotate_pixel(double* image, int64_t idx, int64_t N_distr, double pRM, double pix_weight,
                                        int stokes)
{
    double _RM = pRM * pix_weight;
    _RM = julia_mod_pi(_RM);
    if (stokes) {
        double Q = image[idx], U = image[idx + N_distr];
        double Ipol = sqrt(Q * Q + U * U);
        double psi = 0.5 * atan(U / Q);
        image[idx] = Ipol * cos(2.0 * (psi + _RM));
        image[idx + N_distr] = Ipol * sin(2.0 * (psi + _RM));
    }
}

static void cic_mapping_2d_range(const double* pos, const double* hsml, const double* m, const double* rho,
                                 const double* binq, const double* w, int64_t p0, int64_t p1, int n_images,
                                 double len2pix, int64_t npix, int kid, int kdim, int calc_mean, double* image,
                                 double* wk, double* A, int64_t* fp, s2go_stats* st, const double* rm, int stokes,
                                 uint8_t* touched_pixel)
{
    const int64_t N_distr = npix * npix;
    for (int64_t p = p0; p < p1; p++) {
        /* cic_2D.jl:155-168 */
        int all_zero = 1;
        for (int q = 0; q < n_images; q++)
            if (binq[(int64_t)n_images * p + q] != 0.0) all_zero = 0;
        if (fp) { fp[4 * p] = 0; fp[4 * p + 1] = -1; fp[4 * p + 2] = 0; fp[4 * p + 3] = -1; }
        if (all_zero && !calc_mean) continue;

        /* get_quantities_2D cic_2D.jl:80-91 */
        double h = hsml[p] * len2pix;
        double hinv = 1.0 / h;
        double area = (2.0 * h) * (2.0 * h);
        double rho_p = rho[p] * (1.0 / (len2pix * len2pix * len2pix));
        double dz = m[p] / rho_p / area;
        double los_weight = w[p];

        /* get_xyz cic_shared.jl:85-100 */
        double x = pos[3 * p + 0] * len2pix;
        double y = pos[3 * p + 1] * len2pix;
        x += 0.5 * (double)npix;
        y += 0.5 * (double)npix;

        int64_t iMin, iMax, jMin, jMax;
        pix_index_min_max(x, h, npix, &iMin, &iMax);
        pix_index_min_max(y, h, npix, &jMin, &jMax);
        if (fp) { fp[4 * p] = iMin; fp[4 * p + 1] = iMax; fp[4 * p + 2] = jMin; fp[4 * p + 3] = jMax; }

        /* calculate_weights cic_2D.jl:11-72 */
        int64_t n_distr_pix = 0, n_tot_pix = 0;
        double distr_weight = 0.0, distr_area = 0.0;
        for (int64_t i = iMin; i <= iMax; i++) {
            double x_dist, dx;
            get_x_dx(x, h, i, &x_dist, &dx);
            for (int64_t j = jMin; j <= jMax; j++) {
                double y_dist, dy;
                get_x_dx(y, h, j, &y_dist, &dy);
                double u = get_d_hsml_2d(x_dist, y_dist, hinv);
                double dxdy = dx * dy;
                int64_t idx = s2go_calculate_index_2d(i, j, npix);
                /* get_weight_per_pixel cic_shared.jl:9-39 */
                A[idx] = dxdy;
                distr_area += dxdy;
                n_tot_pix += 1;
                if (u <= 1.0) {
                    double _wk = s2go_kernel_value(kid, kdim, u, hinv);
                    distr_weight += _wk * dxdy;
                    n_distr_pix += 1;
                    wk[idx] = _wk;
                } else
                    wk[idx] = 0.0;
            }
        }
        double weight_per_pix;
        if (distr_weight == 0.0) { /* cic_2D.jl:51-66 */
            n_distr_pix = n_tot_pix;
            for (int64_t i = iMin; i <= iMax; i++)
                for (int64_t j = jMin; j <= jMax; j++) wk[s2go_calculate_index_2d(i, j, npix)] = 1.0;
            if (distr_area != 0.0)
                weight_per_pix = (double)n_distr_pix / distr_area;
            else
                weight_per_pix = 1.0;
            if (st && n_tot_pix > 0) st->n_fallback++;
        } else
            weight_per_pix = (double)n_distr_pix / distr_weight;

        /* cic_2D.jl:186-188 */
        double kernel_norm = area / (double)n_distr_pix;
        double area_norm = kernel_norm * weight_per_pix * los_weight * dz;

        if (st && n_tot_pix > 0) { st->n_mapped++; st->footprint_pixels += n_tot_pix; }

        /* cic_2D.jl:193-222, update_image! cic_shared.jl:111-121 */
        for (int64_t i = iMin; i <= iMax; i++)
            for (int64_t j = jMin; j <= jMax; j++) {
                int64_t idx = s2go_calculate_index_2d(i, j, npix);
                double pix_weight = wk[idx] * A[idx] * area_norm;
                /* cic_2D.jl:201-209: Faraday-rotate what the pixel holds so far (only if something was deposited) */
                if (rm && touched_pixel[idx]) faraday_rotate_pixel(image, idx, N_distr, rm[p], pix_weight, stokes);
                if (pix_weight != 0.0) {
                    if (rm) touched_pixel[idx] = 1; /* cic_2D.jl:214-217 */
                    image[idx + N_distr * n_images] += pix_weight;
                    if (all_zero) /* bin_q collapsed to scalar 0.0 (cic_2D.jl:160-162) */
                        image[idx] += 0.0 * pix_weight;
                    else
                        for (int q = 0; q < n_images; q++)
                            image[idx + N_distr * q] += binq[(int64_t)n_images * p + q] * pix_weight;
                    if (st) st->touched_pixels++;
                }
            }
    }
}

S2GO_API int s2go_cic_mapping_2d(const double* pos, const double* hsml, const double* m, const double* rho,
                                 const double* binq, const double* w, int64_t n, int n_images, double len2pix,
                                 int64_t npix, int kid, int kdim, int calc_mean, double* image, int64_t* fp,
                                 int64_t* stats4)
{
    const int64_t N_distr = npix * npix;
    double* wk = (double*)calloc((size_t)N_distr, sizeof(double));
    double* A = (double*)malloc((size_t)N_distr * sizeof(double));
    if (!wk || !A) { free(wk); free(A); return -1; }
    s2go_stats st = {0, 0, 0, 0};
    cic_mapping_2d_range(pos, hsml, m, rho, binq, w, 0, n, n_images, len2pix, npix, kid, kdim, calc_mean, image, wk, A,
                         fp, &st, NULL, 0, NULL);
    if (stats4) { stats4[0] = st.n_mapped; stats4[1] = st.footprint_pixels; stats4[2] = st.touched_pixels; stats4[3] = st.n_fallback; }
    free(wk); free(A);
    return 0;
}

/* cic_mapping_2D with the RM argument (cic_2D.jl:103-244, RM !== nothing): particles are processed strictly in the
 * given order (the caller sorted them far -> near, cic_interpolation.jl:74-83); serial by construction. */
S2GO_API int s2go_cic_mapping_2d_rm(const double* pos, const double* hsml, const double* m, const double* rho,
                                    const double* binq, const double* w, const double* rm, int64_t n, int n_images,
                                    double len2pix, int64_t npix, int kid, int kdim, int calc_mean, int stokes,
                                    double* image, int64_t* stats4)
{
    const int64_t N_distr = npix * npix;
    double* wk = (double*)calloc((size_t)N_distr, sizeof(double));
    double* A = (double*)malloc((size_t)N_distr * sizeof(double));
    uint8_t* touched = (uint8_t*)calloc((size_t)N_distr, 1);
    if (!wk || !A || !touched) { free(wk); free(A); free(touched); return -1; }
    s2go_stats st = {0, 0, 0, 0};
    cic_mapping_2d_range(pos, hsml, m, rho, binq, w, 0, n, n_images, len2pix, npix, kid, kdim, calc_mean, image, wk, A,
                         NULL, &st, rm, stokes, touched);
    if (stats4) { stats4[0] = st.n_mapped; stats4[1] = st.footprint_pixels; stats4[2] = st.touched_pixels; stats4[3] = st.n_fallback; }
    free(wk); free(A); free(touched);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* 3D deposit: src/cic_interpolation/cic_3D.jl:13-78, :87-97, :110-209        */
/* image: npix^3 x 2 planes (quantity, weight)                                */
/* ------------------------------------------------------------------------- */
static void cic_mapping_3d_range(const double* pos, const double* hsml, const double* m, const double* rho,
                                 const double* binq, const double* w, int64_t p0, int64_t p1, double len2pix,
                                 int64_t npix, int kid, int kdim, int calc_mean, double* image, double* wk, double* V,
                                 int64_t* fp, s2go_stats* st, double* mass2)
{
    const int64_t N_distr = npix * npix * npix;
    double grid_mass = 0.0, particle_mass = 0.0;
    for (int64_t p = p0; p < p1; p++) {
        double bin_q = binq[p];
        if (fp) { fp[6 * p] = 0; fp[6 * p + 1] = -1; fp[6 * p + 2] = 0; fp[6 * p + 3] = -1; fp[6 * p + 4] = 0; fp[6 * p + 5] = -1; }
        if (bin_q == 0.0 && !calc_mean) continue;
        /* get_quantities_3D cic_3D.jl:87-97 */
        double h = hsml[p] * len2pix;
        double hinv = 1.0 / h;
        double rho_p = rho[p] / ((len2pix * len2pix) * len2pix);
        double vol = m[p] / rho_p;
        double los_weight = w[p];
        /* get_xyz */
        double x = pos[3 * p + 0] * len2pix, y = pos[3 * p + 1] * len2pix, z = pos[3 * p + 2] * len2pix;
        x += 0.5 * (double)npix; y += 0.5 * (double)npix; z += 0.5 * (double)npix;
        int64_t iMin, iMax, jMin, jMax, kMin, kMax;
        pix_index_min_max(x, h, npix, &iMin, &iMax);
        pix_index_min_max(y, h, npix, &jMin, &jMax);
        pix_index_min_max(z, h, npix, &kMin, &kMax);
        if (fp) { fp[6 * p] = iMin; fp[6 * p + 1] = iMax; fp[6 * p + 2] = jMin; fp[6 * p + 3] = jMax; fp[6 * p + 4] = kMin; fp[6 * p + 5] = kMax; }

        int64_t n_distr_pix = 0, n_tot_pix = 0;
        double distr_weight = 0.0, distr_volume = 0.0;
        for (int64_t i = iMin; i <= iMax; i++) {
            double x_dist, dx; get_x_dx(x, h, i, &x_dist, &dx);
            for (int64_t j = jMin; j <= jMax; j++) {
                double y_dist, dy; get_x_dx(y, h, j, &y_dist, &dy);
                for (int64_t k = kMin; k <= kMax; k++) {
                    double z_dist, dzz; get_x_dx(z, h, k, &z_dist, &dzz);
                    int64_t idx = s2go_calculate_index_3d(i, j, k, npix, npix);
                    double dxdydz = dx * dy * dzz;
                    double u = get_d_hsml_3d(x_dist, y_dist, z_dist, hinv);
                    V[idx] = dxdydz;
                    distr_volume += dxdydz;
                    n_tot_pix += 1;
                    if (u <= 1.0) {
                        double _wk = s2go_kernel_value(kid, kdim, u, hinv);
                        distr_weight += _wk * dxdydz;
                        n_distr_pix += 1;
                        wk[idx] = _wk;
                    } else
                        wk[idx] = 0.0;
                }
            }
        }
        double weight_per_pix;
        if (distr_weight == 0.0) {
            n_distr_pix = n_tot_pix;
            for (int64_t i = iMin; i <= iMax; i++)
                for (int64_t j = jMin; j <= jMax; j++)
                    for (int64_t k = kMin; k <= kMax; k++) wk[s2go_calculate_index_3d(i, j, k, npix, npix)] = 1.0;
            weight_per_pix = (distr_volume != 0.0) ? (double)n_distr_pix / distr_volume : 1.0;
            if (st && n_tot_pix > 0) st->n_fallback++;
        } else
            weight_per_pix = (double)n_distr_pix / distr_weight;

        /* cic_3D.jl:167-169 */
        double kernel_norm = vol / (double)n_distr_pix;
        double volume_norm = kernel_norm * weight_per_pix * los_weight * len2pix;
        if (st && n_tot_pix > 0) { st->n_mapped++; st->footprint_pixels += n_tot_pix; }

        for (int64_t i = iMin; i <= iMax; i++)
            for (int64_t j = jMin; j <= jMax; j++)
                for (int64_t k = kMin; k <= kMax; k++) {
                    int64_t idx = s2go_calculate_index_3d(i, j, k, npix, npix);
                    double pix_weight = wk[idx] * V[idx] * volume_norm;
                    if (pix_weight != 0.0) {
                        image[idx + N_distr] += pix_weight;
                        image[idx] += bin_q * pix_weight;
                        if (st) st->touched_pixels++;
                    }
                    grid_mass += rho[p] * wk[idx] * V[idx] / ((len2pix * len2pix) * len2pix); /* cic_3D.jl:186 */
                }
        particle_mass += m[p];
    }
    if (mass2) { mass2[0] += grid_mass; mass2[1] += particle_mass; }
}

S2GO_API int s2go_cic_mapping_3d(const double* pos, const double* hsml, const double* m, const double* rho,
                                 const double* binq, const double* w, int64_t n, double len2pix, int64_t npix, int kid,
                                 int kdim, int calc_mean, double* image, int64_t* fp, int64_t* stats4, double* mass2)
{
    const int64_t N_distr = npix * npix * npix;
    double* wk = (double*)calloc((size_t)N_distr, sizeof(double));
    double* V = (double*)malloc((size_t)N_distr * sizeof(double));
    if (!wk || !V) { free(wk); free(V); return -1; }
    s2go_stats st = {0, 0, 0, 0};
    if (mass2) { mass2[0] = 0.0; mass2[1] = 0.0; }
    cic_mapping_3d_range(pos, hsml, m, rho, binq, w, 0, n, len2pix, npix, kid, kdim, calc_mean, image, wk, V, fp, &st,
                         mass2);
    if (stats4) { stats4[0] = st.n_mapped; stats4[1] = st.footprint_pixels; stats4[2] = st.touched_pixels; stats4[3] = st.n_fallback; }
    free(wk); free(V);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* parallel=true: src/cic_interpolation/cic_interpolation.jl:171-215 (2D),    */
/* :236-271 (3D): contiguous slices (domain_decomposition), one private full  */
/* image per worker, image = sum(fetch.(futures)) in worker order.            */
/* This is the timed CPU baseline (OpenMP threads stand in for Julia workers).*/
/* ------------------------------------------------------------------------- */
S2GO_API int s2go_cic_mapping_parallel(int dims, const double* pos, const double* hsml, const double* m,
                                       const double* rho, const double* binq, const double* w, int64_t n, int n_images,
                                       double len2pix, int64_t npix, int kid, int kdim, int calc_mean, int n_workers,
                                       double* image, int64_t* stats4 /* may be NULL: counters summed over the slices */)
{
    if (n_workers < 1) n_workers = 1;
    s2go_stats* wst = (s2go_stats*)calloc((size_t)n_workers * 8, sizeof(s2go_stats)); /* stride 8: one cache line pair per worker */
    const int64_t N_distr = dims == 2 ? npix * npix : npix * npix * npix;
    const int planes = dims == 2 ? n_images + 1 : 2;
    int64_t* start = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_workers);
    int64_t* end = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_workers);
    double** partial = (double**)calloc((size_t)n_workers, sizeof(double*));
    s2go_domain_decomposition(n, n_workers, start, end);
    int fail = 0;
#pragma omp parallel for num_threads(n_workers) schedule(static, 1)
    for (int t = 0; t < n_workers; t++) {
        double* img = (double*)calloc((size_t)(N_distr * planes), sizeof(double));
        double* wk = (double*)calloc((size_t)N_distr, sizeof(double));
        double* A = (double*)malloc((size_t)N_distr * sizeof(double));
        if (!img || !wk || !A) {
            fail = 1;
        } else if (dims == 2)
            cic_mapping_2d_range(pos, hsml, m, rho, binq, w, start[t], end[t], n_images, len2pix, npix, kid, kdim,
                                 calc_mean, img, wk, A, NULL, &wst[8 * t], NULL, 0, NULL);
        else
            cic_mapping_3d_range(pos, hsml, m, rho, binq, w, start[t], end[t], len2pix, npix, kid, kdim, calc_mean, img,
                                 wk, A, NULL, &wst[8 * t], NULL);
        free(wk); free(A);
        partial[t] = img;
    }
    if (!fail) {
        /* sum(fetch.(futures)): ((p1 + p2) + p3) + ... element-wise */
#pragma omp parallel for num_threads(n_workers) schedule(static)
        for (int64_t e = 0; e < N_distr * planes; e++) {
            double s = partial[0][e];
            for (int t = 1; t < n_workers; t++) s += partial[t][e];
            image[e] = s;
        }
    }
    if (stats4) {
        stats4[0] = stats4[1] = stats4[2] = stats4[3] = 0;
        for (int t = 0; t < n_workers; t++) {
            stats4[0] += wst[8 * t].n_mapped; stats4[1] += wst[8 * t].footprint_pixels;
            stats4[2] += wst[8 * t].touched_pixels; stats4[3] += wst[8 * t].n_fallback;
        }
    }
    for (int t = 0; t < n_workers; t++) free(partial[t]);
    free(partial); free(start); free(end); free(wst);
    return fail ? -1 : 0;
}

S2GO_API int s2go_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* src/cic_interpolation/reduce_image.jl:8-31                                 */
/* out: Julia Array{Float64,3}(N,N,I) memory, i.e. out[ix + N*iy + N*N*n]     */
/*      = image[ix*N + iy, n] (/ weight where reduce && weight > 0)           */
/* ------------------------------------------------------------------------- */
S2GO_API void s2go_reduce_image_2d(const double* image, int64_t x_pixels, int64_t y_pixels, int n_images,
                                   int reduce_image, double* out)
{
    const int64_t N_distr = x_pixels * y_pixels;
    int64_t k = 0;
    for (int64_t i = 0; i < y_pixels; i++)
        for (int64_t j = 0; j < x_pixels; j++) {
            for (int q = 0; q < n_images; q++) {
                double v = image[k + N_distr * q];
                double wgt = image[k + N_distr * n_images];
                if (reduce_image && (wgt > 0.0)) v /= wgt;
                /* im_plot[j,i] then transposed -> final[i,j]; i = x index, j = y index */
                out[i + y_pixels * j + N_distr * q] = v;
            }
            k++;
        }
}

/* reduce_image.jl:39-55 (+ cic_interpolation.jl:230-232: image[:,2] .= 1 when !reduce_image) */
S2GO_API void s2go_reduce_image_3d(const double* image, int64_t npix, int reduce_image, double* out)
{
    const int64_t N_distr = npix * npix * npix;
    for (int64_t mm = 0; mm < N_distr; mm++) {
        double v = image[mm];
        double wgt = reduce_image ? image[mm + N_distr] : 1.0;
        if (image[mm] > 0.0) v /= wgt; /* Q7: gate on the quantity plane */
        out[mm] = v;
    }
}

/* distributed_mapping/cic.jl:62-70, healpix.jl:44-52: finite-guarded accumulate */
S2GO_API void s2go_accumulate_finite(double* sum, const double* local, int64_t n)
{
    for (int64_t i = 0; i < n; i++)
        if (!isnan(local[i]) && !isinf(local[i])) sum[i] += local[i];
}

/* ------------------------------------------------------------------------- */
/* HEALPix RING pixelisation (Healpix.jl v4, third-party, restated from the    */
/* HEALPix C/C++ algorithms it ports).  0-based pixel numbers.                 */
/* ------------------------------------------------------------------------- */
static const double S2GO_PI = 3.14159265358979323846;
static const double S2GO_TWOPI = 6.28318530717958647692;

/* ring index 1..4nside-1, first pixel, pixels in ring, shifted flag */
static inline void hp_ring_info(int64_t nside, int64_t ring, int64_t* startpix, int64_t* ringpix, int* shifted)
{
    int64_t npix = 12 * nside * nside, ncap = 2 * nside * (nside - 1);
    if (ring < nside) {
        *ringpix = 4 * ring; *startpix = 2 * ring * (ring - 1); *shifted = 1;
    } else if (ring <= 3 * nside) {
        *ringpix = 4 * nside; *startpix = ncap + (ring - nside) * 4 * nside; *shifted = (((ring - nside) & 1) == 0);
    } else {
        int64_t nr = 4 * nside - ring;
        *ringpix = 4 * nr; *startpix = npix - 2 * nr * (nr + 1); *shifted = 1;
    }
}

static inline int64_t hp_ring_above(int64_t nside, double z)
{
    double az = fabs(z);
    if (az <= 2.0 / 3.0) return (int64_t)((double)nside * (2.0 - 1.5 * z));
    int64_t iring = (int64_t)((double)nside * sqrt(3.0 * (1.0 - az)));
    return (z > 0) ? iring : 4 * nside - iring - 1;
}

static inline double hp_ring2z(int64_t nside, int64_t ring)
{
    double fact2 = 4.0 / (double)(12 * nside * nside);
    double fact1 = (double)(2 * nside) * fact2;
    if (ring < nside) return 1.0 - (double)(ring * ring) * fact2;
    if (ring <= 3 * nside) return (double)(2 * nside - ring) * fact1;
    ring = 4 * nside - ring;
    return (double)(ring * ring) * fact2 - 1.0;
}

S2GO_API int64_t s2go_hp_ang2pix_ring(int64_t nside, double theta, double phi)
{
    double z = cos(theta), za = fabs(z);
    double tt = fmod(phi, S2GO_TWOPI);
    if (tt < 0) tt += S2GO_TWOPI;
    tt = tt / (0.5 * S2GO_PI); /* in [0,4) */
    int64_t nl4 = 4 * nside, npix = 12 * nside * nside, ncap = 2 * nside * (nside - 1);
    if (za <= 2.0 / 3.0) {
        double temp1 = (double)nside * (0.5 + tt);
        double temp2 = (double)nside * z * 0.75;
        int64_t jp = (int64_t)floor(temp1 - temp2);
        int64_t jm = (int64_t)floor(temp1 + temp2);
        int64_t ir = nside + 1 + jp - jm; /* in {1,2n+1} */
        int64_t kshift = 1 - (ir & 1);
        int64_t ip = (jp + jm - nside + kshift + 1) / 2; /* in {0,4n-1} */
        ip = ((ip % nl4) + nl4) % nl4;
        return ncap + (ir - 1) * nl4 + ip;
    } else {
        double tp = tt - floor(tt);
        double tmp = (double)nside * sqrt(3.0 * (1.0 - za));
        int64_t jp = (int64_t)floor(tp * tmp);
        int64_t jm = (int64_t)floor((1.0 - tp) * tmp);
        int64_t ir = jp + jm + 1;
        int64_t ip = (int64_t)floor(tt * (double)ir);
        ip = ((ip % (4 * ir)) + 4 * ir) % (4 * ir);
        if (z > 0) return 2 * ir * (ir - 1) + ip;
        return npix - 2 * ir * (ir + 1) + ip;
    }
}

/* pix2ang_ring (classic formulation) */
S2GO_API void s2go_hp_pix2ang_ring(int64_t nside, int64_t pix, double* theta, double* phi)
{
    int64_t npix = 12 * nside * nside, ncap = 2 * nside * (nside - 1), nl2 = 2 * nside, nl4 = 4 * nside;
    double fact1 = 1.5 * (double)nside, fact2 = 3.0 * (double)nside * (double)nside;
    int64_t ipix1 = pix + 1;
    if (ipix1 <= ncap) {
        double hip = (double)ipix1 / 2.0;
        double fihip = floor(hip);
        int64_t iring = (int64_t)floor(sqrt(hip - sqrt(fihip))) + 1;
        int64_t iphi = ipix1 - 2 * iring * (iring - 1);
        *theta = acos(1.0 - (double)(iring * iring) / fact2);
        *phi = ((double)iphi - 0.5) * S2GO_PI / (2.0 * (double)iring);
    } else if (ipix1 <= nl2 * (5 * nside + 1)) {
        int64_t ip = ipix1 - ncap - 1;
        int64_t iring = ip / nl4 + nside;
        int64_t iphi = ip % nl4 + 1;
        double fodd = 0.5 * (double)(1 + ((iring + nside) & 1));
        *theta = acos((double)(nl2 - iring) / fact1);
        *phi = ((double)iphi - fodd) * S2GO_PI / (2.0 * (double)nside);
    } else {
        int64_t ip = npix - ipix1 + 1;
        double hip = (double)ip / 2.0;
        double fihip = floor(hip);
        int64_t iring = (int64_t)floor(sqrt(hip - sqrt(fihip))) + 1;
        int64_t iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
        *theta = acos(-1.0 + (double)(iring * iring) / fact2);
        *phi = ((double)iphi - 0.5) * S2GO_PI / (2.0 * (double)iring);
    }
}

S2GO_API void s2go_hp_pix2vec_ring(int64_t nside, int64_t pix, double v[3])
{
    double theta, phi;
    s2go_hp_pix2ang_ring(nside, pix, &theta, &phi);
    double st = sin(theta);
    v[0] = st * cos(phi);
    v[1] = st * sin(phi);
    v[2] = cos(theta);
}

S2GO_API void s2go_hp_vec2ang(double x, double y, double z, double* theta, double* phi)
{
    double norm = sqrt(x * x + y * y + z * z);
    *theta = acos(z / norm);
    double p = atan2(y, x);
    if (p < 0) p += S2GO_TWOPI;
    *phi = p;
}

/* non-inclusive query_disc, RING scheme (pixels whose CENTRE lies in the disc).
 * Appends 0-based pixels to out (capacity cap); returns the count (or -needed). */
S2GO_API int64_t s2go_hp_query_disc_ring(int64_t nside, double theta, double phi, double radius, int64_t* out,
                                         int64_t cap)
{
    int64_t npix = 12 * nside * nside;
    int64_t cnt = 0;
#define HP_APPEND_RANGE(a, b)                                                                                          \
    do {                                                                                                               \
        for (int64_t _p = (a); _p < (b); _p++) {                                                                       \
            if (cnt < cap) out[cnt] = _p;                                                                              \
            cnt++;                                                                                                     \
        }                                                                                                              \
    } while (0)
    double rsmall = radius, rbig = radius;
    if (rsmall >= S2GO_PI) { HP_APPEND_RANGE(0, npix); return cnt <= cap ? cnt : -cnt; }
    rbig = rbig < S2GO_PI ? rbig : S2GO_PI;
    double cosrbig = cos(rbig);
    double z0 = cos(theta);
    double xa = 1.0 / sqrt((1.0 - z0) * (1.0 + z0));
    double rlat1 = theta - rsmall;
    double zmax = cos(rlat1);
    int64_t irmin = hp_ring_above(nside, zmax) + 1;
    if ((rlat1 <= 0) && (irmin > 1)) { /* north pole in the disc */
        int64_t sp, rp; int sh;
        hp_ring_info(nside, irmin - 1, &sp, &rp, &sh);
        HP_APPEND_RANGE(0, sp + rp);
    }
    double rlat2 = theta + rsmall;
    double zmin = cos(rlat2);
    int64_t irmax = hp_ring_above(nside, zmin);
    for (int64_t iz = irmin; iz <= irmax; iz++) {
        double z = hp_ring2z(nside, iz);
        double x = (cosrbig - z * z0) * xa;
        double ysq = 1.0 - z * z - x * x;
        double dphi = (ysq <= 0) ? 0.0 : atan2(sqrt(ysq), x);
        if (dphi > 0) {
            int64_t ipix1, nr; int shifted;
            hp_ring_info(nside, iz, &ipix1, &nr, &shifted);
            double shift = shifted ? 0.5 : 0.0;
            int64_t ipix2 = ipix1 + nr - 1;
            int64_t ip_lo = (int64_t)floor((double)nr / S2GO_TWOPI * (phi - dphi) - shift) + 1;
            int64_t ip_hi = (int64_t)floor((double)nr / S2GO_TWOPI * (phi + dphi) - shift);
            if (ip_hi >= nr) { ip_lo -= nr; ip_hi -= nr; }
            if (ip_lo < 0) {
                HP_APPEND_RANGE(ipix1, ipix1 + ip_hi + 1);
                HP_APPEND_RANGE(ipix1 + ip_lo + nr, ipix2 + 1);
            } else
                HP_APPEND_RANGE(ipix1 + ip_lo, ipix1 + ip_hi + 1);
        }
    }
    if ((rlat2 >= S2GO_PI) && (irmax + 1 < 4 * nside)) { /* south pole in the disc */
        int64_t sp, rp; int sh;
        hp_ring_info(nside, irmax + 1, &sp, &rp, &sh);
        HP_APPEND_RANGE(sp, npix);
    }
#undef HP_APPEND_RANGE
    return cnt <= cap ? cnt : -cnt;
}

/* src/healpix_interpolation/constributing_pixels.jl:7-22: disc ∪ centre pixel, unique! (first occurrence kept) */
S2GO_API int64_t s2go_hp_contributing_pixels(int64_t nside, const double pos[3], double radius, int64_t* out,
                                             int64_t cap)
{
    double theta, phi;
    s2go_hp_vec2ang(pos[0], pos[1], pos[2], &theta, &phi);
    int64_t n = s2go_hp_query_disc_ring(nside, theta, phi, radius, out, cap - 1);
    if (n < 0) return n - 1;
    int64_t cpix = s2go_hp_ang2pix_ring(nside, theta, phi);
    out[n++] = cpix;
    /* unique!: the ring ranges are disjoint unless the disc wraps a whole ring; do the general thing */
    int64_t m = 0;
    for (int64_t a = 0; a < n; a++) {
        int dup = 0;
        /* ranges are sorted runs; a linear back-scan is only needed for the pushed centre pixel and wrap overlaps */
        if (a == n - 1 || (a > 0 && out[a] <= out[a - 1])) {
            for (int64_t b = 0; b < m; b++)
                if (out[b] == out[a]) { dup = 1; break; }
        }
        if (!dup) out[m++] = out[a];
    }
    return m;
}

/* ------------------------------------------------------------------------- */
/* HEALPix particle loop: src/healpix_interpolation/main.jl:143-213,           */
/* pixel_weights.jl:6-140, main.jl:25-63, shared.jl:1-10.                      */
/* pos is relative to the observer (already centred+filtered by the caller).   */
/* maps must be zero-filled by the caller; 0-based RING storage.               */
/* ------------------------------------------------------------------------- */
S2GO_API int s2go_healpix_deposit(const double* pos, const double* hsml, const double* m, const double* rho,
                                  const double* binq, const double* w, int64_t n, int64_t nside, int kid, int kdim,
                                  int calc_mean, double* allsky_map, double* weight_map, int64_t* stats4)
{
    const int64_t npix = 12 * nside * nside;
    const double ang_pix = sqrt(4.0 * S2GO_PI / (double)npix); /* main.jl:144 */
    int64_t cap = 1024;
    int64_t* pixidx = (int64_t*)malloc(sizeof(int64_t) * (size_t)cap);
    double* wk = (double*)malloc(sizeof(double) * (size_t)cap);
    double* A = (double*)malloc(sizeof(double) * (size_t)cap);
    s2go_stats st = {0, 0, 0, 0};
    for (int64_t ip = 0; ip < n; ip++) {
        if (!calc_mean && binq[ip] == 0.0) continue; /* main.jl:160-165 */
        const double* P = pos + 3 * ip;
        /* get_norm shared.jl:1-10 */
        double dx2 = 0.0;
        for (int d = 0; d < 3; d++) dx2 += P[d] * P[d];
        double Dx = sqrt(dx2);
        if (Dx < hsml[ip]) continue; /* main.jl:172-174 */
        double proj_hsml = asin(hsml[ip] / Dx); /* :182 */
        int64_t np;
        for (;;) {
            np = s2go_hp_contributing_pixels(nside, P, proj_hsml, pixidx, cap);
            if (np >= 0) break;
            cap = -np + 16;
            pixidx = (int64_t*)realloc(pixidx, sizeof(int64_t) * (size_t)cap);
            wk = (double*)realloc(wk, sizeof(double) * (size_t)cap);
            A = (double*)realloc(A, sizeof(double) * (size_t)cap);
        }
        /* particle_area_and_depth main.jl:56-63, :193 */
        double dz = 2.0 * hsml[ip];
        double area = (m[ip] / rho[ip]) / dz;
        dz /= (ang_pix * Dx) * (ang_pix * Dx);
        /* calculate_weights pixel_weights.jl:87-140 */
        double hsml_inv = 1.0 / proj_hsml;
        int64_t n_distr_pix = 0, n_tot_pix = 0;
        double distr_weight = 0.0, distr_area = 0.0;
        for (int64_t k = 0; k < np; k++) {
            /* weight_per_index pixel_weights.jl:34-76 */
            double c[3];
            s2go_hp_pix2vec_ring(nside, pixidx[k], c);
            double d = 0.0;
            for (int q = 0; q < 3; q++) d += P[q] * c[q];
            double t = d / Dx;
            double ddx = acos(t < 1.0 ? t : 1.0);
            double u = ddx * hsml_inv;
            double inner = fabs(proj_hsml - (ddx - 0.5 * ang_pix));
            double mn = ang_pix < inner ? ang_pix : inner;
            double _A = (0.0 > mn ? 0.0 : mn) / ang_pix;
            _A /= (ang_pix * Dx) * (ang_pix * Dx);
            distr_area += _A;
            n_tot_pix += 1;
            double _wk;
            if (u <= 1.0) {
                _wk = s2go_kernel_value(kid, kdim, u, hsml_inv);
                distr_weight += _wk * _A;
                n_distr_pix += 1;
            } else
                _wk = 0.0;
            A[k] = _A;
            wk[k] = _wk;
        }
        double weight_per_pix;
        if (distr_weight == 0.0) {
            n_distr_pix = n_tot_pix;
            for (int64_t k = 0; k < np; k++) wk[k] = 1.0;
            weight_per_pix = (distr_area != 0.0) ? (double)n_distr_pix / distr_area : 1.0;
            st.n_fallback++;
        } else
            weight_per_pix = (double)n_distr_pix / distr_weight;
        /* update_image! main.jl:25-45 */
        double kernel_norm = area / (double)n_distr_pix;
        double area_norm = kernel_norm * weight_per_pix * w[ip] * dz;
        for (int64_t k = 0; k < np; k++) {
            double pix_weight = area_norm * wk[k] * A[k];
            allsky_map[pixidx[k]] += binq[ip] * pix_weight;
            weight_map[pixidx[k]] += pix_weight;
        }
        st.n_mapped++;
        st.footprint_pixels += np;
        st.touched_pixels += np;
    }
    if (stats4) { stats4[0] = st.n_mapped; stats4[1] = st.footprint_pixels; stats4[2] = st.touched_pixels; stats4[3] = st.n_fallback; }
    free(pixidx); free(wk); free(A);
    return 0;
}

/* OpenMP slices + private maps + ordered sum, mirroring how distributed_allsky_map sums per-worker maps */
S2GO_API int s2go_healpix_deposit_parallel(const double* pos, const double* hsml, const double* m, const double* rho,
                                           const double* binq, const double* w, int64_t n, int64_t nside, int kid,
                                           int kdim, int calc_mean, int n_workers, double* allsky_map,
                                           double* weight_map)
{
    if (n_workers < 1) n_workers = 1;
    const int64_t npix = 12 * nside * nside;
    int64_t* start = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_workers);
    int64_t* end = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_workers);
    double** pm = (double**)calloc((size_t)n_workers, sizeof(double*));
    double** pw = (double**)calloc((size_t)n_workers, sizeof(double*));
    s2go_domain_decomposition(n, n_workers, start, end);
#pragma omp parallel for num_threads(n_workers) schedule(static, 1)
    for (int t = 0; t < n_workers; t++) {
        pm[t] = (double*)calloc((size_t)npix, sizeof(double));
        pw[t] = (double*)calloc((size_t)npix, sizeof(double));
        int64_t s = start[t], cntp = end[t] - start[t];
        s2go_healpix_deposit(pos + 3 * s, hsml + s, m + s, rho + s, binq + s, w + s, cntp, nside, kid, kdim, calc_mean,
                             pm[t], pw[t], NULL);
    }
#pragma omp parallel for num_threads(n_workers) schedule(static)
    for (int64_t e = 0; e < npix; e++) {
        double a = pm[0][e], b = pw[0][e];
        for (int t = 1; t < n_workers; t++) { a += pm[t][e]; b += pw[t][e]; }
        allsky_map[e] = a; weight_map[e] = b;
    }
    for (int t = 0; t < n_workers; t++) { free(pm[t]); free(pw[t]); }
    free(pm); free(pw); free(start); free(end);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* CIC / TSC stencils.  NO live reference code (tsc_interpolation.jl:1-183 is  */
/* commented out and its dependency is not in Project.toml) -> semantics       */
/* defined here ("parity unpinned"):                                           */
/*   grid coordinate g = pos*len2pix + 0.5*npix (same get_xyz convention as    */
/*   the Smac path, so cell c covers [c, c+1) with centre c+0.5);              */
/*   CIC: multilinear weights on the 2^d cells nearest to g-0.5;               */
/*   TSC: quadratic-spline weights on 3^d cells around the cell containing g:  */
/*        d = g - (c+0.5); w0 = 0.75 - d^2, w(+-1) = 0.5*(0.5 +- d)^2;         */
/*   cells outside the grid are dropped (non-periodic) or wrapped (periodic);  */
/*   field plane += q*w, weight plane += w; average=true => field/weight where */
/*   weight > 0 (done by s2go_stencil_average).                                */
/* layout: same flat index as the Smac path (indices.jl), planes [field, wgt]. */
/* ------------------------------------------------------------------------- */
static inline int wrap_or_drop(int64_t* c, int64_t n, int periodic)
{
    if (*c >= 0 && *c < n) return 1;
    if (!periodic) return 0;
    *c = ((*c % n) + n) % n;
    return 1;
}

S2GO_API void s2go_stencil_deposit(int order /*2=CIC,3=TSC*/, int dims, const double* pos, const double* q, int64_t n,
                                   double len2pix, int64_t npix, int periodic, double* image)
{
    const int64_t N_distr = dims == 2 ? npix * npix : npix * npix * npix;
    for (int64_t p = 0; p < n; p++) {
        double g[3];
        int64_t c0[3];
        double wgt[3][3];
        int cnt = order;
        for (int d = 0; d < 3; d++) {
            g[d] = pos[3 * p + d] * len2pix;
            g[d] += 0.5 * (double)npix;
            if (order == 2) {
                double s = g[d] - 0.5;
                double f = floor(s);
                double fr = s - f;
                c0[d] = (int64_t)f;
                wgt[d][0] = 1.0 - fr; wgt[d][1] = fr; wgt[d][2] = 0.0;
            } else {
                double f = floor(g[d]);
                double dd = g[d] - (f + 0.5);
                c0[d] = (int64_t)f - 1;
                wgt[d][0] = 0.5 * ((0.5 - dd) * (0.5 - dd));
                wgt[d][1] = 0.75 - dd * dd;
                wgt[d][2] = 0.5 * ((0.5 + dd) * (0.5 + dd));
            }
        }
        for (int a = 0; a < cnt; a++) {
            int64_t i = c0[0] + a;
            if (!wrap_or_drop(&i, npix, periodic)) continue;
            for (int b = 0; b < cnt; b++) {
                int64_t j = c0[1] + b;
                if (!wrap_or_drop(&j, npix, periodic)) continue;
                if (dims == 2) {
                    double ww = wgt[0][a] * wgt[1][b];
                    int64_t idx = s2go_calculate_index_2d(i, j, npix);
                    image[idx] += q[p] * ww;
                    image[idx + N_distr] += ww;
                } else
                    for (int c = 0; c < cnt; c++) {
                        int64_t k = c0[2] + c;
                        if (!wrap_or_drop(&k, npix, periodic)) continue;
                        double ww = wgt[0][a] * wgt[1][b] * wgt[2][c];
                        int64_t idx = s2go_calculate_index_3d(i, j, k, npix, npix);
                        image[idx] += q[p] * ww;
                        image[idx + N_distr] += ww;
                    }
            }
        }
    }
}

S2GO_API void s2go_stencil_average(const double* image, int64_t n_cells, double* out)
{
    for (int64_t e = 0; e < n_cells; e++) {
        double wv = image[e + n_cells];
        out[e] = wv > 0.0 ? image[e] / wv : image[e];
    }
}
