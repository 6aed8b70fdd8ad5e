// s2g_hpgather.cu — HEALPix deposit, tile-gather strategy: pass B of the healpix_map particle loop
// (update_image!, src/healpix_interpolation/main.jl:25-45, with the weights of weight_per_index,
// pixel_weights.jl:34-76) WITHOUT global atomics in the inner loop.
//
// The scatter walk (s2g_healpix.cu) issues two red.global.add.f64 per (particle, pixel): 2.3e12 reds for BASELINE
// config 4 — 8.6 s at the measured L2 red peak before a single FP64 instruction is counted.  Here the sphere is cut
// into TILES = (band of HPG_BR consecutive rings) x (sector of <= HPG_SW consecutive pixels of each of those rings); a
// CTA owns a tile for the duration of a work item, every thread owns HPG_PPT pixels of ONE ring and keeps
//   * the unit vectors of its pixel centres (pix2vecRing, pixel_weights.jl:42) and
//   * the weight and quantity sums of those pixels
// in registers while the particle records stream through shared memory.  Per (pixel, particle) the angular distance is
// the chord between two unit vectors (three subtractions, two FMAs) — the per-pixel sin/cos of the scatter walk is gone —
// then dx = 2 asin(chord/2) by its series, the kernel polynomial and two FMAs.  One coalesced red.add flush per work item.
//
//   pass A (calculate_weights, pixel_weights.jl:87-140)  stays the ring walk of s2g_healpix.cu (REC mode): it needs the
//          EXACT pixel list (n_distr, the centre pixel, the `distr_weight == 0` branch) and no atomics; it writes one
//          64-byte record per particle.  Particles in the fallback branch or with a non-finite normalisation are handed
//          back to the scatter walk.
//   pairs  (tile, record) for every tile that holds a pixel of the particle's disc — from the same query_disc ring
//          runs as the reference's pixel list (ring_run), so no tile is missed and none is listed in vain;
//          stable CUB radix sort by tile (deterministic order inside a tile), tile ranges, chunks of <= 4096 pairs.
//   pass B k_hp_gather below.  Membership is geometric (t = 1 - dx/proj_hsml > 0): queryDiscRing is "pixel centre inside
//          the disc", so the two sets differ only where w -> 0.
//
// Sector rule: ring r of band b has L_r pixels and the band ns_b = ceil(max_r L_r / HPG_SW) sectors; pixel j of ring r
// belongs to sector floor(j * ns_b / L_r), i.e. sector k owns j in [ceil(k L_r/ns_b), ceil((k+1) L_r/ns_b)) — at most
// HPG_SW pixels, contiguous, and the same rule in the polar caps (L_r = 4r) and the belt (L_r = 4 Nside).
#include <cub/cub.cuh>
#include <cstdlib>

#include "s2g_healpix.cuh"

namespace {

constexpr int HPG_BR = 16;        // rings per band (= tile height)
constexpr int HPG_SW = 64;        // max pixels of a ring per sector (= tile width)
constexpr int HPG_PPT = 4;        // pixels per thread: slots lane16 + 16 m of the thread's ring
constexpr int HPG_THREADS = 256;  // 16 rings x 16 lanes
constexpr int HPG_CTAS = 3;
constexpr int HPG_BATCH = 256;    // records staged in shared memory at a time
constexpr int HPG_CHUNK = 4096;   // max pairs per work item

// record as staged in shared memory: everything the inner loop needs, derived once per (pair) by the staging thread
struct __align__(16) HRecS {
    double ux, uy, uz;   // unit vector to the particle
    double c2max;        // squared chord of the disc rim, 4 sin^2(proj_h / 2)
    double hinv;         // 1 / proj_h
    double php;          // proj_h / ang_pix
    double an, anq;      // area_norm / (ang_pix Dx)^2  and the same times the quantity
    int rmin, rmax;      // disc rings
    int big, pad1;       // big: proj_h + 2 ang_pix >= 0.2 rad -> asin itself, not its series
};

// asin(x)/x = 1 + x^2/6 + 3x^4/40 + ... written in c2 = (2x)^2 (the squared chord): G(c2) = sum kG[i] c2^i
__constant__ double kG[8] = {1.0,
                             1.0 / 6.0 / 4.0,
                             3.0 / 40.0 / 16.0,
                             15.0 / 336.0 / 64.0,
                             105.0 / 3456.0 / 256.0,
                             945.0 / 42240.0 / 1024.0,
                             10395.0 / 599040.0 / 4096.0,
                             135135.0 / 9676800.0 / 16384.0};

// ---- classification: which path deposits particle p
//   heavy[p]      very large disc the gather cannot take (non-finite quantity or normalisation, >= 1.5 rad)
//                 -> cooperative scatter launch
//   gath[p]       disc for the tile-gather: 1 = below 0.073 rad (short series), 2 = up to 0.2 rad
//   gath_heavy[p] disc above 0.2 rad for the tile-gather (asin instead of its series)
//   skip[p] = 1   the ordinary scatter launch must NOT take it (any of the three above)
__global__ void __launch_bounds__(256) k_hp_classify(s2g_particles P, HpGeom g, int calc_mean,
                                                     const unsigned char* __restrict__ take, double heavy_radius,
                                                     double gather_radius, int gather_on,
                                                     unsigned char* __restrict__ heavy, unsigned char* __restrict__ gath,
                                                     unsigned char* __restrict__ gath_heavy,
                                                     unsigned char* __restrict__ skip)
{
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const double x = ld_pos(P, p, 0), y = ld_pos(P, p, 1), z = ld_pos(P, p, 2);
    const double dx = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    const double hs = ld_in(P.hsml, p, P.in_dtype);
    const double q = ld_in(P.binq, p, P.in_dtype);
    unsigned char h = 0, ga = 0, gh = 0;
    const bool alive = (!take || take[p]) && (calc_mean || q != 0.0) && (dx >= hs);
    if (alive) {
        // asin(hs/dx) >= r  <=>  hs >= dx*sin(r)  (r < pi/2)
        h = (heavy_radius > 0.0 && hs >= dx * sin(heavy_radius)) ? 1 : 0;
        if (gather_on) {
            const double ph = asin(__ddiv_rn(hs, dx));
            const double m = 2.0 * g.ang_pix;
            const double an_probe = ld_in(P.m, p, P.in_dtype) / ld_in(P.rho, p, P.in_dtype) * ld_in(P.w, p, P.in_dtype);
            const bool ok = ph >= gather_radius && ph < 1.5 && isfinite(q) && isfinite(an_probe) && an_probe != 0.0;
            if (ok) {
                // ga = 1: small-angle disc (5 series coefficients are exact to 1e-16 below 0.073 rad), 2: up to 0.2 rad
                // (8 coefficients); gh: above — asin itself (BIG instantiation, 2 CTAs/SM)
                if (ph + m >= 0.2) gh = 1; else ga = (ph + m < 0.073) ? 1 : 2;
                h = 0;
            }
        }
    }
    heavy[p] = h;
    gath[p] = ga;
    gath_heavy[p] = gh;
    skip[p] = (h || ga || gh) ? 1 : 0;
}

// ---- tile table of a resolution
struct HpTiles {
    int nbands, ntiles;
    const int* band_base;  // [nbands+1] first tile id of a band
    const int* band_ns;    // [nbands]   sectors of a band
};

// ---- records without a ring walk: everything of main.jl:143-193 that does not need the pixel list.  `an` holds the
// UN-normalised area·w·dz/(ang_pix Δx)² until k_hp_normalise divides it by Σ wk·A' (pass A as a tile-gather, below):
// area_norm = kernel_norm·wpp·w·dz = (area/N)(N/Σ)·w·dz, N cancels (main.jl:32-33, pixel_weights.jl:119-137).
__global__ void __launch_bounds__(256) k_hp_records(s2g_particles P, HpGeom g, const unsigned* __restrict__ list,
                                                    long long n_list, HRec* __restrict__ recs)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_list) return;
    const long long p = list[t];
    Disc d;
    d.px = ld_pos(P, p, 0); d.py = ld_pos(P, p, 1); d.pz = ld_pos(P, p, 2);
    const double hs = ld_in(P.hsml, p, P.in_dtype);
    d.Dx = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d.px, d.px), __dmul_rn(d.py, d.py)), __dmul_rn(d.pz, d.pz)));
    d.proj_h = asin(__ddiv_rn(hs, d.Dx));
    d.hinv = __ddiv_rn(1.0, d.proj_h);
    make_disc(g, d);
    HRec r;
    r.ux = d.px / d.Dx; r.uy = d.py / d.Dx; r.uz = d.pz / d.Dx;
    r.ph = d.proj_h;
    // particle_area_and_depth (main.jl:56-63) and :193
    double dz = __dmul_rn(2.0, hs);
    const double area = __ddiv_rn(__ddiv_rn(ld_in(P.m, p, P.in_dtype), ld_in(P.rho, p, P.in_dtype)), dz);
    const double aD = __dmul_rn(g.ang_pix, d.Dx);
    dz = __ddiv_rn(dz, __dmul_rn(aD, aD));
    // pix_weight = area_norm·wk·A with A = A'/(ang_pix Δx)² (pixel_weights.jl:53) and area_norm = area·w·dz/Σ(wk·A):
    // the two (ang_pix Δx)² cancel, pix_weight = (area·w·dz / Σ wk·A')·wk·A'
    r.an = area * ld_in(P.w, p, P.in_dtype) * dz;
    r.anq = ld_in(P.binq, p, P.in_dtype);          // the quantity; becomes an*q in k_hp_normalise
    // all rings of the walk: the disc's rings [irmin, irmax] plus, when a pole lies inside the disc, the whole rings
    // between that pole and the disc's first / last ring (query_disc appends them entirely; they are geometrically
    // inside the disc, so the gather's membership test agrees)
    const bool ok = !d.full_sky && d.ring_first <= d.ring_last;
    r.rmin = ok ? (int)d.ring_first : 1;
    r.rmax = ok ? (int)d.ring_last : 0;
    r.ntot = 0;
    r.pad = (int)p;
    recs[t] = r;
}

// after pass A: an := an / Σ, anq := an·q.  A record whose Σ is zero (`distr_weight == 0` branch, pixel_weights.jl:121)
// or not finite, or whose centre pixel is not in its own disc walk (ntot < 0, set by k_hp_pairs), is dropped here and
// handed back to the scatter walk through skip[p] = 0.
__global__ void __launch_bounds__(256) k_hp_normalise(HRec* __restrict__ recs, const double* __restrict__ S, long long n,
                                                      unsigned char* __restrict__ skip)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    HRec r = recs[t];
    if (r.rmin > r.rmax) { skip[r.pad] = 0; return; }
    const double sw = S[t];
    const double an = r.an / sw;
    if (!(sw > 0.0) || !isfinite(an) || r.ntot < 0) {
        recs[t].rmin = 1; recs[t].rmax = 0; recs[t].ntot = 0;
        skip[r.pad] = 0;
        return;
    }
    recs[t].an = an;
    recs[t].anq = an * r.anq;
}

// ---- (tile, record) pairs.  One warp per record; half-warps take alternate bands, lanes are the rings of a band.
// The sector range of a ring is that of its query_disc pixel run (ring_run: the reference's own list), the band's
// range the union over its rings, expressed relative to the sector holding the disc centre (the runs are intervals
// around the particle's azimuth, possibly wrapping).
// The count pass parks (first sector, sector count) of every band of a record in `bandinfo` (HPG_NBMAX slots per record),
// so that the write pass of a record with at most HPG_NBMAX bands (every non-heavy disc) copies them out instead of
// evaluating the ring runs (an atan2 per ring) a second time.
constexpr int HPG_NBMAX = 40;

template <bool WRITE>
__global__ void __launch_bounds__(256) k_hp_pairs(HRec* __restrict__ recs, long long n_rec, HpGeom g, HpTiles T,
                                                  const unsigned* __restrict__ off, unsigned* __restrict__ npairs,
                                                  unsigned* __restrict__ keys, unsigned* __restrict__ vals, int fill_ntot,
                                                  unsigned* __restrict__ bandinfo)
{
    const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
    const long long t = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (t >= n_rec) return;
    const HRec r = recs[t];
    if (r.rmin > r.rmax) {
        if (!WRITE && lane == 0) npairs[t] = 0;
        return;
    }
    if (WRITE && bandinfo && (r.rmax - 1) / HPG_BR - (r.rmin - 1) / HPG_BR < HPG_NBMAX) {
        // replay the parked band ranges: lanes over the sectors of a band
        const int b0 = (r.rmin - 1) / HPG_BR, b1 = (r.rmax - 1) / HPG_BR;
        unsigned o = off[t];
        for (int band = b0; band <= b1; ++band) {
            const unsigned bi = bandinfo[t * HPG_NBMAX + (band - b0)];
            const int sec0 = (int)(bi >> 16), nsec = (int)(bi & 0xffffu), ns = T.band_ns[band], base = T.band_base[band];
            for (int k = lane; k < nsec; k += 32) {
                int sec = sec0 + k;
                if (sec >= ns) sec -= ns;
                keys[o + k] = (unsigned)(base + sec);
                vals[o + k] = (unsigned)t;
            }
            o += (unsigned)nsec;
        }
        return;
    }
    Disc d;
    d.px = r.ux; d.py = r.uy; d.pz = r.uz;   // make_disc only uses the direction
    d.Dx = 1.0;
    d.proj_h = r.ph;
    d.hinv = 1.0 / r.ph;
    make_disc(g, d);   // (from the unit vector: its ring range may differ from the record's by a ring that barely touches)
    const int b0 = (r.rmin - 1) / HPG_BR, b1 = (r.rmax - 1) / HPG_BR;
    unsigned count = 0;
    unsigned o = WRITE ? off[t] : 0u;
    long long ntot = 0;        // Σ run lengths = length of the reference's pixel list (when the centre pixel is in it)
    bool found_c = false;
    for (int bb = b0; bb <= b1; bb += 2) {
        const int band = bb + half;
        const long long ring = (long long)band * HPG_BR + 1 + l16;
        const bool band_ok = band <= b1;
        const int ns = band_ok ? T.band_ns[band] : 1;
        int kref = (int)floor(d.phi * (double)ns / kTwoPi);
        kref = kref < 0 ? 0 : (kref >= ns ? ns - 1 : kref);
        int lo_rel = 1 << 30, hi_rel = -(1 << 30);
        if (band_ok && ring >= r.rmin && ring <= r.rmax) {
            long long sp, nr, j0, cnt;
            bool sh;
            hp_ring_info(g, ring, sp, nr, sh);
            ring_run(g, d, ring, nr, sh, j0, cnt);
            if (!WRITE && fill_ntot) {
                ntot += cnt;
                const long long jc_ = d.cpix - sp;   // the centre pixel, if it lies in this ring
                if (jc_ >= 0 && jc_ < nr && cnt > 0) {
                    long long rel = jc_ - j0;
                    if (rel < 0) rel += nr;
                    found_c = found_c || rel < cnt;
                }
            }
            if (cnt > 0) {
                // unwrapped sector interval of the run: pixels j0 .. j0+cnt-1 (j may exceed nr), sector = floor(j ns / nr)
                const int k_lo = (int)((j0 * ns) / nr);
                const int k_hi = (int)(((j0 + cnt - 1) * ns) / nr);
                const int len = min(ns, k_hi - k_lo + 1);
                if (len >= ns) {
                    lo_rel = -(ns / 2); hi_rel = lo_rel + ns - 1;
                } else {
                    // start of the run relative to the sector of the disc centre's azimuth, kref (ring independent): the
                    // run is an interval around that azimuth, so its start lies at or before kref (+1 for the pixel grid)
                    int a = (k_lo - kref) % ns;
                    if (a < 0) a += ns;            // 0 .. ns-1
                    if (a > 1) a -= ns;            // -(ns-2) .. 1
                    lo_rel = a; hi_rel = a + len - 1;
                }
            }
        }
        // union over the 16 lanes of the half-warp
#pragma unroll
        for (int s = 8; s > 0; s >>= 1) {
            lo_rel = min(lo_rel, __shfl_xor_sync(0xffffffffu, lo_rel, s));
            hi_rel = max(hi_rel, __shfl_xor_sync(0xffffffffu, hi_rel, s));
        }
        int nsec = (band_ok && hi_rel >= lo_rel) ? min(ns, hi_rel - lo_rel + 1) : 0;
        if (!WRITE && bandinfo && band_ok && l16 == 0 && band - b0 < HPG_NBMAX) {
            int sec0 = (kref + lo_rel) % ns;
            if (sec0 < 0) sec0 += ns;
            bandinfo[t * HPG_NBMAX + (band - b0)] = ((unsigned)sec0 << 16) | (unsigned)nsec;
        }
        // the other half-warp's count, to keep the two bands' outputs in order
        const int n_other = __shfl_xor_sync(0xffffffffu, nsec, 16);
        const int base = T.band_base[band_ok ? band : 0];
        if (WRITE) {
            const unsigned my_off = o + (half ? (unsigned)n_other : 0u);
            for (int k = l16; k < nsec; k += 16) {
                int sec = (kref + lo_rel + k) % ns;
                if (sec < 0) sec += ns;
                keys[my_off + k] = (unsigned)(base + sec);
                vals[my_off + k] = (unsigned)t;
            }
            o += (unsigned)(nsec + n_other);
        } else
            count += (unsigned)(nsec + n_other);
    }
    if (!WRITE && fill_ntot) {
        ntot = warp_sum_ll(ntot);
        found_c = __any_sync(0xffffffffu, found_c);
        // push! + unique! (constributing_pixels.jl:16-19): a centre pixel outside its own disc walk would be one more
        // list entry, which the gather has no place for -> negative count = "hand this particle to the scatter walk"
        if (lane == 0) recs[t].ntot = found_c ? (int)ntot : -1;
    }
    if (!WRITE && lane == 0) npairs[t] = count;
}

// counters of the records of an accepted slice (the reference's n_tot = length of the pixel list, main.jl:36-43)
__global__ void __launch_bounds__(256) k_hpg_count(const HRec* __restrict__ recs, long long n_rec,
                                                   unsigned long long* __restrict__ counters)
{
    typedef cub::BlockReduce<unsigned long long, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    unsigned long long n = 0, px = 0;
    if (t < n_rec) {
        const HRec r = recs[t];
        if (r.rmin <= r.rmax) { n = 1; px = (unsigned long long)r.ntot; }
    }
    const unsigned long long both = BR(tmp).Sum((px << 20) | n);   // n <= 256 per block: 20 bits are plenty
    if (threadIdx.x == 0 && both) {
        const unsigned long long nn = both & 0xfffffull, pp = both >> 20;
        atomicAdd(&counters[CNT_MAPPED], nn); atomicAdd(&counters[CNT_GATHER], nn);
        atomicAdd(&counters[CNT_TOUCHED], pp); atomicAdd(&counters[CNT_FOOTPRINT], pp);
    }
}

__global__ void __launch_bounds__(256) k_hpg_tile_bounds(const unsigned* __restrict__ keys, long long m,
                                                         unsigned* __restrict__ tile_beg, unsigned* __restrict__ tile_end)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= m) return;
    const unsigned k = keys[t];
    if (t == 0 || keys[t - 1] != k) tile_beg[k] = (unsigned)t;
    if (t == m - 1 || keys[t + 1] != k) tile_end[k] = (unsigned)(t + 1);
}

__global__ void __launch_bounds__(256) k_hpg_tile_chunks(const unsigned* __restrict__ tile_beg,
                                                         const unsigned* __restrict__ tile_end, int ntiles,
                                                         unsigned* __restrict__ nchunks)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    nchunks[t] = (tile_end[t] - tile_beg[t] + HPG_CHUNK - 1) / HPG_CHUNK;
}

// ---- pass B
// PASSA = true: calculate_weights (pixel_weights.jl:87-140) in the same tile form — per (tile, record) pair the CTA
// sums wk·A' over the tile's pixels (warp shuffle, then shared-memory atomics across the 8 warps) and adds the pair's
// partial sum to Ssum[record].  No pixel-centre trig per (pixel, particle), perfect load balance for huge discs.
// NT: coefficients of G(c2) = asin(x)/x kept (8: exact to 1e-16 up to 0.2 rad; 5: up to 0.073 rad, the bulk of a survey
// volume — a compile-time constant: choosing it per record inside the loop cost more than it saved).
template <int KID, bool BIG, bool PASSA, int NT>
__global__ void __launch_bounds__(HPG_THREADS, (BIG && !PASSA) ? 2 : HPG_CTAS) k_hp_gather(const HRec* __restrict__ recs,
                                                                     const unsigned* __restrict__ vals,
                                                                     const unsigned* __restrict__ tile_beg,
                                                                     const unsigned* __restrict__ tile_end,
                                                                     const unsigned* __restrict__ chunk_begin,
                                                                     HpGeom g, HpTiles T, unsigned total_chunks,
                                                                     double* __restrict__ amap, double* __restrict__ wmap,
                                                                     unsigned long long* __restrict__ counters,
                                                                     double* __restrict__ Ssum)
{
    __shared__ HRecS s_rec[HPG_BATCH];
    __shared__ double s_part[PASSA ? HPG_BATCH : 1];
    // pass A: the lanes' partial sums of 8 records are parked here and reduced together (8 loads + 7 adds + 2 shuffle
    // steps per lane for 8 records, instead of 5 shuffle steps per record); 33-double rows: at most 2-way bank conflicts
    __shared__ double s_red[PASSA ? 8 : 1][PASSA ? 8 : 1][PASSA ? 33 : 1];
    __shared__ unsigned s_work[4];
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    const int rt = tid >> 4, l16 = tid & 15;   // ring of the tile, lane within the ring
    const double inv_ang = 1.0 / g.ang_pix;

    for (;;) {
        if (tid == 0) {
            const unsigned w = (unsigned)atomicAdd(&counters[CNT_WORK], 1ull);
            unsigned tile = 0xffffffffu, b = 0, e = 0, band = 0;
            if (w < total_chunks) {
                int lo = 0, hi = T.ntiles;  // last tile with chunk_begin[tile] <= w
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (chunk_begin[mid] <= w) lo = mid; else hi = mid;
                }
                tile = (unsigned)lo;
                const unsigned c = w - chunk_begin[lo];
                b = tile_beg[lo] + c * HPG_CHUNK;
                e = min(b + HPG_CHUNK, tile_end[lo]);
                int bl = 0, bh = T.nbands;  // last band with band_base[band] <= tile
                while (bh - bl > 1) {
                    const int mid = (bl + bh) >> 1;
                    if (T.band_base[mid] <= (int)tile) bl = mid; else bh = mid;
                }
                band = (unsigned)bl;
            }
            s_work[0] = tile; s_work[1] = b; s_work[2] = e; s_work[3] = band;
        }
        __syncthreads();
        const unsigned tile = s_work[0], wb = s_work[1], we = s_work[2];
        const int band = (int)s_work[3];
        if (tile == 0xffffffffu) break;
        const int sec = (int)tile - T.band_base[band], ns = T.band_ns[band];
        // this thread's ring and pixels
        const long long ring = (long long)band * HPG_BR + 1 + rt;
        const bool ring_ok = ring < g.nl4;
        long long sp = 0, nr = 4;
        bool sh = true;
        RingTrig tr;
        tr.st = 0.0; tr.ct = 2.0; tr.off = 0.5; tr.inv_den = 0.25;   // ct = 2: nothing is ever inside a disc
        if (ring_ok) {
            hp_ring_info(g, ring, sp, nr, sh);
            tr = hp_ring_trig(g, ring);
        }
        const long long jb = ((long long)sec * nr + ns - 1) / ns, je = ((long long)(sec + 1) * nr + ns - 1) / ns;
        double cx[HPG_PPT], cy[HPG_PPT], acc_w[HPG_PPT], acc_q[HPG_PPT];
#pragma unroll
        for (int m = 0; m < HPG_PPT; ++m) {
            const long long j = jb + l16 + 16 * m;
            const bool valid = ring_ok && j < je;
            double s_, c_;
            sincospi(((double)(j + 1) - tr.off) * tr.inv_den, &s_, &c_);   // phi = (iphi - off) pi / den
            cx[m] = valid ? tr.st * c_ : 4.0;   // far away: the membership test fails without a separate flag
            cy[m] = valid ? tr.st * s_ : 4.0;
            acc_w[m] = 0.0; acc_q[m] = 0.0;
        }
        const double cz = tr.ct;
        // rings of this warp (uniform): 2 wq, 2 wq + 1 of the tile
        const int wr_lo = band * HPG_BR + 1 + 2 * wq, wr_hi = wr_lo + 1;

        for (unsigned b = wb; b < we; b += HPG_BATCH) {
            const int nb = (int)min((unsigned)HPG_BATCH, we - b);
            __syncthreads();  // previous batch fully consumed
            for (int t = tid; t < nb; t += HPG_THREADS) {
                const HRec r = recs[vals[b + t]];
                HRecS s;
                s.ux = r.ux; s.uy = r.uy; s.uz = r.uz;
                const double sh_ = sin(0.5 * r.ph);
                s.c2max = 4.0 * sh_ * sh_;
                s.hinv = 1.0 / r.ph;
                s.php = r.ph * inv_ang;
                s.an = r.an; s.anq = r.anq;
                s.rmin = r.rmin; s.rmax = r.rmax;
                s.big = (r.ph + 2.0 * g.ang_pix < 0.2) ? 0 : 1; s.pad1 = 0;
                s_rec[t] = s;
                if (PASSA) s_part[t] = 0.0;
            }
            __syncthreads();
            int kslot = 0, my_e = 0;   // pass A: records parked in s_red[wq]; lane k remembers the record of slot k
            auto flush_slots = [&]() {
                __syncwarp();
                const int k = lane & 7, c0 = (lane >> 3) * 8;
                double sum = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) sum += s_red[PASSA ? wq : 0][PASSA ? k : 0][PASSA ? c0 + j : 0];
                sum += __shfl_xor_sync(0xffffffffu, sum, 8);
                sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                if (lane < kslot && sum != 0.0) atomicAdd(&s_part[PASSA ? my_e : 0], sum);
                kslot = 0;
                __syncwarp();
            };
            for (int e = 0; e < nb; ++e) {
                const HRecS& r = s_rec[e];
                if (wr_hi < r.rmin || wr_lo > r.rmax) continue;   // warp-uniform: none of this warp's rings in the disc
                const double ez = cz - r.uz;
                const double ez2 = fma(ez, ez, 1e-300);           // c2 > 0 even when a pixel centre sits on the particle
                const double c2max = r.c2max;
                if (!__any_sync(0xffffffffu, ez2 < c2max)) continue;
                const double ux = r.ux, uy = r.uy;
                double c2[HPG_PPT];
                bool in[HPG_PPT];
#pragma unroll
                for (int m = 0; m < HPG_PPT; ++m) {
                    const double ex = cx[m] - ux, ey = cy[m] - uy;
                    c2[m] = fma(ex, ex, fma(ey, ey, ez2));
                    in[m] = c2[m] < c2max;
                }
                const double hinv = r.hinv, php = r.php, an = r.an, anq = r.anq;
                double part = 0.0;
#pragma unroll
                for (int m = 0; m < HPG_PPT; ++m) {
                    if (!__any_sync(0xffffffffu, in[m])) continue;   // this 16-pixel group of both rings is outside
                    double t;                                         // 1 - u,  u = dx / proj_h
                    if (BIG && r.big) {                               // (warp-uniform) dx = 2 asin(chord / 2)
                        const double hc = 0.5 * (c2[m] * hp_rsqrt(c2[m]));
                        t = fma(-2.0 * asin(hc < 1.0 ? hc : 1.0), hinv, 1.0);
                    } else {
                        // chord / proj_h = sqrt(c2)/h: MUFU seed y0, then the third-order correction applied to the
                        // PRODUCT c2*y0/h (one instruction fewer than refining y0 first)
                        double y0;
                        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(c2[m]));
                        const double sq0 = c2[m] * y0;
                        const double e = fma(-sq0, y0, 1.0);
                        const double ce = fma(e, kRsq[0], kRsq[1]) * e;
                        const double sqh0 = sq0 * hinv;
                        const double sqh = fma(ce, sqh0, sqh0);
                        double G;
                        if (NT >= 8) {
                            G = fma(kG[7], c2[m], kG[6]);
                            G = fma(G, c2[m], kG[5]);
                            G = fma(G, c2[m], kG[4]);
                        } else
                            G = kG[4];
                        G = fma(G, c2[m], kG[3]);
                        G = fma(G, c2[m], kG[2]);
                        G = fma(G, c2[m], kG[1]);
                        G = fma(G, c2[m], kG[0]);
                        t = fma(-sqh, G, 1.0);
                    }
                    // contributing_area (pixel_weights.jl:6-8): min(ang, |proj_h - (dx - ang/2)|)/ang = min(1, t php + 1/2)
                    const double ap = fma(t, php, 0.5);
                    const double a1 = ap < 1.0 ? ap : 1.0;
                    const double wk = hp_shape_t<KID>(t) * a1;
                    const double wka = in[m] ? wk : 0.0;
                    if (PASSA)
                        part += wka;
                    else {
                        acc_w[m] = fma(wka, an, acc_w[m]);
                        acc_q[m] = fma(wka, anq, acc_q[m]);
                    }
                }
                if (PASSA) {
                    s_red[PASSA ? wq : 0][PASSA ? kslot : 0][PASSA ? lane : 0] = part;
                    if (lane == kslot) my_e = e;
                    if (++kslot == 8) flush_slots();
                }
            }
            if (PASSA && kslot > 0) {
                // unused slots hold stale numbers: only lanes < kslot add their sum
                flush_slots();
            }
            if (PASSA) {
                __syncthreads();
                for (int t = tid; t < nb; t += HPG_THREADS)
                    if (s_part[t] != 0.0) atomicAdd(&Ssum[vals[b + t]], s_part[t]);
            }
        }
        // one flush per work item; lanes 0..15 of a ring write 16 consecutive pixels per group
        if (!PASSA && ring_ok) {
#pragma unroll
            for (int m = 0; m < HPG_PPT; ++m) {
                const long long j = jb + l16 + 16 * m;
                if (j < je && (acc_w[m] != 0.0 || acc_q[m] != 0.0)) {
                    red_add(wmap + sp + j, acc_w[m]);
                    red_add(amap + sp + j, acc_q[m]);
                }
            }
        }
        __syncthreads();  // s_work reuse
    }
}

long long env_ll(const char* name, long long dflt)
{
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    return atoll(s);
}

template <int KID>
int launch_gather_k(s2g_ctx* ctx, const HRec* recs, const unsigned* vals, const unsigned* tbeg, const unsigned* tend,
                    const unsigned* cbeg, const HpGeom& g, const HpTiles& T, unsigned chunks, double* amap, double* wmap,
                    int big, double* Ssum, int nt)
{
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters + CNT_WORK, 0, sizeof(unsigned long long), ctx->stream));
    const int blocks = std::max((int)std::min<long long>((long long)chunks, (long long)ctx->sm_count * ((big && !Ssum) ? 2 : HPG_CTAS)), 1);   // asin + deposit: 100 registers
#define HPG_LAUNCH(B, A, N) k_hp_gather<KID, B, A, N><<<blocks, HPG_THREADS, 0, ctx->stream>>>(recs, vals, tbeg, tend, cbeg, g, T, \
                                                                                               chunks, amap, wmap, ctx->d_counters, Ssum)
    if (Ssum) { if (big) HPG_LAUNCH(true, true, 8); else if (nt >= 8) HPG_LAUNCH(false, true, 8); else HPG_LAUNCH(false, true, 5); }
    else      { if (big) HPG_LAUNCH(true, false, 8); else if (nt >= 8) HPG_LAUNCH(false, false, 8); else HPG_LAUNCH(false, false, 5); }
#undef HPG_LAUNCH
    S2G_CUDA(cudaGetLastError());
    return S2G_OK;
}

int launch_gather(s2g_ctx* ctx, int kernel, const HRec* recs, const unsigned* vals, const unsigned* tbeg,
                  const unsigned* tend, const unsigned* cbeg, const HpGeom& g, const HpTiles& T, unsigned chunks,
                  double* amap, double* wmap, int big, double* Ssum, int nt)
{
    switch (kernel) {
#define HPG_CASE(K) case K: return launch_gather_k<K>(ctx, recs, vals, tbeg, tend, cbeg, g, T, chunks, amap, wmap, big, Ssum, nt);
        HPG_CASE(S2G_KERNEL_CUBIC)
        HPG_CASE(S2G_KERNEL_QUINTIC)
        HPG_CASE(S2G_KERNEL_WENDLAND_C2)
        HPG_CASE(S2G_KERNEL_WENDLAND_C4)
        HPG_CASE(S2G_KERNEL_WENDLAND_C6)
        HPG_CASE(S2G_KERNEL_WENDLAND_C8)
#undef HPG_CASE
    }
    s2g_set_error("unknown kernel id %d", kernel);
    return S2G_EINVAL;
}

}  // namespace

// host tile table of a resolution, cached per context (device copy in the scratch pool)
static int hp_tiles(s2g_ctx* ctx, long long nside, HpTiles& T)
{
    const long long nrings = 4 * nside - 1;
    const int nbands = (int)((nrings + HPG_BR - 1) / HPG_BR);
    std::vector<int> base(nbands + 1), ns(nbands);
    int acc = 0;
    for (int b = 0; b < nbands; ++b) {
        long long lmax = 0;
        for (long long r = (long long)b * HPG_BR + 1; r <= std::min<long long>(nrings, (long long)(b + 1) * HPG_BR); ++r) {
            const long long L = r < nside ? 4 * r : (r <= 3 * nside ? 4 * nside : 4 * (4 * nside - r));
            lmax = std::max(lmax, L);
        }
        ns[b] = (int)((lmax + HPG_SW - 1) / HPG_SW);
        base[b] = acc;
        acc += ns[b];
    }
    base[nbands] = acc;
    void *d_base, *d_ns;
    S2G_TRY(s2g_scratch(ctx, "hpg_band_base", sizeof(int) * (nbands + 1), &d_base));
    S2G_TRY(s2g_scratch(ctx, "hpg_band_ns", sizeof(int) * nbands, &d_ns));
    // pageable source + stream-ordered copy: synchronise before the vectors go out of scope
    S2G_CUDA(cudaMemcpyAsync(d_base, base.data(), sizeof(int) * (nbands + 1), cudaMemcpyHostToDevice, ctx->stream));
    S2G_CUDA(cudaMemcpyAsync(d_ns, ns.data(), sizeof(int) * nbands, cudaMemcpyHostToDevice, ctx->stream));
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    T.nbands = nbands; T.ntiles = acc; T.band_base = (const int*)d_base; T.band_ns = (const int*)d_ns;
    return S2G_OK;
}

int s2g_hp_classify(s2g_ctx* ctx, const s2g_particles& P, long long nside, int calc_mean, const unsigned char* take,
                    double heavy_radius, double gather_radius, int gather_on, unsigned char* heavy, unsigned char* gath,
                    unsigned char* gath_heavy, unsigned char* skip)
{
    const HpGeom g = make_hp(nside);
    k_hp_classify<<<(int)((P.n + 255) / 256), 256, 0, ctx->stream>>>(P, g, calc_mean, take, heavy_radius, gather_radius,
                                                                     gather_on, heavy, gath, gath_heavy, skip);
    S2G_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return S2G_OK;
}

// The gather pipeline over the particles listed in `list` (device, n_list entries): records by the ring-walk pass A
// (s2g_hp_launch_records), pairs, sort, pass B.  Particles that pass A hands back (fallback branch, non-finite
// normalisation) get skip[p] = 0 there and are deposited by the scatter launch that FOLLOWS this call.
int s2g_hp_gather_pipeline(s2g_ctx* ctx, const s2g_particles& P, long long nside, int kernel, int calc_mean,
                           const unsigned* list, long long n_list, unsigned char* skip, double* amap, double* wmap,
                           int coop_records, int big, int series_nt)
{
    if (n_list <= 0) return S2G_OK;
    const HpGeom g = make_hp(nside);
    HpTiles T;
    S2G_TRY(hp_tiles(ctx, nside, T));
    const long long batch_max = env_ll("S2G_HP_BATCH_PARTICLES", 4LL << 20);
    const long long pair_cap = env_ll("S2G_PAIR_CAP", 512LL << 20);
    cudaStream_t st = ctx->stream;
    long long p0 = 0, batch = std::min(batch_max, n_list);
    // pass A: "gather" (default) = the tile form above, no ring walk at all; "walk" (S2G_HP_PASSA=walk) = the ring walk
    // of s2g_healpix.cu in record mode (one warp, or one CTA for a heavy disc, per particle)
    const char* e_pa = getenv("S2G_HP_PASSA");
    const bool passa_gather = !(e_pa && e_pa[0] == 'w');
    while (p0 < n_list) {
        const long long nb = std::min(batch, n_list - p0);
        void *d_recs, *d_np, *d_off, *d_tmp, *d_S = nullptr;
        S2G_TRY(s2g_scratch(ctx, "hpg_recs", sizeof(HRec) * nb, &d_recs));
        S2G_TRY(s2g_scratch(ctx, "hpg_np", sizeof(unsigned) * (nb + 1), &d_np));
        S2G_TRY(s2g_scratch(ctx, "hpg_off", sizeof(unsigned) * (nb + 1), &d_off));
        int ph = s2g_phase_begin(ctx, PH_NORM);
        if (passa_gather) {
            S2G_TRY(s2g_scratch(ctx, "hpg_S", sizeof(double) * nb, &d_S));
            S2G_CUDA(cudaMemsetAsync(d_S, 0, sizeof(double) * nb, st));
            k_hp_records<<<(int)((nb + 255) / 256), 256, 0, st>>>(P, g, list + p0, nb, (HRec*)d_recs);
            S2G_CUDA(cudaGetLastError());
            ctx->launches += 1;
        } else
            S2G_TRY(s2g_hp_launch_records(ctx, P, nside, kernel, calc_mean, list + p0, nb, (HRec*)d_recs, skip,
                                          coop_records));
        s2g_phase_end(ctx, ph);
        ph = s2g_phase_begin(ctx, PH_SORT);
        S2G_CUDA(cudaMemsetAsync((unsigned*)d_np + nb, 0, sizeof(unsigned), st));
        const int wblocks = (int)((nb * 32 + 255) / 256);
        void* d_bi = nullptr;
        S2G_TRY(s2g_scratch(ctx, "hpg_bandinfo", sizeof(unsigned) * (size_t)nb * HPG_NBMAX, &d_bi));
        k_hp_pairs<false><<<wblocks, 256, 0, st>>>((HRec*)d_recs, nb, g, T, nullptr, (unsigned*)d_np, nullptr, nullptr,
                                                   passa_gather ? 1 : 0, (unsigned*)d_bi);
        S2G_CUDA(cudaGetLastError());
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, (const unsigned*)d_np, (unsigned*)d_off, (int)(nb + 1), st);
        S2G_TRY(s2g_scratch(ctx, "g_tmp", tb + 16, &d_tmp));
        S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, (const unsigned*)d_np, (unsigned*)d_off, (int)(nb + 1), st));
        unsigned h_m = 0;
        S2G_CUDA(cudaMemcpyAsync(&h_m, (unsigned*)d_off + nb, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        S2G_CUDA(cudaStreamSynchronize(st));
        ctx->launches += 3;
        const long long m = h_m;
        if (m > pair_cap && nb > 1024) {
            // too many pairs for one slice: shrink it and redo (pass A of this slice is repeated; the counters are only
            // added for the slice that is kept, by k_hpg_count below)
            s2g_phase_end(ctx, ph);
            batch = std::max<long long>(1024, nb / 2);
            continue;
        }
        if (!passa_gather) {
            k_hpg_count<<<(int)((nb + 255) / 256), 256, 0, st>>>((const HRec*)d_recs, nb, ctx->d_counters);
            S2G_CUDA(cudaGetLastError());
        }
        if (m > 0) {
            void *d_keys, *d_vals, *d_keys2, *d_vals2, *d_tend, *d_tbeg, *d_nch, *d_cbeg;
            const int nt = T.ntiles;
            S2G_TRY(s2g_scratch(ctx, "g_keys", sizeof(unsigned) * m, &d_keys));
            S2G_TRY(s2g_scratch(ctx, "g_vals", sizeof(unsigned) * m, &d_vals));
            S2G_TRY(s2g_scratch(ctx, "g_keys2", sizeof(unsigned) * m, &d_keys2));
            S2G_TRY(s2g_scratch(ctx, "g_vals2", sizeof(unsigned) * m, &d_vals2));
            S2G_TRY(s2g_scratch(ctx, "hpg_tend", sizeof(unsigned) * (nt + 1), &d_tend));
            S2G_TRY(s2g_scratch(ctx, "hpg_tbeg", sizeof(unsigned) * (nt + 1), &d_tbeg));
            S2G_TRY(s2g_scratch(ctx, "hpg_nch", sizeof(unsigned) * (nt + 1), &d_nch));
            S2G_TRY(s2g_scratch(ctx, "hpg_cbeg", sizeof(unsigned) * (nt + 1), &d_cbeg));
            k_hp_pairs<true><<<wblocks, 256, 0, st>>>((HRec*)d_recs, nb, g, T, (const unsigned*)d_off, nullptr,
                                                      (unsigned*)d_keys, (unsigned*)d_vals, 0, (unsigned*)d_bi);
            S2G_CUDA(cudaGetLastError());
            int bits = 1;
            while ((1 << bits) < nt) ++bits;
            size_t sb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, sb, (const unsigned*)d_keys, (unsigned*)d_keys2,
                                            (const unsigned*)d_vals, (unsigned*)d_vals2, (int)m, 0, bits, st);
            S2G_TRY(s2g_scratch(ctx, "g_sort_tmp", sb + 16, &d_tmp));
            S2G_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, sb, (const unsigned*)d_keys, (unsigned*)d_keys2,
                                                     (const unsigned*)d_vals, (unsigned*)d_vals2, (int)m, 0, bits, st));
            S2G_CUDA(cudaMemsetAsync(d_tend, 0, sizeof(unsigned) * (nt + 1), st));
            S2G_CUDA(cudaMemsetAsync(d_tbeg, 0, sizeof(unsigned) * (nt + 1), st));
            k_hpg_tile_bounds<<<(int)((m + 255) / 256), 256, 0, st>>>((const unsigned*)d_keys2, m, (unsigned*)d_tbeg,
                                                                      (unsigned*)d_tend);
            S2G_CUDA(cudaGetLastError());
            S2G_CUDA(cudaMemsetAsync(d_nch, 0, sizeof(unsigned) * (nt + 1), st));
            k_hpg_tile_chunks<<<(nt + 255) / 256, 256, 0, st>>>((const unsigned*)d_tbeg, (const unsigned*)d_tend, nt,
                                                                (unsigned*)d_nch);
            S2G_CUDA(cudaGetLastError());
            size_t tb3 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb3, (const unsigned*)d_nch, (unsigned*)d_cbeg, nt + 1, st);
            S2G_TRY(s2g_scratch(ctx, "g_tmp", tb3 + 16, &d_tmp));
            size_t tbb = tb3 + 16;
            S2G_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tbb, (const unsigned*)d_nch, (unsigned*)d_cbeg, nt + 1, st));
            unsigned h_chunks = 0;
            S2G_CUDA(cudaMemcpyAsync(&h_chunks, (unsigned*)d_cbeg + nt, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            s2g_phase_end(ctx, ph);
            ph = -1;
            ctx->launches += 7;
            S2G_CUDA(cudaStreamSynchronize(st));
            if (passa_gather) {
                const int pha = s2g_phase_begin(ctx, PH_NORM);
                int rca = launch_gather(ctx, kernel, (const HRec*)d_recs, (const unsigned*)d_vals2, (const unsigned*)d_tbeg,
                                        (const unsigned*)d_tend, (const unsigned*)d_cbeg, g, T, h_chunks, amap, wmap, big,
                                        (double*)d_S, series_nt);
                if (rca == S2G_OK) {
                    k_hp_normalise<<<(int)((nb + 255) / 256), 256, 0, st>>>((HRec*)d_recs, (const double*)d_S, nb, skip);
                    k_hpg_count<<<(int)((nb + 255) / 256), 256, 0, st>>>((const HRec*)d_recs, nb, ctx->d_counters);
                    if (cudaGetLastError() != cudaSuccess) rca = S2G_ECUDA;
                }
                s2g_phase_end(ctx, pha);
                S2G_TRY(rca);
                ctx->launches += 3;
            }
            const int phg = s2g_phase_begin(ctx, PH_DEPOSIT);
            const int rc = launch_gather(ctx, kernel, (const HRec*)d_recs, (const unsigned*)d_vals2, (const unsigned*)d_tbeg,
                                         (const unsigned*)d_tend, (const unsigned*)d_cbeg, g, T, h_chunks, amap, wmap, big,
                                         nullptr, series_nt);
            s2g_phase_end(ctx, phg);
            S2G_TRY(rc);
            ctx->launches += 1;
            ctx->host_pairs += m;
        }
        else if (passa_gather) {   // no pair at all: every record of the slice goes back to the scatter walk
            k_hp_normalise<<<(int)((nb + 255) / 256), 256, 0, st>>>((HRec*)d_recs, (const double*)d_S, nb, skip);
            S2G_CUDA(cudaGetLastError());
        }
        if (ph >= 0) s2g_phase_end(ctx, ph);
        p0 += nb;
    }
    return S2G_OK;
}
