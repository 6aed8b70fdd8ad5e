// s2g_cic3d.cuh — per-particle cell-space record of the 3D Smac deposit (shared by the scatter and gather kernels).
#pragma once
#include "s2g_cic2d.cuh"

#ifdef __CUDACC__
struct Rec3 {
    double x, y, z, h, hinv, vol, w, q;
    int lo[3], hi[3];
};

__device__ __forceinline__ bool make_rec3(const s2g_particles& P, const s2g_geom& G, long long p, Rec3& r)
{
    r.q = ld_in(P.binq, p, P.in_dtype);
    if (r.q == 0.0 && !G.calc_mean) return false;  // cic_3D.jl:142
    const double px = ld_pos(P, p, 0), py = ld_pos(P, p, 1), pz = ld_pos(P, p, 2);
    if (P.fuse_center && !in_image(P, px, py, pz)) return false;
    const double hs = ld_in(P.hsml, p, P.in_dtype);
    const double mm = ld_in(P.m, p, P.in_dtype);
    const double rh = ld_in(P.rho, p, P.in_dtype);
    r.w = ld_in(P.w, p, P.in_dtype);
    r.h = __dmul_rn(hs, G.len2pix);
    r.hinv = __ddiv_rn(1.0, r.h);
    r.vol = __ddiv_rn(mm, __ddiv_rn(rh, G.l3));
    r.x = __dadd_rn(__dmul_rn(px, G.len2pix), G.half_n);
    r.y = __dadd_rn(__dmul_rn(py, G.len2pix), G.half_n);
    r.z = __dadd_rn(__dmul_rn(pz, G.len2pix), G.half_n);
    const int n1 = (int)G.npix - 1;
    const double c[3] = {r.x, r.y, r.z};
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        r.lo[d] = max(floor_to_int(__dadd_rn(c[d], -r.h)), 0);
        r.hi[d] = min(floor_to_int(__dadd_rn(c[d], r.h)), n1);
        ok = ok && (r.lo[d] <= r.hi[d]);
    }
    return ok;
}

#endif
