// spread_plane.cuh -- double-precision 3-D spreading with wide stencils (ns >= 9: tol <= 1e-8), gpu_method 2.
// Replaces Spread_3d_Subprob[_Horner] (src/3d/spreadinterp3d.cu:180-389) for the plans whose stencil is too large
// for register accumulators (spreadinterp.cuh: Geo::MERGE == false; config 5's type-1 twin: 1000 cells per point).
//
// spread_sm_kernel gives every warp a PRIVATE tile: with ns = 10 that is 62 KB even for an 8 x 8 x 2 sub-bin, three
// warps per SM, and a warp alone cannot hide the load -> FMA -> store chain of 34 passes per point (measured 3.8 ns per
// point and GPU, 37 % of the shared-memory pipe; smaller sub-bins make it slower: the tile flushes grow faster than the
// occupancy, profiles/r03g).  Here the BLOCK owns one tile of a whole reference bin with its halo (26 x 26 x 12 cells
// = 130 KB) and the warps own its z PLANES: every warp walks every point of the batch and updates only the 10 x 10
// cells of its own plane(s) -- no two warps ever touch the same cell, so there are no atomics and no barriers inside
// a batch, and 12 warps share the latency.  Per point and plane: lane = (row group g, column ix), rows g, g + RG, ...:
// one 8-byte load of kx[ix], one of ky per row, one 16-byte load of (kz c_re, kz c_im) for the plane, and per row
// LDS.128 / 2 DFMA / STS.128 on the tile.  The row stride of the tile is chosen = ns (mod 8) cells, which makes the
// 16-byte cells of a quarter-warp fall on 8 different bank groups (plan_tile_geometry).
//   phase A (thread per point, 256 points per batch): record, strength, 3 x ns kernel values -> the batch scratch
//   phase B (warp per plane): as above
//   flush (whole block): tile -> fine grid with RED over the non-zero cells, tile cleared on the way
#pragma once
#include "spreadinterp.cuh"

#ifndef CFB_PLANE_PB
#define CFB_PLANE_PB 256
#endif

namespace cfb {

template <int NS> struct GeoP {
    static constexpr int RG = 32 / NS;                       // row groups per pass
    static constexpr int LANES = RG * NS;
    static constexpr int NPASS = (NS + RG - 1) / RG;
    static constexpr int PB = CFB_PLANE_PB;                  // points per batch
    static constexpr int SLOT = (((4 * NS) / 2) | 1) * 2;    // doubles per point: kx[NS] | ky[NS] | (kz c_re, kz c_im)[NS], odd in 16-byte units
    static constexpr size_t SCRATCH = (size_t)PB * SLOT * sizeof(double) + 2 * PB * sizeof(int);
    static constexpr int MAXW = 16;                          // warps per block (one per tile plane when ez <= 16)
};

constexpr bool plane_engine_ns(int ns) { return ns >= 9; }     // = the widths with Geo<double, 3, ns>::MERGE == false

template <int NS, bool HORNER>
__global__ void __launch_bounds__(32 * GeoP<NS>::MAXW)
spread_plane_kernel(const SIArgs<double> a_in)
{
    using G = GeoP<NS>;
    using C = double2;
    SIArgs<double> a = a_in;
    a.horner = HORNER ? 1 : 0;
    extern __shared__ __align__(16) unsigned char smem[];
    double *s_hc = reinterpret_cast<double *>(smem);
    C *tile = reinterpret_cast<C *>(smem + 18 * 16 * sizeof(double));
    double *slots = reinterpret_cast<double *>(tile + a.tile_cells);
    int *s_xy = reinterpret_cast<int *>(slots + G::PB * G::SLOT), *s_z = s_xy + G::PB;
    __shared__ long long s_work;
    stage_horner<double, NS>(a, s_hc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const bool active = lane < G::LANES;
    const int g = active ? lane / NS : 0, ix = active ? lane - g * NS : 0;
    const int nsub = *a.nsub;
    const long long total = (long long)nsub * a.nt;
    const int ex = a.ex, ey = a.ey, ez = a.ez;

    for (int i = threadIdx.x; i < a.tile_cells; i += blockDim.x) tile[i] = C{0.0, 0.0};   // clean at every item start: here once, then by every flush

    for (;;) {
        __syncthreads();                                  // the previous flush is complete (and the initial clear)
        if (threadIdx.x == 0) s_work = atomicAdd(a.counter, 1);
        __syncthreads();
        const long long w = s_work;
        if (w >= total) break;
        const int t = (int)(w / nsub), s = (int)(w - (long long)t * nsub);
        int pstart, n, ox, oy, oz;
        decode_subproblem<double, 3>(a, s, pstart, n, ox, oy, oz);
        const C *cin = a.c + (size_t)t * a.M;
        C *fwt = a.fw + (size_t)t * a.fwstride;

        for (int base = 0; base < n; base += G::PB) {
            const int cnt = min(G::PB, n - base);
            // ---- phase A: one point per thread
            if ((int)threadIdx.x < cnt) {
                const PtRec<double> rec = load_rec(a.recs + pstart + base + threadIdx.x);
                const C cv = cin[rec_index(rec)];
                double kx[NS], ky[NS], kz[NS];
                const int xs = stencil_start(rec.x, NS), ys = stencil_start(rec.y, NS), zs = stencil_start(rec.z, NS);
                kernel_vector<double, NS, true>(kx, (double)xs - rec.x, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
                kernel_vector<double, NS, true>(ky, (double)ys - rec.y, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
                kernel_vector<double, NS, true>(kz, (double)zs - rec.z, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
                double2 *slot = reinterpret_cast<double2 *>(slots + (size_t)threadIdx.x * G::SLOT);
#pragma unroll
                for (int i = 0; i < NS / 2; ++i) slot[i] = make_double2(kx[2 * i], kx[2 * i + 1]);
                if (NS & 1) slots[(size_t)threadIdx.x * G::SLOT + NS - 1] = kx[NS - 1];
#pragma unroll
                for (int i = 0; i < NS; ++i) slots[(size_t)threadIdx.x * G::SLOT + NS + i] = ky[i];
#pragma unroll
                for (int i = 0; i < NS; ++i) slot[NS + i] = make_double2(kz[i] * cv.x, kz[i] * cv.y);
                s_xy[threadIdx.x] = clampi(ys - oy, 0, ey - NS) * a.sy + clampi(xs - ox, 0, ex - NS);
                s_z[threadIdx.x] = clampi(zs - a.zshift - oz, 0, ez - NS);      // weights from the global coordinate, grid index slab-local
            }
            __syncthreads();
            // ---- phase B: every warp walks the batch and updates its own planes
            for (int q = 0; q < cnt; ++q) {
                const int zo = s_z[q];
                const double *slot = slots + (size_t)q * G::SLOT;
                for (int pl = warp; pl < ez; pl += nwarps) {
                    const int j = pl - zo;
                    if (j < 0 || j >= NS) continue;                              // warp-uniform
                    const double2 kzc = *reinterpret_cast<const double2 *>(slot + 2 * NS + 2 * j);
                    const double kxv = active ? slot[ix] : 0.0;
                    C *cell0 = tile + (size_t)pl * a.sz + s_xy[q] + ix;
                    C v[G::NPASS];
                    double wy[G::NPASS];
#pragma unroll
                    for (int ps = 0; ps < G::NPASS; ++ps) {
                        const int iy = g + G::RG * ps;
                        const bool on = active && iy < NS;
                        wy[ps] = on ? slot[NS + iy] : 0.0;
                        v[ps] = on ? cell0[iy * a.sy] : C{0.0, 0.0};
                    }
#pragma unroll
                    for (int ps = 0; ps < G::NPASS; ++ps) {
                        const double wxy = kxv * wy[ps];
                        v[ps].x = fma(wxy, kzc.x, v[ps].x);
                        v[ps].y = fma(wxy, kzc.y, v[ps].y);
                    }
#pragma unroll
                    for (int ps = 0; ps < G::NPASS; ++ps) {
                        const int iy = g + G::RG * ps;
                        if (active && iy < NS) cell0[iy * a.sy] = v[ps];
                    }
                }
                __syncwarp();      // the next point's loads (other lanes, possibly the same cells) come after these stores
            }
            __syncthreads();       // the batch scratch is free again
        }

        // ---- tile -> fine grid: vector RED over the non-zero cells, clearing the tile on the way; single periodic
        // wrap (the reference's guard ix < nf + pad, src/3d/spreadinterp3d.cu, is implied: cells past it stay zero)
        {
            const int ncell = ex * ey * ez;
            const float inv_ex = 1.0f / (float)ex, inv_ey = 1.0f / (float)ey;
            for (int idx = threadIdx.x; idx < ncell; idx += blockDim.x) {
                const int row = (int)(((float)idx + 0.5f) * inv_ex);
                const int lx = idx - row * ex;
                const int lz = (int)(((float)row + 0.5f) * inv_ey);
                const int ly = row - lz * ey;
                C *tp = tile + (size_t)lz * a.sz + ly * a.sy + lx;
                const C v = *tp;
                if (v.x != 0.0 || v.y != 0.0) {
                    const size_t gi = (size_t)wrap_index(ox + lx, a.nf1) + (size_t)wrap_index(oy + ly, a.nf2) * a.nf1 +
                                      (size_t)wrap_index(oz + lz, a.nf3) * a.nf1 * a.nf2;
                    red_add(fwt + gi, v.x, v.y);
                    *tp = C{0.0, 0.0};
                }
            }
        }
    }
}

}  // namespace cfb
