// spread.cu -- dimension dispatch for the spread / interp stages and the plan-time
// choice of the shared-memory tile geometry for the SM spread engine.
#include <algorithm>
#include <cstdlib>
#include "spreadinterp.cuh"
#include "spread_sm2.cuh"
#include "spread_plane.cuh"

#ifndef CFB_TILE_PAD_PCT
#define CFB_TILE_PAD_PCT 130
#endif

namespace cfb {

template <typename T>
int stage_spread(Plan<T> &p, const typename Plan<T>::C *c, typename Plan<T>::C *fw, int nt)
{
    if (p.M <= 0 || nt <= 0) return 0;
    switch (p.dim) {
        case 1: return launch_spread<T, 1>(p, c, fw, nt);
        case 2: return launch_spread<T, 2>(p, c, fw, nt);
        default: return launch_spread<T, 3>(p, c, fw, nt);
    }
}

template <typename T>
int stage_interp(Plan<T> &p, typename Plan<T>::C *c, const typename Plan<T>::C *fw, int nt)
{
    if (p.M <= 0 || nt <= 0) return 0;
    switch (p.dim) {
        case 1: return launch_interp<T, 1>(p, c, fw, nt);
        case 2: return launch_interp<T, 2>(p, c, fw, nt);
        default: return launch_interp<T, 3>(p, c, fw, nt);
    }
}

// Shared-memory wavefronts one point costs for a candidate (sy, rows-per-plane) tile
// layout: lanes (r, ix) of each pass hit cell r_off + ix; an access of 8-byte cells is
// served per half-warp over 16 8-byte bank pairs, 16-byte cells per quarter-warp over 8.
static int layout_cost(int dim, int ns, int sy, int sz, int cell_bytes)
{
    const int group = cell_bytes == 8 ? 16 : 8, nbank = group;
    if (cell_bytes == 8 && dim >= 2 && sm2_applies(dim, ns) && !sm2_map(dim, ns).single) {
        // second-generation fp32 engine (spread_sm2.cuh): lane = (row slot, column pair); a run flush reads
        // and writes, per pass, the cell at column xp and the cell at column xp + XP of the lane's row
        const Sm2Map m = sm2_map(dim, ns);
        int cost = 0;
        for (int it = 0; it < m.npass; ++it)
            for (int half = 0; half < 2; ++half)
                for (int g0 = 0; g0 < 32; g0 += group) {
                    int mult[16] = {0}, worst = 0;
                    for (int lane = g0; lane < g0 + group; ++lane) {
                        const int r = lane / m.xp, xp = lane - r * m.xp, row = it * m.r + r, col = xp + half * m.xp;
                        if (lane >= m.r * m.xp || row >= m.rows || col >= ns) continue;
                        const int iz = dim == 3 ? row / ns : 0, iy = dim == 3 ? row - iz * ns : row;
                        worst = std::max(worst, ++mult[(iz * sz + iy * sy + col) % nbank]);
                    }
                    cost += worst;
                }
        return cost;
    }
    const int R = 32 / ns, rows = dim == 1 ? 1 : (dim == 2 ? ns : ns * ns), iters = (rows + R - 1) / R;
    int cost = 0;
    for (int it = 0; it < iters; ++it) {
        for (int g0 = 0; g0 < 32; g0 += group) {
            int mult[16] = {0}, worst = 0;
            for (int lane = g0; lane < g0 + group; ++lane) {
                int r = lane / ns, ix = lane - r * ns, row = it * R + r;
                if (lane >= R * ns || row >= rows) continue;
                int iz = dim == 3 ? row / ns : 0, iy = dim == 3 ? row - iz * ns : row;
                int cell = iz * sz + iy * sy + ix;
                worst = std::max(worst, ++mult[cell % nbank]);
            }
            cost += worst;
        }
    }
    return cost;
}

// Tile = bin + ceil(ns/2) halo on every side (reference: src/2d/spreadinterp2d.cu:171,
// shared size check src/2d/spread2d_wrapper.cu:647-654), strides padded so that the
// lane-per-cell passes are bank-conflict free.  Also fixes how many private tiles
// (= warps) one block carries; 0 means the tile does not fit and the GM engine is used.
template <typename T>
void plan_tile_geometry(Plan<T> &p)
{
    const int cell = (int)sizeof(typename Plan<T>::C);
    p.tile_pad = (p.ns + 1) / 2;
    const int ex = p.ibs[0] + 2 * p.tile_pad;
    const int ey = p.dim > 1 ? p.ibs[1] + 2 * p.tile_pad : 1;
    const int ez = p.dim > 2 ? p.ibs[2] + 2 * p.tile_pad : 1;
    long long best_score = -1;
    int best_sy = ex, best_rows = ey, best_cost = 0;
    const long long base_cells = (long long)ex * ey * ez;
    for (int sy = ex; sy < ex + 16; ++sy) {
        for (int rows = ey; rows < ey + (p.dim > 2 ? 16 : 1); ++rows) {
            long long cells = p.dim == 1 ? sy : (long long)sy * rows * ez;
            if (cells * 100 > base_cells * CFB_TILE_PAD_PCT && !(sy == ex && rows == ey)) continue;   // <= 30 % padding
            int cost = layout_cost(p.dim, p.ns, sy, sy * rows, cell);
            long long score = (long long)cost * 1000000 + cells;
            if (best_score < 0 || score < best_score) { best_score = score; best_sy = sy; best_rows = rows; best_cost = cost; }
        }
        if (p.dim == 1) break;
    }
    if (p.dim == 1) { best_sy = ex; best_rows = 1; }
    p.tile_cost = best_cost;
    p.tile_sy = p.dim == 1 ? ((ex + 1) & ~1) : best_sy;
    p.tile_sz = p.tile_sy * best_rows;
    // fp32, 3-D, ns = 6 (config 3; spread_sm2_kernel<3, 6>): the order in which the 36 stencil rows are dealt to the
    // (pass, row slot) pairs is a compile-time table, and for two stride classes (mod 16) a table is built in that makes
    // the run flush free of bank conflicts (spread_sm2.cuh: sm2_rowmap_tab).  The plane stride need not be a multiple
    // of the row stride: the natural 14 x 14 x 8 tile of config 3 gets sz = 197 instead of 196 -- half a percent of
    // padding instead of the 36 % a conflict-free stride costs with the rows in natural order.
    p.sm2_rmc = 0;
    if (cell == 8 && p.dim == 3 && p.ns == 6) {
        static const int classes[][3] = {{14, 5, 1}, {6, 3, 2}};                  // sy mod 16, sz mod 16, table
        long long best_cells = -1;
        for (int sy = ex; sy <= ex + 3; ++sy)
            for (int sz = sy * ey; sz < sy * ey + 16; ++sz)
                for (const auto &m : classes) {
                    if (m[0] != sy % 16 || m[1] != sz % 16) continue;
                    const long long c = (long long)sz * ez;
                    if (c * 10 > base_cells * 11) continue;                        // <= 10 % padding
                    if (best_cells < 0 || c < best_cells) {
                        best_cells = c;
                        p.tile_sy = sy; p.tile_sz = sz; p.tile_cost = 16; p.sm2_rmc = m[2];
                    }
                }
    }
    long long cells = p.dim == 1 ? p.tile_sy : (p.dim == 2 ? (long long)p.tile_sy * ey : (long long)p.tile_sz * ez);
    cells = (cells + 1) & ~1LL;
    p.tile_cells = (int)cells;
    size_t per_warp;
    switch (p.dim) {
        case 1: per_warp = sm_spread_smem_per_warp<T, 1>(p.ns, p.tile_cells); break;
        case 2: per_warp = sm_spread_smem_per_warp<T, 2>(p.ns, p.tile_cells); break;
        default: per_warp = sm_spread_smem_per_warp<T, 3>(p.ns, p.tile_cells); break;
    }
    size_t avail = (size_t)p.max_smem_optin - 18 * 16 * sizeof(T) - 1024;
    long long w = cells > (1 << 24) ? 0 : (long long)(avail / per_warp);
    const int maxw = sm_spread_max_warps(p.dim, p.ns, (int)sizeof(T));
    if (w > 16) w = 8;                  // several blocks per SM instead of one huge block
    if (w > maxw) w = maxw;
    p.sm_warps = (int)w;
}

// The plane-owner engine (spread_plane.cuh: double precision, 3-D, ns >= 9, gpu_method 2): the block's tile is the
// whole reference bin with its halo, the row stride = ns (mod 8) cells (conflict-free 16-byte cells per quarter-warp).
// Returns false when it does not apply or the tile + batch scratch do not fit; the warp-private engine then serves.
template <typename T>
static bool plane_geometry(Plan<T> &p)
{
    p.plane_engine = false;
    if constexpr (sizeof(T) == 8) {
        static const bool off = [] { const char *e = getenv("CFB_PLANE"); return e && e[0] == '0'; }();   // A/B measurements
        if (off || !(p.type == 1 && p.method == 2 && p.sorted && p.dim == 3 && plane_engine_ns(p.ns))) return false;
        const int pad = (p.ns + 1) / 2;
        const int ex = p.bs[0] + 2 * pad, ey = p.bs[1] + 2 * pad, ez = p.bs[2] + 2 * pad;
        const size_t slot = (size_t)((((4 * p.ns) / 2) | 1) * 2);
        auto bytes = [&](long long cells) {
            return 18 * 16 * sizeof(T) + (size_t)cells * 2 * sizeof(T) + CFB_PLANE_PB * slot * sizeof(double) + 2 * CFB_PLANE_PB * sizeof(int) + 1024;
        };
        int sy = ex;
        while (sy % 8 != p.ns % 8) ++sy;
        if (bytes((long long)sy * ey * ez) > (size_t)p.max_smem_optin) sy = ex;      // no room for the conflict-free stride
        const long long cells = (long long)sy * ey * ez;
        if (cells > (1 << 24) || bytes(cells) > (size_t)p.max_smem_optin) return false;
        for (int d = 0; d < 3; ++d) { p.ibs[d] = p.bs[d]; p.spb[d] = 1; }
        p.nibins = p.nbins;
        p.tile_pad = pad; p.tile_sy = sy; p.tile_sz = sy * ey; p.tile_cells = (int)cells; p.tile_cost = 0;
        p.sm_warps = ez < 16 ? (ez < CFB_PLANE_PB / 32 ? CFB_PLANE_PB / 32 : ez) : 16;     // phase A takes CFB_PLANE_PB threads
        p.plane_engine = true;
        return true;
    }
    return false;
}

// Internal bins for the SM spread engine.  The reference's bins (16x16x2 in 3-D, 32x32 in 2-D) give
// every warp a private tile of 20-40 KB, i.e. 5-12 resident warps per SM, and the kernel is then
// bound by the dependent-issue rate of those few warps (profiles/r01k: 30 % issue slots used, stall
// reason "wait").  Splitting each reference bin into sub-bins shrinks the tile and multiplies the
// resident warps; the halo (flush) overhead per point grows, so the split stops at ~14 KB per warp
// (16 warps per SM) and needs >= 64 points per sub-bin on average.  Reference-facing arrays are
// unaffected: the sort stays bin-major and their values are sums over the sub-bins (setpts.cu).
template <typename T>
void choose_internal_bins(Plan<T> &p, long long M)
{
    for (int d = 0; d < 3; ++d) { p.ibs[d] = p.bs[d]; p.spb[d] = 1; }
    p.nibins = p.nbins;
    p.imaxsub = p.opts.gpu_maxsubprobsize;
    p.ilist = false;
    p.plane_engine = false;
    plan_tile_geometry(p);
    if (!(p.type == 1 && p.method == 2 && p.sorted) || p.dim == 1) return;
    if (plane_geometry(p)) {
        // one tile per block: no sub-bins; larger work items still save flushes on dense inputs
        const long long per_item = M / (16LL * p.num_sms);
        const long long lo = p.opts.gpu_maxsubprobsize, hi = 4096;
        p.imaxsub = (int)(per_item < lo ? lo : (per_item > hi ? (hi > lo ? hi : lo) : per_item));
        p.ilist = p.imaxsub != p.opts.gpu_maxsubprobsize;
        return;
    }
    // Work-item size of the engines' own list: every item ends with a flush of its tile to the fine
    // grid (15 % of the config-3 kernel's instructions with 1024-point items), so dense inputs get
    // items of up to 4096 points -- as long as >= 16 items per resident warp remain for balance.
    {
        const long long per_item = M / (16LL * p.num_sms * 16);
        static const long long hi_env = [] { const char *e = getenv("CFB_SM_MAXITEM"); return e ? atoll(e) : 0LL; }();   // experiments
        const long long lo = p.opts.gpu_maxsubprobsize, hi = hi_env > 0 ? hi_env : 4096;
        p.imaxsub = (int)(per_item < lo ? lo : (per_item > hi ? (hi > lo ? hi : lo) : per_item));
    }
    static const size_t target_env = [] { const char *e = getenv("CFB_SM_TARGET_KB"); return e ? (size_t)atoll(e) * 1024 : (size_t)0; }();   // experiments
    const size_t target = target_env ? target_env : 14 * 1024;
    for (;;) {
        size_t per_warp;
        switch (p.dim) {
            case 2: per_warp = sm_spread_smem_per_warp<T, 2>(p.ns, p.tile_cells); break;
            default: per_warp = sm_spread_smem_per_warp<T, 3>(p.ns, p.tile_cells); break;
        }
        if (p.sm_warps > 0 && per_warp <= target) break;
        // halve one dimension (even, >= 8 cells): the one whose halved tile has the cheapest
        // bank-conflict-free layout, then the fewest cells (in 3-D that keeps x = 16 + halo = 22
        // cells, whose natural stride is conflict free for ns = 6, and halves y)
        // (the density rule yields when the tile leaves fewer than four warps per SM: wide fp64
        // stencils in 3-D, where one 130 KB tile per SM would serialise everything)
        const long long nib = (long long)p.nibins * 2;
        if ((M < 64 * nib && p.sm_warps >= 4) || nib > (1LL << 28)) break;
        int best = -1;
        long long best_score = 0;
        for (int d = 0; d < p.dim; ++d) {
            if (p.ibs[d] % 2 != 0 || p.ibs[d] < 8) continue;
            p.ibs[d] /= 2;
            plan_tile_geometry(p);
            const long long score = (long long)p.tile_cost * (1LL << 32) + p.tile_cells;
            p.ibs[d] *= 2;
            if (best < 0 || score < best_score) { best = d; best_score = score; }
        }
        if (best < 0) { plan_tile_geometry(p); break; }
        p.ibs[best] /= 2; p.spb[best] *= 2; p.nibins = (int)nib;
        plan_tile_geometry(p);
    }
    p.ilist = p.nibins != p.nbins || p.imaxsub != p.opts.gpu_maxsubprobsize;
}

// Whether the tile interpolation engine serves this plan's points (the launcher's rule, spreadinterp_launch.cuh:
// do_interp_tile, without the occupancy query): setpts needs to know, because it orders the points of a bin by
// shared-memory bank class for that engine and by stencil cell for the gather engine.
template <typename T>
bool interp_tile_applies(const Plan<T> &p)
{
    using C = typename Plan<T>::C;
    if (!p.sorted || p.interp_engine == 1 || p.M <= 0) return false;
    const int pad = (p.ns + 1) / 2;
    const size_t cells = (size_t)(p.ibs[0] + 2 * pad) * (p.dim > 1 ? p.ibs[1] + 2 * pad : 1) * (p.dim > 2 ? p.ibs[2] + 2 * pad : 1);
    const size_t smem = 18 * 16 * sizeof(T) + cells * sizeof(C);
    if (smem + 1024 > (size_t)p.max_smem_optin) return false;
    const bool sparse = (unsigned long long)p.M * 64ull < (unsigned long long)p.nbins * cells;
    const bool planes_fit_l2 = p.dim < 3 || (long long)p.ns * p.nf1 * p.nf2 * (long long)sizeof(C) <= p.l2_bytes / 2;
    return !(p.interp_engine == 0 && sparse && planes_fit_l2);
}

template int stage_spread<float>(Plan<float> &, const float2 *, float2 *, int);
template int stage_spread<double>(Plan<double> &, const double2 *, double2 *, int);
template int stage_interp<float>(Plan<float> &, float2 *, const float2 *, int);
template int stage_interp<double>(Plan<double> &, double2 *, const double2 *, int);
template bool interp_tile_applies<float>(const Plan<float> &);
template bool interp_tile_applies<double>(const Plan<double> &);
template void plan_tile_geometry<float>(Plan<float> &);
template void plan_tile_geometry<double>(Plan<double> &);
template void choose_internal_bins<float>(Plan<float> &, long long);
template void choose_internal_bins<double>(Plan<double> &, long long);

}  // namespace cfb
