// spread_sm2.cuh -- the single-precision shared-memory spreading engine (gpu_method 2), second
// generation.  Replaces Spread_{2,3}d_Subprob[_Horner] (src/2d/spreadinterp2d.cu:153-312,
// src/3d/spreadinterp3d.cu:180-389) for fp32 in 2-D and in 3-D up to ns = 7 (BASELINE configs 1, 3, 4);
// spread_sm_kernel (spreadinterp.cuh) keeps serving fp64, 1-D and the very wide 3-D stencils.
//
// Same decomposition as spread_sm_kernel -- a warp owns a private padded tile of one (internal) bin,
// phase A is thread-per-point, phase B lane-per-cell with lane = (row slot r, column ix), runs of
// points with one stencil origin are accumulated in registers and added to the tile once -- but the
// instruction stream is rebuilt around what ncu showed for the first generation on config 3
// (profiles/r01y: 77 warp instructions per point, 8 of them FFMA2):
//   * lane = (row slot, column PAIR): a lane owns columns xp and xp + ceil(ns/2) of its rows.  Phase A
//     writes the operands in the form phase B needs them: per pair (c_re kx_a, c_re kx_b, c_im kx_a,
//     c_im kx_b) and per row slot the row weights of all passes (ky*kz).  Phase B per point and lane is
//     ONE 16-byte load for the pair, one for (up to four) passes' weights, and per pass two FFMA2 whose
//     scalar operand is the row weight (FFMA2 broadcasts a 32-bit source): no multiplies, no packing.
//     The second generation's first cut (lane per cell, 8 passes) spent 12 of its 19 shared-memory
//     wavefronts per point on these loads and ran at 68 % of the shared-memory pipe (profiles/r02b);
//     pairs of columns need 8.
//   * one loop over the points of a batch; a new run is a warp-uniform branch around the flush
//     (no per-run loop set-up, no separate point-by-point path); the next point's operands are
//     loaded before the current point's FFMA2s (explicit two-stage pipeline: the kernel runs with
//     ~2.5 warps per scheduler, latency has to be hidden inside the warp).
//   * the run flush uses per-lane row offsets kept in registers: LDS.64 / 2 FADD / STS.64 per pass.
//   * tile -> fine grid row by row over the touched box (row arithmetic is warp-uniform, lanes over x,
//     two rows in flight), zero cells skipped, vector RED.
//   * stencils that fit one pass of lane-per-cell (2-D ns <= 5: configs 1 and 4) pair (re, im) instead:
//     one FFMA2 per point and lane.
#pragma once
#include <type_traits>
#include "spreadinterp.cuh"

namespace cfb {

template <int DIM, int NS> struct Geo2 {
    static_assert(DIM == 2 || DIM == 3, "2-D and 3-D only");
    static constexpr int ROWS = DIM == 2 ? NS : NS * NS;               // stencil rows (y, or (y, z) flattened)
    // SINGLE: the whole stencil in one pass with lane = (row, column); the FFMA2 pairs (re, im)
    static constexpr bool SINGLE = ROWS <= 32 / NS;
    // otherwise lane = (row slot, column pair): a lane owns columns xp and xp + XP of its rows; the FFMA2
    // pairs those two columns and takes the row weight as its scalar (broadcast) operand
    static constexpr int XP = (NS + 1) / 2;
    static constexpr int R = SINGLE ? 32 / NS : 32 / XP;               // stencil rows per pass
    static constexpr int LANES = SINGLE ? R * NS : R * XP;
    static constexpr int NPASS = SINGLE ? 1 : (ROWS + R - 1) / R;
    static constexpr int WS = SINGLE ? 2 : (NPASS <= 2 ? NPASS : roundup(NPASS, 4));   // floats per row slot
    static constexpr int CKSEG = SINGLE ? roundup(2 * NS, 4) : 4 * XP;
    static constexpr int WSEG = roundup(R * WS, 4);
    static constexpr int SLOT = (((CKSEG + WSEG) / 4) | 1) * 4;        // odd in 16-byte units: phase A's vector stores are conflict free
    static constexpr int NSLOT = 33;                                   // one spare slot: the pipeline reads one point ahead
    static constexpr size_t SCRATCH = (size_t)NSLOT * SLOT * sizeof(float) + 32 * sizeof(int);
    static constexpr int MAXW = NPASS > 8 ? 8 : 16;                    // warps per block (register budget)
};

// Row orders of the 3-D ns = 6 kernel (config 3): stencil row (iz*6 + iy) of (pass, row slot), -1 = none.  Class 0 is
// the natural order.  For the tile strides (sy, sz) = (14, 5) mod 16 [class 1: the 14 x 14 x 8 tile of 8 x 8 x 2 sub-bins
// with one cell of padding per plane] and (6, 3) mod 16 [class 2: the 22 x 22 x 8 tile of the reference's own bins] the
// orders below make the 16 lanes of every half-warp of a run flush touch 16 different 8-byte bank pairs
// (tools/search_sm2_rowmap.py; all classes it found: sm2_rowmaps.inc).  The table is a compile-time constant so that
// phase A still writes the row weights of a row slot's passes next to each other (one LDS.128 in phase B).
constexpr int SM2_RMC_COUNT = 3;
// stencil row served by row slot rr in pass it (-1: none); RMC > 0 only for DIM == 3, NS == 6
template <int DIM, int NS, int RMC>
__host__ __device__ constexpr int sm2_row(int it, int rr)
{
    constexpr int XP = (NS + 1) / 2, R = 32 / XP, ROWS = DIM == 2 ? NS : NS * NS;
    constexpr signed char tab[SM2_RMC_COUNT][40] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, -1, -1, -1, -1},
    {0, 2, 5, 7, 21, 9, 3, 6, 10, 16, 1, 4, 8, 11, 15, 22, 12, 17, 19, 23, 13, 18, 20, 25, 32, 27, 24, 29, 31, 34, 14, 26, 28, 30, 33, 35, -1, -1, -1, -1},
    {0, 1, 2, 6, 7, 8, 3, 4, 9, 10, 5, 11, 15, 17, 21, 23, 13, 14, 19, 20, 12, 18, 24, 25, 30, 31, 16, 22, 26, 32, 27, 28, 29, 33, 34, 35, -1, -1, -1, -1},
};
    if (RMC > 0) return tab[RMC][it * R + rr];
    return it * R + rr < ROWS ? it * R + rr : -1;
}

constexpr bool sm2_applies(int dim, int ns) { return dim == 2 || (dim == 3 && ns <= 7); }

// host mirror of the lane mapping, for the tile-layout search (spread.cu: layout_cost)
struct Sm2Map { bool single; int xp, r, rows, npass; };
inline Sm2Map sm2_map(int dim, int ns)
{
    Sm2Map m;
    m.rows = dim == 2 ? ns : ns * ns;
    m.single = m.rows <= 32 / ns;
    m.xp = (ns + 1) / 2;
    m.r = m.single ? 32 / ns : 32 / m.xp;
    m.npass = m.single ? 1 : (m.rows + m.r - 1) / m.r;
    return m;
}

// packed FP32 FMA (FFMA2) through the compiler's own builtin: the accumulators stay ordinary float2
// values for the register allocator, and a (w, w) operand becomes the scalar-broadcast form of FFMA2
__device__ __forceinline__ void fma2(float2 &d, float2 a, float2 b) { d = __ffma2_rn(a, b, d); }

template <int DIM, int NS, bool HORNER, int RMC = 0>
__global__ void __launch_bounds__(32 * Geo2<DIM, NS>::MAXW)
spread_sm2_kernel(const SIArgs<float> a_in)
{
    using G = Geo2<DIM, NS>;
    using C = float2;
    SIArgs<float> a = a_in;
    a.horner = HORNER ? 1 : 0;
    extern __shared__ __align__(16) unsigned char smem[];
    float *s_hc = reinterpret_cast<float *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = (size_t)a.tile_cells * sizeof(C) + G::SCRATCH;
    unsigned char *wbase = smem + 18 * 16 * sizeof(float) + warp * per_warp;
    C *tile = reinterpret_cast<C *>(wbase);
    float *slots = reinterpret_cast<float *>(wbase + (size_t)a.tile_cells * sizeof(C));
    int *s_off = reinterpret_cast<int *>(slots + G::NSLOT * G::SLOT);
    stage_horner<float, NS>(a, s_hc);

    constexpr int NCOL = G::SINGLE ? NS : G::XP;                       // lanes per row
    const bool active = lane < G::LANES;
    const int r = active ? lane / NCOL : 0, ix = active ? lane - r * NCOL : 0;
    // stencil row of this lane's row slot per pass (SINGLE: row = r)
    auto row_of = [&](int it) -> int {
        if constexpr (G::SINGLE) return r < G::ROWS ? r : -1;
        else {
            int row = -1;
#pragma unroll
            for (int rr = 0; rr < G::R; ++rr) if (rr == r) row = sm2_row<DIM, NS, RMC>(it, rr);
            return row;
        }
    };
    const bool last_ok = active && row_of(G::NPASS - 1) >= 0;          // the last pass may run past the stencil
    const bool b_ok = G::SINGLE || ix + G::XP < NS;                    // odd widths: the last pair has one column only
    int tb[G::NPASS];                                                  // tile offset of this lane's (first) cell per pass
#pragma unroll
    for (int it = 0; it < G::NPASS; ++it) {
        int row = row_of(it);
        if (row < 0) row = 0;
        if (DIM == 2) tb[it] = row * a.sy + ix;
        else { const int iz = row / NS, iy = row - iz * NS; tb[it] = iz * a.sz + iy * a.sy + ix; }
    }
    // this lane's operands inside a point slot
    // (the idle lanes 30, 31 of a 10 x 3 mapping read what their neighbour reads: with r = 0 they hit the bank group of
    // row slot 8 at another address, a fifth wavefront for every weight load -- profiles/r03n)
    const float *my_ck = slots + (active ? ix : NCOL - 1) * (G::SINGLE ? 2 : 4);
    const float *my_w = slots + G::CKSEG + (active ? r : G::R - 1) * G::WS;

    const int nsub = *a.nsub;
    const long long total = (long long)nsub * a.nt;
    for (int i = lane; i < a.tile_cells; i += 32) tile[i] = C{0.0f, 0.0f};   // clean at every item start: here once, then by every flush

    for (;;) {
        long long w = 0;
        if (lane == 0) w = atomicAdd(a.counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total) break;
        const int t = (int)(w / nsub), s = (int)(w - (long long)t * nsub);
        int pstart, n, ox, oy, oz;
        decode_subproblem<float, DIM>(a, s, pstart, n, ox, oy, oz);
        const C *cin = a.c + (size_t)t * a.M;
        C *fwt = a.fw + (size_t)t * a.fwstride;
        const PtRec<float> *recs = a.recs + pstart;

        // SINGLE: are[0] = (re, im) of the lane's cell.  Else per pass: are = re of (column xp, column xp + XP), aim = im
        float2 are[G::NPASS], aim[G::NPASS];
#pragma unroll
        for (int j = 0; j < G::NPASS; ++j) { are[j] = make_float2(0.0f, 0.0f); aim[j] = make_float2(0.0f, 0.0f); }
        int cur = -1;                                      // tile offset of the open run
        int ylo = 1 << 30, yhi = -1, zlo = 1 << 30, zhi = -1;

        auto flush_run = [&]() {
            __syncwarp();                                  // earlier tile updates of other lanes come first
            C *cell0 = tile + cur;
            if constexpr (G::SINGLE) {
                if (last_ok) {
                    C v = cell0[tb[0]];
                    v.x += are[0].x; v.y += are[0].y;
                    cell0[tb[0]] = v;
                }
                are[0] = make_float2(0.0f, 0.0f);
            } else {
                C va[G::NPASS], vb[G::NPASS];
#pragma unroll
                for (int it = 0; it < G::NPASS; ++it)
                    if (it < G::NPASS - 1 ? active : last_ok) {
                        va[it] = cell0[tb[it]];
                        if (b_ok) vb[it] = cell0[tb[it] + G::XP];
                    }
#pragma unroll
                for (int it = 0; it < G::NPASS; ++it) {
                    va[it].x += are[it].x; va[it].y += aim[it].x;
                    vb[it].x += are[it].y; vb[it].y += aim[it].y;
                }
#pragma unroll
                for (int it = 0; it < G::NPASS; ++it)
                    if (it < G::NPASS - 1 ? active : last_ok) {
                        cell0[tb[it]] = va[it];
                        if (b_ok) cell0[tb[it] + G::XP] = vb[it];
                    }
#pragma unroll
                for (int j = 0; j < G::NPASS; ++j) { are[j] = make_float2(0.0f, 0.0f); aim[j] = make_float2(0.0f, 0.0f); }
            }
        };

        // software pipeline over the batches: records two batches ahead, strengths one batch ahead
        PtRec<float> rec_cur = lane < n ? load_rec(recs + lane) : null_rec<float>();
        PtRec<float> rec_nxt = 32 + lane < n ? load_rec(recs + 32 + lane) : null_rec<float>();
        C c_cur = lane < n ? __ldg(cin + rec_index(rec_cur)) : C{0, 0};

        for (int base = 0; base < n; base += 32) {
            const int cnt = min(32, n - base);
            const C c_nxt = base + 32 + lane < n ? __ldg(cin + rec_index(rec_nxt)) : C{0, 0};
            const PtRec<float> rec_nn = base + 64 + lane < n ? load_rec(recs + base + 64 + lane) : null_rec<float>();
            __syncwarp();                                  // phase B of the previous batch is done with the slots
            int myoff = -2;
            if (lane < cnt) {
                // ---- phase A: kernel values and the operands of phase B, one point per thread
                float kx[NS], ky[NS], kz[DIM == 3 ? NS : 1];
                const int xs = stencil_start(rec_cur.x, NS), ys = stencil_start(rec_cur.y, NS);
                kernel_vector<float, NS, true>(kx, (float)xs - rec_cur.x, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
                kernel_vector<float, NS, true>(ky, (float)ys - rec_cur.y, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
                int zs = 0;
                if (DIM == 3) {
                    zs = stencil_start(rec_cur.z, NS);
                    kernel_vector<float, NS, true>(kz, (float)zs - rec_cur.z, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
                    zs -= a.zshift;                        // weights from the global coordinate, grid index slab-local
                }
                float4 *slot = reinterpret_cast<float4 *>(slots + lane * G::SLOT);
                auto wrow = [&](int row) -> float {
                    if (row >= G::ROWS) return 0.0f;
                    if (DIM == 2) return ky[row];
                    return ky[row % NS] * kz[row / NS];
                };
                auto ckv = [&](int i, int comp) -> float { return i < NS ? (comp ? c_cur.y : c_cur.x) * kx[i] : 0.0f; };
                if constexpr (G::SINGLE) {
#pragma unroll
                    for (int i = 0; i < G::CKSEG / 4; ++i)
                        slot[i] = make_float4(ckv(2 * i, 0), ckv(2 * i, 1), ckv(2 * i + 1, 0), ckv(2 * i + 1, 1));
                    auto wv = [&](int rr) -> float { return rr < G::R ? wrow(rr) : 0.0f; };
#pragma unroll
                    for (int i = 0; i < G::WSEG / 4; ++i)
                        slot[G::CKSEG / 4 + i] = make_float4(wv(2 * i), wv(2 * i), wv(2 * i + 1), wv(2 * i + 1));
                } else {
#pragma unroll
                    for (int i = 0; i < G::XP; ++i)
                        slot[i] = make_float4(ckv(i, 0), ckv(i + G::XP, 0), ckv(i, 1), ckv(i + G::XP, 1));
                    // row slot rr, pass it -> flat float index rr*WS + it inside the W segment
                    auto wv = [&](int f) -> float {
                        const int rr = f / G::WS, it = f - rr * G::WS;
                        if (!(rr < G::R && it < G::NPASS)) return 0.0f;
                        const int row = sm2_row<DIM, NS, RMC>(it, rr);     // compile-time after unrolling
                        return row >= 0 ? wrow(row) : 0.0f;
                    };
#pragma unroll
                    for (int i = 0; i < G::WSEG / 4; ++i)
                        slot[G::CKSEG / 4 + i] = make_float4(wv(4 * i), wv(4 * i + 1), wv(4 * i + 2), wv(4 * i + 3));
                }
                myoff = clampi(xs - ox, 0, a.ex - NS);
                { const int yo = clampi(ys - oy, 0, a.ey - NS); myoff += yo * a.sy; ylo = min(ylo, yo); yhi = max(yhi, yo); }
                if (DIM == 3) { const int zo = clampi(zs - oz, 0, a.ez - NS); myoff += zo * a.sz; zlo = min(zlo, zo); zhi = max(zhi, zo); }
                s_off[lane] = myoff;
            } else if (cnt < 32) {
                // short (last) batch: phase B runs whole groups of four points, the surplus ones add zeros
                // (both operand segments: 0 x a stale NaN would still be NaN)
                float4 *slot = reinterpret_cast<float4 *>(slots + lane * G::SLOT);
#pragma unroll
                for (int i = 0; i < (G::CKSEG + G::WSEG) / 4; ++i) slot[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
            __syncwarp();
            int prev = __shfl_up_sync(0xffffffffu, myoff, 1);
            if (lane == 0) prev = cur;
            const unsigned starts = __ballot_sync(0xffffffffu, lane < cnt && myoff != prev);

            // ---- phase B: one loop over the points, operands one point ahead
            if constexpr (G::SINGLE) {
                float2 A = *reinterpret_cast<const float2 *>(my_ck);
                float2 W = *reinterpret_cast<const float2 *>(my_w);
#pragma unroll 4
                for (int q = 0; q < cnt; ++q) {
                    const float2 An = *reinterpret_cast<const float2 *>(my_ck + (q + 1) * G::SLOT);
                    const float2 Wn = *reinterpret_cast<const float2 *>(my_w + (q + 1) * G::SLOT);
                    if ((starts >> q) & 1u) {
                        if (cur >= 0) flush_run();
                        cur = s_off[q];
                    }
                    fma2(are[0], A, W);
                    A = An; W = Wn;
                }
            } else {
                struct Ops { float4 a; float w[G::WS]; };
                auto load_ops = [&](int q) -> Ops {
                    Ops o;
                    o.a = *reinterpret_cast<const float4 *>(my_ck + q * G::SLOT);
                    const float *pw = my_w + q * G::SLOT;
                    if constexpr (G::WS == 1) o.w[0] = pw[0];
                    else if constexpr (G::WS == 2) { const float2 v = *reinterpret_cast<const float2 *>(pw); o.w[0] = v.x; o.w[1] = v.y; }
                    else {
#pragma unroll
                        for (int j = 0; j < G::WS / 4; ++j) {
                            const float4 v = reinterpret_cast<const float4 *>(pw)[j];
                            o.w[4 * j] = v.x; o.w[4 * j + 1] = v.y; o.w[4 * j + 2] = v.z; o.w[4 * j + 3] = v.w;
                        }
                    }
                    return o;
                };
                auto accumulate = [&](const Ops &o) {
                    const float2 pre = make_float2(o.a.x, o.a.y), pim = make_float2(o.a.z, o.a.w);
#pragma unroll
                    for (int it = 0; it < G::NPASS; ++it) {
                        const float2 ww = make_float2(o.w[it], o.w[it]);
                        fma2(are[it], ww, pre);
                        fma2(aim[it], ww, pim);
                    }
                };
                Ops o = load_ops(0);
                // groups of four points: a group without a run start (about half of them on dense inputs) is
                // straight-line code; the others test every point
                for (int q0 = 0; q0 < cnt; q0 += 4) {
                    const unsigned m = (starts >> q0) & 0xfu;
                    if (m == 0u) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const Ops on = load_ops(q0 + j + 1);
                            accumulate(o);
                            o = on;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const Ops on = load_ops(q0 + j + 1);
                            if ((m >> j) & 1u) {
                                if (cur >= 0) flush_run();
                                cur = s_off[q0 + j];
                            }
                            accumulate(o);
                            o = on;
                        }
                    }
                }
            }
            rec_cur = rec_nxt; rec_nxt = rec_nn; c_cur = c_nxt;
        }
        if (cur >= 0) flush_run();
        __syncwarp();

        // ---- tile -> fine grid over the touched box [nz][ny] x ex: lanes over (rows of the box, x) -- several
        // rows per warp pass when the tile is narrow --, two passes in flight; zero cells are skipped, the tile
        // is cleared on the way.  Boxes that do not cross the periodic boundary (all but the outermost bins)
        // take the path without index wrapping.  (The reference's guard ix < nf + pad,
        // src/2d/spreadinterp2d.cu:222-224, is implied: cells beyond it are never touched and stay zero.)
        {
            ylo = __reduce_min_sync(0xffffffffu, ylo); yhi = __reduce_max_sync(0xffffffffu, yhi);
            if (DIM == 3) { zlo = __reduce_min_sync(0xffffffffu, zlo); zhi = __reduce_max_sync(0xffffffffu, zhi); }
            const int ex = a.ex;
            const int ny = yhi < 0 ? 0 : yhi - ylo + NS;
            const int nz = DIM == 3 ? zhi - zlo + NS : 1;
            const int z0 = DIM == 3 ? zlo : 0;
            const int nrows = ny * nz;
            const int rpi = ex >= 32 ? 1 : 32 / ex;                     // box rows per warp pass
            const int lr = ex >= 32 ? 0 : lane / ex, lx0 = lane - lr * ex;
            const bool lane_on = lr < rpi;
            const int gy0 = oy + ylo, gz0 = oz + z0;
            const bool nowrap = ox >= 0 && ox + ex <= a.nf1 && gy0 >= 0 && gy0 + ny <= a.nf2 &&
                                (DIM == 2 || (gz0 >= 0 && gz0 + nz <= a.nf3));
            const size_t plane = (size_t)a.nf1 * a.nf2;
            auto sweep = [&](auto wrap_tag) {
                constexpr bool WRAP = decltype(wrap_tag)::value;
                int ly = lr, lz = 0;                                   // this lane's row of the current pass
                while (ly >= ny && ny > 0) { ly -= ny; ++lz; }
                auto step = [&]() { ly += rpi; while (ly >= ny) { ly -= ny; ++lz; } };
                auto cell_ptrs = [&](int row, C *&tp, C *&gp, bool &on) {
                    on = lane_on && row < nrows;
                    tp = tile + (z0 + lz) * a.sz + (ylo + ly) * a.sy;
                    if (WRAP) {
                        gp = fwt + (size_t)wrap_index(gy0 + ly, a.nf2) * a.nf1;
                        if (DIM == 3) gp += (size_t)wrap_index(gz0 + lz, a.nf3) * plane;
                    } else {
                        gp = fwt + (size_t)(gy0 + ly) * a.nf1 + ox;
                        if (DIM == 3) gp += (size_t)(gz0 + lz) * plane;
                    }
                };
                for (int row0 = 0; row0 < nrows; row0 += 2 * rpi) {
                    C *t0, *t1, *g0, *g1;
                    bool on0, on1;
                    cell_ptrs(row0 + lr, t0, g0, on0);
                    step();
                    cell_ptrs(row0 + rpi + lr, t1, g1, on1);
                    step();
                    for (int lx = lx0; lx < ex; lx += 32) {
                        C v0 = C{0.0f, 0.0f}, v1 = C{0.0f, 0.0f};
                        if (on0) v0 = t0[lx];
                        if (on1) v1 = t1[lx];
                        const int gx = WRAP ? wrap_index(ox + lx, a.nf1) : lx;
                        if (v0.x != 0.0f || v0.y != 0.0f) { red_add(g0 + gx, v0.x, v0.y); t0[lx] = C{0.0f, 0.0f}; }
                        if (v1.x != 0.0f || v1.y != 0.0f) { red_add(g1 + gx, v1.x, v1.y); t1[lx] = C{0.0f, 0.0f}; }
                    }
                }
            };
            if (nowrap) sweep(std::false_type{});
            else sweep(std::true_type{});
        }
        __syncwarp();
    }
}

}  // namespace cfb
