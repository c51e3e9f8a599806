// spreadinterp.cuh -- spreading (type-1 step 1) and interpolation (type-2 step 3)
// kernels of the B200-native cuFINUFFT hot path.
//
// Reference kernels replaced (all of src/{1,2,3}d/spreadinterp{1,2,3}d.cu):
//   Spread_{1,2,3}d_NUptsdriven[_Horner]  -> spread_gm_kernel   (gpu_method 1, GM / GM-sort)
//   Spread_{1,2,3}d_Subprob[_Horner]      -> spread_sm_kernel   (gpu_method 2, SM)
//   Interp_{1,2,3}d_NUptsdriven[_Horner], Interp_{2,3}d_Subprob[_Horner] -> interp_kernel
//
// Design (DESIGN.md "spread" / "interp").  The reference gives every point to one thread that
// does 2*ns^d scalar atomics (shared or global).  Here
//  * a WARP owns a batch of 32 consecutive sorted points.  Phase A is thread-per-point: one
//    16/32-byte record load, the d*ns kernel values (exp/sqrt or Horner), the strength gather,
//    all parked in a small per-warp scratch.  Phase B is lane-per-cell: the 32 lanes are laid
//    over the point's stencil, lane = (row r, column ix), ITERS passes cover all rows.
//  * setpts sorts points by (bin, stencil origin), so consecutive points usually share their
//    WHOLE stencil.  Such a RUN is accumulated in registers (ITERS complex accumulators per
//    lane) and touches memory once per run instead of once per point:
//      spread SM : run -> warp-private padded bin tile in shared memory with plain LDS/FADD/STS
//                  (lanes of one flush hit distinct cells, padded strides keep them on distinct
//                  banks: no shared atomics at all); tile -> fine grid once per subproblem with
//                  vector RED (red.global.add.v2.f32), zero cells skipped
//      spread GM : run -> fine grid directly with vector RED
//      interp    : the stencil's grid values are loaded once per run (coalesced row segments)
//                  and reused from registers for every point of the run; per-point partial sums
//                  of 8 points are reduced together by one transposing shuffle butterfly
//    With one point per run this degenerates to the per-point version; clustered inputs (the
//    reference's worst case: atomic contention) become the best case.
//  * Subproblems (the reference's (bin, <= maxsubprobsize points) units, same subprob_to_bin
//    map) x transforms are pulled from a global work counter by persistent warps.
// Stencils too large for register accumulators (3-D fp64, ns >= 11) use the same kernels with
// MERGE = false: every pass is applied to memory immediately.
#pragma once
#include "cfb_device.cuh"

namespace cfb {

template <typename T>
struct SIArgs {
    using C = typename cplx_of<T>::type;
    const PtRec<T> *recs;       // sorted point records
    C *c;                       // strengths in (spread) / values out (interp), [nt][M]
    C *fw;                      // fine grids [nt][nf3][nf2][nf1]
    const int *binstart, *binsize, *s2b, *substart, *scalars;
    int *counter;               // global work counter (zeroed before launch)
    const T *hcoef;             // Horner coefficients [ncoef][NS] as T (device)
    int M, nt, maxsub;
    int nf1, nf2, nf3;
    int bs1, bs2, bs3, nb1, nb2;
    int pad, ex, ey, ez;        // tile halo and extents (cells)
    int sy, sz, tile_cells;     // padded tile strides
    int horner, ncoef;
    T es_c, es_beta;
    long long fwstride;
};

template <typename T, int DIM, int NS> struct Geo {
    static constexpr int R = 32 / NS;                                    // stencil rows per warp pass
    static constexpr int LANES = R * NS;                                 // active lanes
    static constexpr int ROWS = DIM == 1 ? 1 : (DIM == 2 ? NS : NS * NS);
    static constexpr int ITERS = (ROWS + R - 1) / R;
    static constexpr bool PROD = DIM == 3 && NS <= 7;                    // ky*kz products precomputed per point
    static constexpr int NW = DIM == 1 ? 0 : ((DIM == 2 || PROD) ? ITERS * R : 2 * NS);   // row weights (zero padded)
    static constexpr int KP = (NS + NW) | 1;                             // odd stride: conflict-free scratch
    static constexpr bool MERGE = ITERS * (int)(sizeof(T) / 4) <= 72;    // run accumulators fit in registers
    static constexpr bool TOFF_REGS = sizeof(T) == 4 || ITERS <= 16;     // per-pass tile offsets kept in registers
    static constexpr int SM_MAXW = ITERS * (int)(sizeof(T) / 4) > 24 ? 4 : 16;   // warps per SM-spread block (register budget)
};

// host mirror of Geo::SM_MAXW for the plan-time tile geometry
inline int sm_spread_max_warps(int dim, int ns, int real_bytes)
{
    const int R = 32 / ns, rows = dim == 1 ? 1 : (dim == 2 ? ns : ns * ns), iters = (rows + R - 1) / R;
    return iters * (real_bytes / 4) > 24 ? 4 : 16;
}

// bytes of per-warp scratch: ker[32*KP] T | c[32] C | x0,y0,z0[32] int
template <typename T, int DIM, int NS>
__host__ __device__ constexpr size_t warp_scratch_bytes()
{
    return (size_t)32 * Geo<T, DIM, NS>::KP * sizeof(T) + 32 * 2 * sizeof(T) + 3 * 32 * sizeof(int);
}

template <typename T, int DIM, int NS>
struct Scratch {
    T *ker; typename cplx_of<T>::type *c; int *x0, *y0, *z0;
    __device__ __forceinline__ explicit Scratch(unsigned char *base)
    {
        using G = Geo<T, DIM, NS>;
        ker = reinterpret_cast<T *>(base);
        c = reinterpret_cast<typename cplx_of<T>::type *>(ker + 32 * G::KP);
        x0 = reinterpret_cast<int *>(c + 32);
        y0 = x0 + 32;
        z0 = y0 + 32;
    }
};

// ---- phase A: thread-per-point kernel weights into the per-warp scratch ------
// ker row of the point: [0,NS) x weights; then row weights: 2-D ky[NS] (zero padded to ITERS*R);
// 3-D PROD ky[iy]*kz[iz] at iz*NS+iy (zero padded); 3-D !PROD ky[NS], kz[NS].
template <typename T, int DIM, int NS>
__device__ __forceinline__ void point_weights(const SIArgs<T> &a, const PtRec<T> &rec, T *kp, const T *s_hc,
                                              int &xstart, int &ystart, int &zstart)
{
    using G = Geo<T, DIM, NS>;
    T ker[NS];
    xstart = stencil_start(rec.x, NS);
    kernel_vector<T, NS>(ker, (T)xstart - rec.x, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
#pragma unroll
    for (int i = 0; i < NS; ++i) kp[i] = ker[i];
    if (DIM == 2) {
        ystart = stencil_start(rec.y, NS);
        kernel_vector<T, NS>(ker, (T)ystart - rec.y, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
#pragma unroll
        for (int i = 0; i < NS; ++i) kp[NS + i] = ker[i];
#pragma unroll
        for (int i = NS; i < G::NW; ++i) kp[NS + i] = 0;
    }
    if (DIM == 3) {
        T kz[NS];
        ystart = stencil_start(rec.y, NS);
        kernel_vector<T, NS>(ker, (T)ystart - rec.y, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
        zstart = stencil_start(rec.z, NS);
        kernel_vector<T, NS>(kz, (T)zstart - rec.z, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
        if (G::PROD) {
#pragma unroll
            for (int iz = 0; iz < NS; ++iz)
#pragma unroll
                for (int iy = 0; iy < NS; ++iy) kp[NS + iz * NS + iy] = ker[iy] * kz[iz];
#pragma unroll
            for (int i = NS * NS; i < G::NW; ++i) kp[NS + i] = 0;
        } else {
#pragma unroll
            for (int i = 0; i < NS; ++i) { kp[NS + i] = ker[i]; kp[2 * NS + i] = kz[i]; }
        }
    }
}

// row weight of pass `it` for this lane (row = it*R + r)
template <typename T, int DIM, int NS>
__device__ __forceinline__ T row_weight(const T *kq, int it, int r)
{
    using G = Geo<T, DIM, NS>;
    if (DIM == 1) return (T)1;
    const int row = it * G::R + r;
    if (DIM == 2 || G::PROD) return kq[NS + row];
    const int iz = row / NS, iy = row - iz * NS;
    return row < G::ROWS ? kq[NS + iy] * kq[2 * NS + iz] : (T)0;
}

template <typename T, int NS>
__device__ __forceinline__ const T *stage_horner(const SIArgs<T> &a, T *s_hc)
{
    if (a.horner)
        for (int i = threadIdx.x; i < a.ncoef * NS; i += blockDim.x) s_hc[i] = a.hcoef[i];
    __syncthreads();
    return s_hc;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// =============================================================================
// SM spread: warp-private tile, run accumulation in registers, lane-per-cell flushes.
// dynamic smem: [hcoef 18*16 T][per warp: tile C[tile_cells] | scratch]
// =============================================================================
template <typename T, int DIM, int NS>
__global__ void __launch_bounds__(32 * Geo<T, DIM, NS>::SM_MAXW)
spread_sm_kernel(const SIArgs<T> a)
{
    using C = typename cplx_of<T>::type;
    using G = Geo<T, DIM, NS>;
    extern __shared__ __align__(16) unsigned char smem[];
    T *s_hc = reinterpret_cast<T *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = (size_t)a.tile_cells * sizeof(C) + warp_scratch_bytes<T, DIM, NS>();
    unsigned char *wbase = smem + 18 * 16 * sizeof(T) + warp * per_warp;
    C *tile = reinterpret_cast<C *>(wbase);
    Scratch<T, DIM, NS> sc(wbase + (size_t)a.tile_cells * sizeof(C));
    int *s_off = sc.x0;
    stage_horner<T, NS>(a, s_hc);

    const bool active = lane < G::LANES;
    const int r = active ? lane / NS : 0, ix = active ? lane - r * NS : 0;
    const int nsub = a.scalars[0];
    const long long total = (long long)nsub * a.nt;

    int toff[G::TOFF_REGS ? G::ITERS : 1];
    if (G::TOFF_REGS) {
#pragma unroll
        for (int it = 0; it < G::ITERS; ++it) {
            const int row = it * G::R + r;
            if (DIM == 1) toff[it] = 0;
            else if (DIM == 2) toff[it] = row * a.sy;
            else { const int iz = row / NS, iy = row - iz * NS; toff[it] = iz * a.sz + iy * a.sy; }
        }
    }
    auto tile_off = [&](int it) -> int {
        if (G::TOFF_REGS) return toff[it];
        const int row = it * G::R + r;
        if (DIM == 1) return 0;
        if (DIM == 2) return row * a.sy;
        const int iz = row / NS, iy = row - iz * NS;
        return iz * a.sz + iy * a.sy;
    };

    for (;;) {
        long long w = 0;
        if (lane == 0) w = atomicAdd(a.counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total) break;
        const int t = (int)(w / nsub), s = (int)(w - (long long)t * nsub);
        const int bin = a.s2b[s];
        const int k = s - a.substart[bin];
        const int pstart = a.binstart[bin] + k * a.maxsub;
        const int n = min(a.maxsub, a.binsize[bin] - k * a.maxsub);
        int b1 = bin % a.nb1, b23 = bin / a.nb1;
        int b2 = DIM > 1 ? b23 % a.nb2 : 0, b3 = DIM > 2 ? b23 / a.nb2 : 0;
        const int ox = b1 * a.bs1 - a.pad, oy = b2 * a.bs2 - a.pad, oz = b3 * a.bs3 - a.pad;
        const C *cin = a.c + (size_t)t * a.M;
        C *fwt = a.fw + (size_t)t * a.fwstride;

        for (int i = lane; i < a.tile_cells; i += 32) { tile[i].x = 0; tile[i].y = 0; }

        C acc[G::MERGE ? G::ITERS : 1];
        if (G::MERGE) {
#pragma unroll
            for (int it = 0; it < G::ITERS; ++it) { acc[it].x = 0; acc[it].y = 0; }
        }
        int cur = -1;                                   // tile offset of the open run

        // add the open run's accumulators into the tile (lanes touch distinct cells)
        auto flush_run = [&]() {
            if constexpr (G::MERGE) {
                __syncwarp();
                C *cell0 = tile + cur + ix;
#pragma unroll
                for (int it = 0; it < G::ITERS; ++it) {
                    if (active && it * G::R + r < G::ROWS) {
                        C *cell = cell0 + tile_off(it);
                        C v = *cell;
                        v.x += acc[it].x; v.y += acc[it].y;
                        *cell = v;
                    }
                    acc[it].x = 0; acc[it].y = 0;
                }
            }
        };

        for (int base = 0; base < n; base += 32) {
            const int cnt = min(32, n - base);
            __syncwarp();
            int myoff = -2;
            if (lane < cnt) {
                const PtRec<T> rec = load_rec(a.recs + pstart + base + lane);
                int xs0, ys0 = 0, zs0 = 0;
                point_weights<T, DIM, NS>(a, rec, sc.ker + lane * G::KP, s_hc, xs0, ys0, zs0);
                sc.c[lane] = cin[rec_index(rec)];
                myoff = clampi(xs0 - ox, 0, a.ex - NS);
                if (DIM > 1) myoff += clampi(ys0 - oy, 0, a.ey - NS) * a.sy;
                if (DIM > 2) myoff += clampi(zs0 - oz, 0, a.ez - NS) * a.sz;
                s_off[lane] = myoff;
            }
            __syncwarp();
            if constexpr (G::MERGE) {
                int prev = __shfl_up_sync(0xffffffffu, myoff, 1);
                if (lane == 0) prev = cur;
                const unsigned starts = __ballot_sync(0xffffffffu, lane < cnt && myoff != prev);
                for (int q = 0; q < cnt; ++q) {
                    if ((starts >> q) & 1u) {
                        if (cur >= 0) flush_run();
                        cur = s_off[q];
                    }
                    const T *kq = sc.ker + q * G::KP;
                    const T k1 = active ? kq[ix] : (T)0;
                    const C cv = sc.c[q];
                    const T cr = cv.x * k1, ci = cv.y * k1;
#pragma unroll
                    for (int it = 0; it < G::ITERS; ++it) {
                        const T wgt = row_weight<T, DIM, NS>(kq, it, r);
                        acc[it].x = fma(cr, wgt, acc[it].x);
                        acc[it].y = fma(ci, wgt, acc[it].y);
                    }
                }
            } else {
                // stencil too large for register accumulators: apply every pass to the tile at once
                for (int q = 0; q < cnt; ++q) {
                    const T *kq = sc.ker + q * G::KP;
                    const T k1 = active ? kq[ix] : (T)0;
                    const C cv = sc.c[q];
                    const T cr = cv.x * k1, ci = cv.y * k1;
                    C *cell0 = tile + s_off[q] + ix;
#pragma unroll 4
                    for (int it = 0; it < G::ITERS; ++it) {
                        if (active && it * G::R + r < G::ROWS) {
                            const T wgt = row_weight<T, DIM, NS>(kq, it, r);
                            C *cell = cell0 + tile_off(it);
                            C v = *cell;
                            v.x = fma(cr, wgt, v.x); v.y = fma(ci, wgt, v.y);
                            *cell = v;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (cur >= 0) flush_run();
        __syncwarp();

        // flush: vector RED of touched cells, single periodic wrap (reference guard
        // ix < nf+pad, src/2d/spreadinterp2d.cu:222-224, is implied: cells past it stay zero)
        {
            int lx = lane, ly = 0, lz = 0;
            while (lx >= a.sy) { lx -= a.sy; ++ly; }
            for (int i = lane; i < a.tile_cells; i += 32) {
                if (DIM > 2) { const int rows_per_z = a.sz / a.sy; while (ly >= rows_per_z) { ly -= rows_per_z; ++lz; } }
                C v = tile[i];
                if ((v.x != 0 || v.y != 0) && lx < a.ex) {
                    int gx = wrap_index(ox + lx, a.nf1);
                    size_t o = gx;
                    if (DIM > 1) o += (size_t)wrap_index(oy + ly, a.nf2) * a.nf1;
                    if (DIM > 2) o += (size_t)wrap_index(oz + lz, a.nf3) * a.nf1 * a.nf2;
                    red_add(fwt + o, v.x, v.y);
                }
                lx += 32;
                while (lx >= a.sy) { lx -= a.sy; ++ly; }
            }
        }
        __syncwarp();
    }
}

// =============================================================================
// GM / GM-sort spread: same lane-per-cell mapping and run accumulation, runs go straight
// into the fine grid with vector RED (no tile).  Work unit = 32 consecutive points.
// =============================================================================
template <typename T, int DIM, int NS>
__global__ void __launch_bounds__(256)
spread_gm_kernel(const SIArgs<T> a)
{
    using C = typename cplx_of<T>::type;
    using G = Geo<T, DIM, NS>;
    extern __shared__ __align__(16) unsigned char smem[];
    T *s_hc = reinterpret_cast<T *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Scratch<T, DIM, NS> sc(smem + 18 * 16 * sizeof(T) + warp * warp_scratch_bytes<T, DIM, NS>());
    stage_horner<T, NS>(a, s_hc);

    const bool active = lane < G::LANES;
    const int r = active ? lane / NS : 0, ix = active ? lane - r * NS : 0;
    // a warp takes CH consecutive batches so that runs survive batch boundaries
    constexpr int CH = 8;
    const long long nchunk = ((long long)a.M + 32 * CH - 1) / (32 * CH);
    const long long total = nchunk * a.nt;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    const size_t plane = (size_t)a.nf1 * a.nf2;

    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp; w < total; w += wstride) {
        const int t = (int)(w / nchunk);
        const long long p0 = (w - (long long)t * nchunk) * (32 * CH);
        const int n = (int)min((long long)(32 * CH), a.M - p0);
        const C *cin = a.c + (size_t)t * a.M;
        C *fwt = a.fw + (size_t)t * a.fwstride;

        C acc[G::MERGE ? G::ITERS : 1];
        if (G::MERGE) {
#pragma unroll
            for (int it = 0; it < G::ITERS; ++it) { acc[it].x = 0; acc[it].y = 0; }
        }
        int cx = 0, cy = 0, cz = 0;
        bool open = false;

        auto flush_run = [&]() {
            if constexpr (G::MERGE) {
                const int gx = wrap_index(cx + ix, a.nf1);
#pragma unroll
                for (int it = 0; it < G::ITERS; ++it) {
                    const int row = it * G::R + r;
                    if (active && row < G::ROWS && (acc[it].x != 0 || acc[it].y != 0)) {
                        size_t o = gx;
                        if (DIM == 2) o += (size_t)wrap_index(cy + row, a.nf2) * a.nf1;
                        if (DIM == 3) { const int iz = row / NS, iy = row - iz * NS;
                                        o += (size_t)wrap_index(cy + iy, a.nf2) * a.nf1 + (size_t)wrap_index(cz + iz, a.nf3) * plane; }
                        red_add(fwt + o, acc[it].x, acc[it].y);
                    }
                    acc[it].x = 0; acc[it].y = 0;
                }
            }
        };

        for (int base = 0; base < n; base += 32) {
            const int cnt = min(32, n - base);
            __syncwarp();
            int mx = 0, my = 0, mz = 0;
            if (lane < cnt) {
                const PtRec<T> rec = load_rec(a.recs + p0 + base + lane);
                point_weights<T, DIM, NS>(a, rec, sc.ker + lane * G::KP, s_hc, mx, my, mz);
                sc.c[lane] = cin[rec_index(rec)];
                mx = clampi(mx, -a.nf1, a.nf1); my = clampi(my, -a.nf2, a.nf2); mz = clampi(mz, -a.nf3, a.nf3);
                sc.x0[lane] = mx; sc.y0[lane] = my; sc.z0[lane] = mz;
            }
            __syncwarp();
            int px = __shfl_up_sync(0xffffffffu, mx, 1), py = __shfl_up_sync(0xffffffffu, my, 1), pz = __shfl_up_sync(0xffffffffu, mz, 1);
            if (lane == 0) { px = cx; py = cy; pz = cz; }
            const bool differs = mx != px || my != py || mz != pz || (lane == 0 && !open);
            const unsigned starts = G::MERGE ? __ballot_sync(0xffffffffu, lane < cnt && differs) : 0xffffffffu;
            for (int q = 0; q < cnt; ++q) {
                const T *kq = sc.ker + q * G::KP;
                const T k1 = active ? kq[ix] : (T)0;
                const C cv = sc.c[q];
                const T cr = cv.x * k1, ci = cv.y * k1;
                if constexpr (G::MERGE) {
                    if ((starts >> q) & 1u) {
                        if (open) flush_run();
                        cx = sc.x0[q]; cy = sc.y0[q]; cz = sc.z0[q];
                        open = true;
                    }
#pragma unroll
                    for (int it = 0; it < G::ITERS; ++it) {
                        const T wgt = row_weight<T, DIM, NS>(kq, it, r);
                        acc[it].x = fma(cr, wgt, acc[it].x);
                        acc[it].y = fma(ci, wgt, acc[it].y);
                    }
                } else {
                    const int gx = wrap_index(sc.x0[q] + ix, a.nf1);
                    const int y0 = sc.y0[q], z0 = sc.z0[q];
#pragma unroll 4
                    for (int it = 0; it < G::ITERS; ++it) {
                        const int row = it * G::R + r;
                        if (active && row < G::ROWS) {
                            const T wgt = row_weight<T, DIM, NS>(kq, it, r);
                            size_t o = gx;
                            if (DIM == 2) o += (size_t)wrap_index(y0 + row, a.nf2) * a.nf1;
                            if (DIM == 3) { const int iz = row / NS, iy = row - iz * NS;
                                            o += (size_t)wrap_index(y0 + iy, a.nf2) * a.nf1 + (size_t)wrap_index(z0 + iz, a.nf3) * plane; }
                            red_add(fwt + o, cr * wgt, ci * wgt);
                        }
                    }
                }
            }
        }
        if (open) flush_run();
        __syncwarp();
    }
}

// =============================================================================
// Interpolation: lanes over the stencil (lane = (row r, column ix)); the grid values of a run's
// stencil are loaded once (coalesced row segments; points are sorted so neighbouring runs reuse
// L1/L2 lines) and kept in registers for all points of the run.  Each lane accumulates its
// column over the passes with the row weights (2 FMA per cell), scales by its x-weight once,
// and the 32 partial sums of EIGHT points are reduced together by a transposing butterfly
// (9 shuffles per component per 8 points instead of 40).  Result scattered to c[index].
// =============================================================================
template <typename T>
__device__ __forceinline__ T reduce8(T (&v)[8], int lane)
{
    // after the call every lane holds the full sum of point  4*bit4 + 2*bit3 + bit2  of its lane id
    bool up = lane & 16;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        T send = up ? v[j] : v[j + 4], keep = up ? v[j + 4] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    up = lane & 8;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        T send = up ? v[j] : v[j + 2], keep = up ? v[j + 2] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    up = lane & 4;
    {
        T send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

template <typename T, int DIM, int NS>
__global__ void __launch_bounds__(256)
interp_kernel(const SIArgs<T> a)
{
    using C = typename cplx_of<T>::type;
    using G = Geo<T, DIM, NS>;
    extern __shared__ __align__(16) unsigned char smem[];
    T *s_hc = reinterpret_cast<T *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Scratch<T, DIM, NS> sc(smem + 18 * 16 * sizeof(T) + warp * warp_scratch_bytes<T, DIM, NS>());
    int *s_idx = reinterpret_cast<int *>(sc.c);       // original index of the 32 points of the batch
    stage_horner<T, NS>(a, s_hc);

    const bool active = lane < G::LANES;
    const int r = active ? lane / NS : 0, ix = active ? lane - r * NS : 0;
    constexpr int CH = 8;
    const long long nchunk = ((long long)a.M + 32 * CH - 1) / (32 * CH);
    const long long total = nchunk * a.nt;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    const size_t plane = (size_t)a.nf1 * a.nf2;

    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp; w < total; w += wstride) {
        const int t = (int)(w / nchunk);
        const long long p0 = (w - (long long)t * nchunk) * (32 * CH);
        const int n = (int)min((long long)(32 * CH), a.M - p0);
        C *cout = a.c + (size_t)t * a.M;
        const C *fwt = a.fw + (size_t)t * a.fwstride;

        C v[G::MERGE ? G::ITERS : 1];
        int cx = 0, cy = 0, cz = 0;
        bool open = false;

        auto load_run = [&]() {
            if constexpr (G::MERGE) {
                const C *col = fwt + wrap_index(cx + ix, a.nf1);
#pragma unroll
                for (int it = 0; it < G::ITERS; ++it) {
                    const int row = it * G::R + r;
                    v[it].x = 0; v[it].y = 0;
                    if (active && row < G::ROWS) {
                        size_t o = 0;
                        if (DIM == 2) o = (size_t)wrap_index(cy + row, a.nf2) * a.nf1;
                        if (DIM == 3) { const int iz = row / NS, iy = row - iz * NS;
                                        o = (size_t)wrap_index(cy + iy, a.nf2) * a.nf1 + (size_t)wrap_index(cz + iz, a.nf3) * plane; }
                        v[it] = col[o];
                    }
                }
            }
        };

        for (int base = 0; base < n; base += 32) {
            const int cnt = min(32, n - base);
            __syncwarp();
            int mx = 0, my = 0, mz = 0;
            if (lane < cnt) {
                const PtRec<T> rec = load_rec(a.recs + p0 + base + lane);
                point_weights<T, DIM, NS>(a, rec, sc.ker + lane * G::KP, s_hc, mx, my, mz);
                s_idx[lane] = rec_index(rec);
                mx = clampi(mx, -a.nf1, a.nf1); my = clampi(my, -a.nf2, a.nf2); mz = clampi(mz, -a.nf3, a.nf3);
                sc.x0[lane] = mx; sc.y0[lane] = my; sc.z0[lane] = mz;
            }
            __syncwarp();
            int px = __shfl_up_sync(0xffffffffu, mx, 1), py = __shfl_up_sync(0xffffffffu, my, 1), pz = __shfl_up_sync(0xffffffffu, mz, 1);
            if (lane == 0) { px = cx; py = cy; pz = cz; }
            const bool differs = mx != px || my != py || mz != pz || (lane == 0 && !open);
            const unsigned starts = G::MERGE ? __ballot_sync(0xffffffffu, lane < cnt && differs) : 0xffffffffu;

            for (int q0 = 0; q0 < cnt; q0 += 8) {
                T accr[8], acci[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    accr[j] = 0; acci[j] = 0;
                    const int q = q0 + j;
                    if (q < cnt) {
                        const T *kq = sc.ker + q * G::KP;
                        T sr = 0, si = 0;
                        if constexpr (G::MERGE) {
                            if ((starts >> q) & 1u) {
                                cx = sc.x0[q]; cy = sc.y0[q]; cz = sc.z0[q];
                                open = true;
                                load_run();
                            }
#pragma unroll
                            for (int it = 0; it < G::ITERS; ++it) {
                                const T wgt = row_weight<T, DIM, NS>(kq, it, r);
                                sr = fma(v[it].x, wgt, sr);
                                si = fma(v[it].y, wgt, si);
                            }
                        } else {
                            const C *col = fwt + wrap_index(sc.x0[q] + ix, a.nf1);
                            const int y0 = sc.y0[q], z0 = sc.z0[q];
#pragma unroll 4
                            for (int it = 0; it < G::ITERS; ++it) {
                                const int row = it * G::R + r;
                                if (row < G::ROWS) {
                                    const T wgt = row_weight<T, DIM, NS>(kq, it, r);
                                    size_t o = 0;
                                    if (DIM == 2) o = (size_t)wrap_index(y0 + row, a.nf2) * a.nf1;
                                    if (DIM == 3) { const int iz = row / NS, iy = row - iz * NS;
                                                    o = (size_t)wrap_index(y0 + iy, a.nf2) * a.nf1 + (size_t)wrap_index(z0 + iz, a.nf3) * plane; }
                                    const C g = col[o];
                                    sr = fma(g.x, wgt, sr);
                                    si = fma(g.y, wgt, si);
                                }
                            }
                        }
                        const T k1 = active ? kq[ix] : (T)0;
                        accr[j] = sr * k1; acci[j] = si * k1;
                    }
                }
                const T tr = reduce8(accr, lane), ti = reduce8(acci, lane);
                // lane bits (4,3,2) select the point in reduce8's order: 4*bit4 + 2*bit3 + bit2
                const int qsel = q0 + (((lane >> 4) & 1) << 2) + (((lane >> 3) & 1) << 1) + ((lane >> 2) & 1);
                if ((lane & 3) == 0 && qsel < cnt) { C o; o.x = tr; o.y = ti; cout[s_idx[qsel]] = o; }
            }
        }
        __syncwarp();
    }
}

// ---- host-side launch helpers (one instantiation per (T, DIM) translation unit) ----
template <typename T, int DIM> int launch_spread(Plan<T> &p, const typename Plan<T>::C *c, typename Plan<T>::C *fw, int nt);
template <typename T, int DIM> int launch_interp(Plan<T> &p, typename Plan<T>::C *c, const typename Plan<T>::C *fw, int nt);
template <typename T, int DIM> size_t sm_spread_smem_per_warp(int ns, int tile_cells);

}  // namespace cfb
