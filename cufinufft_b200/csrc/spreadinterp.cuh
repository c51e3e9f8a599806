// spreadinterp.cuh -- spreading (type-1 step 1) and interpolation (type-2 step 3)
// kernels of the B200-native cuFINUFFT hot path.
//
// Reference kernels replaced (all of src/{1,2,3}d/spreadinterp{1,2,3}d.cu):
//   Spread_{1,2,3}d_NUptsdriven[_Horner]  -> spread_gm_kernel   (gpu_method 1, GM / GM-sort)
//   Spread_{1,2,3}d_Subprob[_Horner]      -> spread_sm_kernel   (gpu_method 2, SM)
//   Interp_{1,2,3}d_NUptsdriven[_Horner], Interp_{2,3}d_Subprob[_Horner] -> interp_kernel
//
// Design (DESIGN.md "spread"): on sm_100a a shared-memory fp32/fp64 atomicAdd is a
// CAS spin loop (ATOMS.CAST.SPIN in SASS), ~2 clk per *lane* even without conflicts.
// So instead of one thread per point doing 2*ns^d shared atomics (the reference),
// every WARP owns a private padded bin tile in shared memory (B200 has 227 KB per
// SM: 16 private 2-D tiles fit) and processes its subproblem's points one at a
// time with the 32 lanes laid over the point's stencil: lane = (row r, column ix).
// Within one point all lanes touch distinct cells and the padded strides make the
// access bank-conflict free, so accumulation is a plain LDS / FFMA / STS sequence --
// no atomics at all in shared memory.  Kernel weights are evaluated thread-per-point
// (32 points at a time, all lanes busy) and handed to the lane-per-cell phase
// through a small per-warp scratch.  Tiles are flushed with vector RED
// (red.global.add.v2.f32) skipping untouched cells.  Subproblems (the reference's
// (bin, <=maxsubprobsize points) units, same subprob_to_bin map) x transforms are
// pulled from a global work counter by persistent warps.
#pragma once
#include "cfb_device.cuh"

namespace cfb {

template <typename T>
struct SIArgs {
    using C = typename cplx_of<T>::type;
    const T *xs, *ys, *zs;      // bin-ordered rescaled coordinates
    const int *idx;             // idxnupts
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

template <int DIM, int NS> struct Geo {
    static constexpr int R = 32 / NS;                                    // stencil rows per warp pass
    static constexpr int ROWS = DIM == 1 ? 1 : (DIM == 2 ? NS : NS * NS);
    static constexpr int ITERS = (ROWS + R - 1) / R;
    static constexpr int KV = DIM * NS;                                  // kernel values per point
    static constexpr int KVP = KV | 1;                                   // odd stride: conflict-free scratch
};

// bytes of per-warp scratch: off[32] + cre[32] + cim[32] + ker[32*KVP]
template <typename T, int DIM, int NS>
__host__ __device__ constexpr size_t warp_scratch_bytes()
{
    return 32 * sizeof(int) + 2 * 32 * sizeof(T) + 32 * Geo<DIM, NS>::KVP * sizeof(T);
}

// ---- phase A: thread-per-point kernel weights into the per-warp scratch ------
template <typename T, int DIM, int NS>
__device__ __forceinline__ void point_weights(const SIArgs<T> &a, int p, T *s_ker_lane, const T *s_hc,
                                              int &xstart, int &ystart, int &zstart)
{
    T ker[NS];
    T xr = a.xs[p];
    xstart = stencil_start(xr, NS);
    kernel_vector<T, NS>(ker, (T)xstart - xr, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
#pragma unroll
    for (int i = 0; i < NS; ++i) s_ker_lane[i] = ker[i];
    if (DIM > 1) {
        T yr = a.ys[p];
        ystart = stencil_start(yr, NS);
        kernel_vector<T, NS>(ker, (T)ystart - yr, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
#pragma unroll
        for (int i = 0; i < NS; ++i) s_ker_lane[NS + i] = ker[i];
    }
    if (DIM > 2) {
        T zr = a.zs[p];
        zstart = stencil_start(zr, NS);
        kernel_vector<T, NS>(ker, (T)zstart - zr, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
#pragma unroll
        for (int i = 0; i < NS; ++i) s_ker_lane[2 * NS + i] = ker[i];
    }
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
// SM spread: warp-private tile, lane-per-cell accumulation without atomics.
// dynamic smem: [hcoef 18*16 T][per warp: tile C[tile_cells] | scratch]
// =============================================================================
template <typename T, int DIM, int NS>
__global__ void __launch_bounds__(512)
spread_sm_kernel(const SIArgs<T> a)
{
    using C = typename cplx_of<T>::type;
    using G = Geo<DIM, NS>;
    extern __shared__ __align__(16) unsigned char smem[];
    T *s_hc = reinterpret_cast<T *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = (size_t)a.tile_cells * sizeof(C) + warp_scratch_bytes<T, DIM, NS>();
    unsigned char *wbase = smem + 18 * 16 * sizeof(T) + warp * per_warp;
    C *tile = reinterpret_cast<C *>(wbase);
    T *s_cre = reinterpret_cast<T *>(wbase + (size_t)a.tile_cells * sizeof(C));
    T *s_cim = s_cre + 32;
    T *s_ker = s_cim + 32;
    int *s_off = reinterpret_cast<int *>(s_ker + 32 * G::KVP);
    stage_horner<T, NS>(a, s_hc);

    const int r = lane / NS, ix = lane - r * NS;
    const bool active = lane < G::R * NS;
    const int nsub = a.scalars[0];
    const long long total = (long long)nsub * a.nt;

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
        __syncwarp();

        for (int base = 0; base < n; base += 32) {
            const int cnt = min(32, n - base);
            if (lane < cnt) {
                const int p = pstart + base + lane;
                int xs0, ys0 = 0, zs0 = 0;
                point_weights<T, DIM, NS>(a, p, s_ker + lane * G::KVP, s_hc, xs0, ys0, zs0);
                C cv = cin[a.idx[p]];
                s_cre[lane] = cv.x; s_cim[lane] = cv.y;
                int off = clampi(xs0 - ox, 0, a.ex - NS);
                if (DIM > 1) off += clampi(ys0 - oy, 0, a.ey - NS) * a.sy;
                if (DIM > 2) off += clampi(zs0 - oz, 0, a.ez - NS) * a.sz;
                s_off[lane] = off;
            }
            __syncwarp();
            // lane-per-cell accumulation.  All loads of a chunk of passes (weights + tile cells)
            // are issued before any store so they overlap (the compiler must otherwise assume
            // the tile store of one pass aliases the loads of the next), and the next point's
            // header is fetched before this point's stores.
            int off_n = s_off[0];
            T k1_n = s_ker[ix], cr_n = s_cre[0], ci_n = s_cim[0];
            for (int q = 0; q < cnt; ++q) {
                const T *kq = s_ker + q * G::KVP;
                const T cr = cr_n * k1_n, ci = ci_n * k1_n;
                C *cell0 = tile + off_n + ix;
                const int qn = q + 1 < cnt ? q + 1 : q;
                off_n = s_off[qn]; k1_n = s_ker[qn * G::KVP + ix]; cr_n = s_cre[qn]; ci_n = s_cim[qn];
                constexpr int CH = G::ITERS < 8 ? G::ITERS : 8;
#pragma unroll(G::ITERS <= 16 ? 16 : 1)
                for (int it0 = 0; it0 < G::ITERS; it0 += CH) {
                    C v[CH]; T wgt[CH]; int toff[CH]; bool ok[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        const int row = (it0 + j) * G::R + r;
                        ok[j] = active && row < G::ROWS && (it0 + j) < G::ITERS;
                        if (DIM == 1) { wgt[j] = 1; toff[j] = 0; }
                        else if (DIM == 2) { toff[j] = row * a.sy; wgt[j] = ok[j] ? kq[NS + row] : (T)0; }
                        else { const int iz = row / NS, iy = row - iz * NS;
                               toff[j] = iz * a.sz + iy * a.sy;
                               wgt[j] = ok[j] ? kq[NS + iy] * kq[2 * NS + iz] : (T)0; }
                    }
#pragma unroll
                    for (int j = 0; j < CH; ++j) { if (ok[j]) v[j] = cell0[toff[j]]; else { v[j].x = 0; v[j].y = 0; } }
#pragma unroll
                    for (int j = 0; j < CH; ++j) { v[j].x = fma(cr, wgt[j], v[j].x); v[j].y = fma(ci, wgt[j], v[j].y); }
#pragma unroll
                    for (int j = 0; j < CH; ++j) if (ok[j]) cell0[toff[j]] = v[j];
                }
                __syncwarp();
            }
        }

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
// GM / GM-sort spread: same lane-per-cell mapping, accumulation straight into the
// fine grid with vector RED (no tile).  Work unit = 32 consecutive (sorted) points.
// =============================================================================
template <typename T, int DIM, int NS>
__global__ void __launch_bounds__(256)
spread_gm_kernel(const SIArgs<T> a)
{
    using C = typename cplx_of<T>::type;
    using G = Geo<DIM, NS>;
    extern __shared__ __align__(16) unsigned char smem[];
    T *s_hc = reinterpret_cast<T *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wbase = smem + 18 * 16 * sizeof(T) + warp * (warp_scratch_bytes<T, DIM, NS>() + 2 * 32 * sizeof(int));
    T *s_cre = reinterpret_cast<T *>(wbase);
    T *s_cim = s_cre + 32;
    T *s_ker = s_cim + 32;
    int *s_x0 = reinterpret_cast<int *>(s_ker + 32 * G::KVP);
    int *s_y0 = s_x0 + 32, *s_z0 = s_y0 + 32;
    stage_horner<T, NS>(a, s_hc);

    const int r = lane / NS, ix = lane - r * NS;
    const bool active = lane < G::R * NS;
    const long long nbatch = ((long long)a.M + 31) / 32;
    const long long total = nbatch * a.nt;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);

    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp; w < total; w += wstride) {
        const int t = (int)(w / nbatch);
        const long long pb = (w - (long long)t * nbatch) * 32;
        const int cnt = (int)min((long long)32, a.M - pb);
        const C *cin = a.c + (size_t)t * a.M;
        C *fwt = a.fw + (size_t)t * a.fwstride;
        if (lane < cnt) {
            const int p = (int)(pb + lane);
            int xs0, ys0 = 0, zs0 = 0;
            point_weights<T, DIM, NS>(a, p, s_ker + lane * G::KVP, s_hc, xs0, ys0, zs0);
            C cv = cin[a.idx[p]];
            s_cre[lane] = cv.x; s_cim[lane] = cv.y;
            s_x0[lane] = xs0; s_y0[lane] = ys0; s_z0[lane] = zs0;
        }
        __syncwarp();
        if (active) {
            for (int q = 0; q < cnt; ++q) {
                const T *kq = s_ker + q * G::KVP;
                const T k1 = kq[ix];
                const T cr = s_cre[q] * k1, ci = s_cim[q] * k1;
                const int gx = wrap_index(clampi(s_x0[q], -a.nf1, a.nf1) + ix, a.nf1);
                const int y0 = clampi(s_y0[q], -a.nf2, a.nf2), z0 = clampi(s_z0[q], -a.nf3, a.nf3);
#pragma unroll
                for (int it = 0; it < G::ITERS; ++it) {
                    const int row = it * G::R + r;
                    if (row < G::ROWS) {
                        T wgt = 1; size_t o = gx;
                        if (DIM == 2) { wgt = kq[NS + row]; o += (size_t)wrap_index(y0 + row, a.nf2) * a.nf1; }
                        if (DIM == 3) { const int iz = row / NS, iy = row - iz * NS;
                                        wgt = kq[NS + iy] * kq[2 * NS + iz];
                                        o += (size_t)wrap_index(y0 + iy, a.nf2) * a.nf1 +
                                             (size_t)wrap_index(z0 + iz, a.nf3) * a.nf1 * a.nf2; }
                        red_add(fwt + o, cr * wgt, ci * wgt);
                    }
                }
            }
        }
        __syncwarp();
    }
}

// =============================================================================
// Interpolation: warp per point, lanes over the stencil (lane = (row r, column ix)),
// coalesced row gathers from the fine grid (L1/L2: points are bin-ordered so neighbouring
// points reuse lines).  Each lane accumulates its column over the passes with the row
// weights only (2 FMA per cell), scales by its x-weight once, and the 32 partial sums of
// EIGHT points are reduced together by a transposing butterfly (9 shuffles per component
// per 8 points instead of 40).  Result scattered to c[idxnupts].
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
    using G = Geo<DIM, NS>;
    extern __shared__ __align__(16) unsigned char smem[];
    T *s_hc = reinterpret_cast<T *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wbase = smem + 18 * 16 * sizeof(T) + warp * (warp_scratch_bytes<T, DIM, NS>() + 2 * 32 * sizeof(int));
    T *s_cre = reinterpret_cast<T *>(wbase);          // unused by interp (kept for the common layout)
    T *s_cim = s_cre + 32;
    T *s_ker = s_cim + 32;
    int *s_x0 = reinterpret_cast<int *>(s_ker + 32 * G::KVP);
    int *s_y0 = s_x0 + 32, *s_z0 = s_y0 + 32;
    int *s_idx = reinterpret_cast<int *>(s_cre);      // idxnupts of the 32 points of the batch
    stage_horner<T, NS>(a, s_hc);

    const int r = lane / NS, ix = lane - r * NS;
    const bool active = lane < G::R * NS;
    const long long nbatch = ((long long)a.M + 31) / 32;
    const long long total = nbatch * a.nt;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    const size_t plane = (size_t)a.nf1 * a.nf2;

    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp; w < total; w += wstride) {
        const int t = (int)(w / nbatch);
        const long long pb = (w - (long long)t * nbatch) * 32;
        const int cnt = (int)min((long long)32, a.M - pb);
        C *cout = a.c + (size_t)t * a.M;
        const C *fwt = a.fw + (size_t)t * a.fwstride;
        if (lane < cnt) {
            const int p = (int)(pb + lane);
            int xs0, ys0 = 0, zs0 = 0;
            point_weights<T, DIM, NS>(a, p, s_ker + lane * G::KVP, s_hc, xs0, ys0, zs0);
            s_idx[lane] = a.idx[p];
            s_x0[lane] = clampi(xs0, -a.nf1, a.nf1);
            s_y0[lane] = clampi(ys0, -a.nf2, a.nf2);
            s_z0[lane] = clampi(zs0, -a.nf3, a.nf3);
        }
        __syncwarp();
        for (int q0 = 0; q0 < cnt; q0 += 8) {
            T accr[8], acci[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                accr[j] = 0; acci[j] = 0;
                const int q = q0 + j;
                if (active && q < cnt) {
                    const T *kq = s_ker + q * G::KVP;
                    const C *col = fwt + wrap_index(s_x0[q] + ix, a.nf1);
                    const int y0 = s_y0[q], z0 = s_z0[q];
                    T sr = 0, si = 0;
#pragma unroll(G::ITERS <= 12 ? 12 : 4)
                    for (int it = 0; it < G::ITERS; ++it) {
                        const int row = it * G::R + r;
                        if (row < G::ROWS) {
                            T wgt; size_t o;
                            if (DIM == 1) { wgt = 1; o = 0; }
                            else if (DIM == 2) { wgt = kq[NS + row]; o = (size_t)wrap_index(y0 + row, a.nf2) * a.nf1; }
                            else { const int iz = row / NS, iy = row - iz * NS;
                                   wgt = kq[NS + iy] * kq[2 * NS + iz];
                                   o = (size_t)wrap_index(y0 + iy, a.nf2) * a.nf1 + (size_t)wrap_index(z0 + iz, a.nf3) * plane; }
                            const C v = col[o];
                            sr = fma(v.x, wgt, sr);
                            si = fma(v.y, wgt, si);
                        }
                    }
                    const T k1 = kq[ix];
                    accr[j] = sr * k1; acci[j] = si * k1;
                }
            }
            const T tr = reduce8(accr, lane), ti = reduce8(acci, lane);
            const int q = q0 + ((lane >> 2) & 7);
            // lane bits (4,3,2) select the point in reduce8's order: 4*bit4 + 2*bit3 + bit2
            const int qsel = q0 + (((lane >> 4) & 1) << 2) + (((lane >> 3) & 1) << 1) + ((lane >> 2) & 1);
            (void)q;
            if ((lane & 3) == 0 && qsel < cnt) { C o; o.x = tr; o.y = ti; cout[s_idx[qsel]] = o; }
        }
        __syncwarp();
    }
}

// ---- host-side launch helpers (one instantiation per (T, DIM) translation unit) ----
template <typename T, int DIM> int launch_spread(Plan<T> &p, const typename Plan<T>::C *c, typename Plan<T>::C *fw, int nt);
template <typename T, int DIM> int launch_interp(Plan<T> &p, typename Plan<T>::C *c, const typename Plan<T>::C *fw, int nt);
template <typename T, int DIM> size_t sm_spread_smem_per_warp(int ns, int tile_cells);

}  // namespace cfb
