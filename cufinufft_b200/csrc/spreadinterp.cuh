// spreadinterp.cuh -- spreading (type-1 step 1) and interpolation (type-2 step 3)
// kernels of the B200-native cuFINUFFT hot path.
//
// Reference kernels replaced (all of src/{1,2,3}d/spreadinterp{1,2,3}d.cu):
//   Spread_{1,2,3}d_NUptsdriven[_Horner]  -> spread_gm_kernel   (gpu_method 1, GM / GM-sort)
//   Spread_{1,2,3}d_Subprob[_Horner]      -> spread_sm_kernel   (gpu_method 2, SM)
//   Interp_{1,2,3}d_NUptsdriven[_Horner]  -> interp_kernel      (gather engine)
//   Interp_{2,3}d_Subprob[_Horner]        -> interp_tile_kernel (tile engine, the default when sorted)
//
// Design (DESIGN.md "spread" / "interp").  The reference gives every point to one thread that
// does 2*ns^d scalar atomics (shared or global).  Here
//  * a WARP owns a batch of 32 consecutive sorted points.  Phase A is thread-per-point: one
//    16/32-byte record (prefetched two batches ahead), the d*ns kernel values (exp/sqrt or
//    Horner), the strength gather (prefetched one batch ahead), all parked in a small per-warp
//    scratch.  Phase B is lane-per-cell: the 32 lanes are laid over the point's stencil,
//    lane = (row r, column ix), ITERS passes cover all rows.
//  * setpts sorts points by (bin, sub-bin, stencil origin), so consecutive points usually share
//    their WHOLE stencil.  Such a RUN is accumulated in registers (ITERS complex accumulators per
//    lane, packed f32x2 FMAs in fp32) and touches memory once per run instead of once per point:
//      spread SM : run -> warp-private padded tile in shared memory with plain LDS/FADD/STS
//                  (lanes of one flush hit distinct cells, padded strides keep them on distinct
//                  banks: no shared atomics at all); tile -> fine grid once per subproblem with
//                  vector RED (red.global.add.v2.f32) over the touched box, cleared on the way.
//                  Batches whose runs are short (sparse regions) skip the run bookkeeping and
//                  apply their points to the tile one by one; the choice is per batch.
//      spread GM : run -> fine grid directly with vector RED
//      interp (gather engine): the stencil's grid values are loaded once per run (coalesced row
//                  segments) and reused from registers for every point of the run; the per-lane
//                  partial sums of 8 points are transposed through shared memory and reduced together
//  * Work items = subproblems (<= maxsubprobsize consecutive sorted points of one INTERNAL bin: the
//    reference's bin or a sub-bin of it, setpts.cu / spread.cu) x transforms, pulled from a global
//    work counter by persistent warps (spread) or blocks (tile interp).
//  * The tile interpolation engine (interp_tile_kernel, below) stages the bin's grid tile in shared
//    memory with cp.async and is thread-per-point.
// Stencils too large for register accumulators (3-D fp64, ns >= 9) use the same kernels with
// MERGE = false: every pass is applied to memory immediately (four passes' loads at a time).
// Every kernel is instantiated per evaluator (HORNER): the other evaluator's code is not emitted.
#pragma once
#include "cfb_device.cuh"

namespace cfb {

template <typename T>
struct SIArgs {
    using C = typename cplx_of<T>::type;
    const PtRec<T> *recs;       // sorted point records
    C *c;                       // strengths in (spread) / values out (interp), [nt][M]
    C *fw;                      // fine grids [nt][nf3][nf2][nf1]
    // work list of the tile engines over the INTERNAL bins (== the reference's bins unless setpts
    // split them): keyoff[ib * cpb] = first sorted point of internal bin ib, s2b / substart = the
    // subproblem -> bin map and offsets, *nsub = number of subproblems
    const int *keyoff, *s2b, *substart, *nsub;
    int *counter;               // global work counter (zeroed before launch)
    int cpb;                    // sort keys per internal bin
    int rbs1, rbs2, rbs3;       // reference bin size; bs1..3 below are the internal (tile) bin size
    int spb1, spb2, spbt;       // sub-bins per reference bin: x, y, total
    const T *hcoef;             // Horner coefficients [ncoef][NS] as T (device)
    int M, nt, maxsub;
    int nf1, nf2, nf3;
    int bs1, bs2, bs3, nb1, nb2;
    int pad, ex, ey, ez;        // tile halo and extents (cells)
    int sy, sz, tile_cells;     // padded tile strides
    int horner, ncoef;
    int zshift;                 // slab plans: local plane = global plane - zshift (0 otherwise)
    int zm, nb3, npairs;        // tile interpolation, 3-D: zm > 1 = one tile serves zm z-adjacent bins (npairs = nb1*nb2*ceil(nb3/zm) work items)
    int bankc;                  // > 0: the points of a bin are ordered by shared-memory bank class (cpb == bankc classes per bin)
    int thr_num, thr_den;       // tuning override of the short-run threshold (0 = built-in; env CFB_DIRECT_THR=num/den)
    T es_c, es_beta;
    long long fwstride;
};

constexpr int roundup(int v, int m) { return (v + m - 1) / m * m; }

// Stencil geometry of one (precision, dimension, width) instantiation.
// Scratch row of a point (T units), VEC layout: [kx: NSX][W: R x WS] with W[r][it] the row weight
// of pass `it` for lanes of row-slot r (2-D: ky[it*R+r]; 3-D: ky[iy]*kz[iz], row = it*R+r =
// iz*NS+iy; zero beyond ROWS), every segment 16-byte aligned so phase B reads it with LDS.128.
// SCALAR layout (3-D without precomputed products): [kx: NS][ky: NS][kz: NS].
template <typename T, int DIM, int NS> struct Geo {
    static constexpr int V = 16 / (int)sizeof(T);                         // reals per 16-byte vector
    static constexpr int R = 32 / NS;                                     // stencil rows per warp pass
    static constexpr int LANES = R * NS;                                  // active lanes
    static constexpr int ROWS = DIM == 1 ? 1 : (DIM == 2 ? NS : NS * NS);
    static constexpr int ITERS = (ROWS + R - 1) / R;
    static constexpr bool PROD = DIM == 3 && NS <= 7 && sizeof(T) == 4;   // ky*kz products precomputed per point
    static constexpr bool VEC = DIM == 2 || PROD;
    static constexpr int NSX = VEC ? roundup(NS, V) : NS;
    static constexpr int WS = (VEC && ITERS == 1) ? 1 : roundup(ITERS, V);   // single-pass stencils: one weight per row slot
    static constexpr int KP0 = DIM == 1 ? NS : (VEC ? NSX + roundup(R * WS, V) : 3 * NS);
    // VEC: KP/V odd (vector phase-A stores and phase-B loads conflict-free); else KP odd
    static constexpr int KP = VEC ? (((KP0 / V) | 1) * V) : (KP0 | 1);
    static constexpr int NACC = VEC ? WS : (DIM == 1 ? 1 : roundup(ITERS, 2));   // accumulator slots (padded passes have weight 0)
    static constexpr bool MERGE = NACC * (int)(sizeof(T) / 4) <= 48;      // run accumulators (4 NACC 32-bit registers in fp64) fit in registers
    static constexpr bool TOFF_REGS = ITERS <= (sizeof(T) == 4 ? 32 : 16);  // per-pass tile offsets kept in registers
    static constexpr int SM_MAXW = ITERS * (int)(sizeof(T) / 4) > 24 ? 4 : 16;   // warps per SM-spread block (register budget)
};

// host mirror of Geo::SM_MAXW for the plan-time tile geometry
inline int sm_spread_max_warps(int dim, int ns, int real_bytes)
{
    const int R = 32 / ns, rows = dim == 1 ? 1 : (dim == 2 ? ns : ns * ns), iters = (rows + R - 1) / R;
    return iters * (real_bytes / 4) > 24 ? 4 : 16;
}

// bytes of per-warp scratch: ker[32*KP] T | c[32] C | x0,y0,z0[32] int | red[8*33] C (interp only uses it)
template <typename T, int DIM, int NS, bool WITH_RED = true>
__host__ __device__ constexpr size_t warp_scratch_bytes()
{
    return (size_t)32 * Geo<T, DIM, NS>::KP * sizeof(T) + 32 * 2 * sizeof(T) + 3 * 32 * sizeof(int) +
           (WITH_RED ? 8 * 34 * 2 * sizeof(T) : 0);      // the spread kernels do not carry the reduction buffer
}

template <typename T, int DIM, int NS>
struct Scratch {
    using C = typename cplx_of<T>::type;
    T *ker; C *c; int *x0, *y0, *z0; C *red;
    __device__ __forceinline__ explicit Scratch(unsigned char *base)
    {
        using G = Geo<T, DIM, NS>;
        ker = reinterpret_cast<T *>(base);
        c = reinterpret_cast<C *>(ker + 32 * G::KP);
        x0 = reinterpret_cast<int *>(c + 32);
        y0 = x0 + 32;
        z0 = y0 + 32;
        red = reinterpret_cast<C *>(z0 + 32);
    }
};

template <typename T> struct vec16;
template <> struct vec16<float>  { using type = float4; };
template <> struct vec16<double> { using type = double2; };

// ---- phase A: thread-per-point kernel weights into the per-warp scratch ------
template <typename T, int DIM, int NS>
__device__ __forceinline__ void point_weights(const SIArgs<T> &a, const PtRec<T> &rec, T *kp, const T *s_hc,
                                              int &xstart, int &ystart, int &zstart)
{
    using G = Geo<T, DIM, NS>;
    constexpr bool F32 = sizeof(T) == 4;
    xstart = stencil_start(rec.x, NS);
    if (DIM > 1) ystart = stencil_start(rec.y, NS);
    if (DIM > 2) zstart = stencil_start(rec.z, NS);
    const T x1 = (T)xstart - rec.x;
    if constexpr (!G::VEC) {
        // 1-D, or 3-D with on-the-fly products: plain [kx][ky][kz]
        kernel_vector<T, NS, F32>(kp, x1, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
        if (DIM == 3) {
            kernel_vector<T, NS, F32>(kp + NS, (T)ystart - rec.y, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
            kernel_vector<T, NS, F32>(kp + 2 * NS, (T)zstart - rec.z, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
        }
    } else if constexpr (F32) {
        // everything in registers, written with 16-byte stores
        float kx[G::NSX], ky[NS], kz[NS];
        kernel_vector<T, NS, true>(kx, x1, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
#pragma unroll
        for (int i = NS; i < G::NSX; ++i) kx[i] = 0;
        kernel_vector<T, NS, true>(ky, (T)ystart - rec.y, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
        if (DIM == 3) kernel_vector<T, NS, true>(kz, (T)zstart - rec.z, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
        float4 *dst = reinterpret_cast<float4 *>(kp);
#pragma unroll
        for (int i = 0; i < G::NSX / 4; ++i) dst[i] = make_float4(kx[4 * i], kx[4 * i + 1], kx[4 * i + 2], kx[4 * i + 3]);
        auto wgt = [&](int r, int it) -> float {
            const int row = it * G::R + r;
            if (it >= G::ITERS || row >= G::ROWS) return 0.0f;
            if (DIM == 2) return ky[row];
            return ky[row % NS] * kz[row / NS];
        };
        if constexpr (G::WS == 1) {
            auto wr = [&](int r) -> float { return r < G::R ? wgt(r, 0) : 0.0f; };
#pragma unroll
            for (int r0 = 0; r0 < G::R; r0 += 4)
                dst[(G::NSX + r0) / 4] = make_float4(wr(r0), wr(r0 + 1), wr(r0 + 2), wr(r0 + 3));
        } else {
#pragma unroll
            for (int r = 0; r < G::R; ++r)
#pragma unroll
                for (int it = 0; it < G::WS; it += 4)
                    dst[(G::NSX + r * G::WS + it) / 4] = make_float4(wgt(r, it), wgt(r, it + 1), wgt(r, it + 2), wgt(r, it + 3));
        }
    } else {
        // fp64 2-D: evaluate straight into the slots (no unrolling: the evaluation is long)
        kernel_vector<T, NS, false>(kp, x1, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
#pragma unroll
        for (int i = NS; i < G::NSX; ++i) kp[i] = 0;
        const T y1 = (T)ystart - rec.y;
        if (!a.horner) {
#pragma unroll 1
            for (int slot = 0; slot < G::R * G::WS; ++slot) {
                const int r = slot / G::WS, it = slot - r * G::WS, row = it * G::R + r;
                T v = 0;
                if (it < G::ITERS && row < G::ROWS) {
                    T ax = y1 + (T)row;
                    ax = ax < 0 ? -ax : ax;
                    v = es_eval(ax, a.es_c, a.es_beta, (T)(NS * 0.5));
                }
                kp[G::NSX + slot] = v;
            }
        } else {
            T ky[NS];
            kernel_vector<T, NS, true>(ky, y1, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
#pragma unroll
            for (int r = 0; r < G::R; ++r)
#pragma unroll
                for (int it = 0; it < G::WS; ++it) {
                    const int row = it * G::R + r;
                    kp[G::NSX + r * G::WS + it] = (it < G::ITERS && row < G::ROWS) ? ky[row] : (T)0;
                }
        }
    }
    if (DIM > 2) zstart -= a.zshift;      // weights from the global coordinate, grid index slab-local
}

// row weights of all passes for this lane (row-slot r) into w[NACC]
template <typename T, int DIM, int NS>
__device__ __forceinline__ void load_row_weights(const T *kq, int r, T (&w)[Geo<T, DIM, NS>::NACC])
{
    using G = Geo<T, DIM, NS>;
    if constexpr (DIM == 1) {
        w[0] = (T)1;
    } else if constexpr (G::VEC && G::WS == 1) {
        w[0] = kq[G::NSX + r];
    } else if constexpr (G::VEC) {
        using V16 = typename vec16<T>::type;
        const V16 *src = reinterpret_cast<const V16 *>(kq + G::NSX + r * G::WS);
#pragma unroll
        for (int j = 0; j < G::WS / G::V; ++j) {
            const V16 v = src[j];
            if constexpr (sizeof(T) == 4) { w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w; }
            else { w[2 * j] = v.x; w[2 * j + 1] = v.y; }
        }
    } else {
#pragma unroll
        for (int it = 0; it < G::NACC; ++it) {
            const int row = it * G::R + r;
            const int iz = row / NS, iy = row - iz * NS;
            w[it] = (it < G::ITERS && row < G::ROWS) ? kq[NS + iy] * kq[2 * NS + iz] : (T)0;
        }
    }
}

// acc[it] += (cr, ci) * w[it] for all passes.  fp32: packed f32x2 FMAs over PAIRS OF PASSES
// (accumulators are stored as (x_it, x_it+1), (y_it, y_it+1)); fp64: scalar DFMA.
template <typename T, int N> struct RunAcc;
template <int N> struct RunAcc<float, N> {
    static_assert(N % 2 == 0 || N == 1, "padded pass count");
    static constexpr int NP = (N + 1) / 2;
    unsigned long long ax[NP], ay[NP];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int j = 0; j < NP; ++j) { ax[j] = 0ull; ay[j] = 0ull; }
    }
    __device__ __forceinline__ static unsigned long long pack(float lo, float hi) {
        unsigned long long v;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
        return v;
    }
    __device__ __forceinline__ void fma(float cr, float ci, const float (&w)[N]) {
        const unsigned long long cr2 = pack(cr, cr), ci2 = pack(ci, ci);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const unsigned long long w2 = pack(w[2 * j], N > 1 ? w[2 * j + (N > 1 ? 1 : 0)] : 0.0f);
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ax[j]) : "l"(cr2), "l"(w2));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ay[j]) : "l"(ci2), "l"(w2));
        }
    }
    __device__ __forceinline__ float2 get(int it) const {
        float xl, xh, yl, yh;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(xl), "=f"(xh) : "l"(ax[it >> 1]));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(yl), "=f"(yh) : "l"(ay[it >> 1]));
        return (it & 1) ? make_float2(xh, yh) : make_float2(xl, yl);
    }
};
template <> struct RunAcc<float, 1> {          // single-pass stencils: plain scalar FMAs
    float ax, ay;
    __device__ __forceinline__ void zero() { ax = 0.0f; ay = 0.0f; }
    __device__ __forceinline__ void fma(float cr, float ci, const float (&w)[1]) { ax = fmaf(cr, w[0], ax); ay = fmaf(ci, w[0], ay); }
    __device__ __forceinline__ float2 get(int) const { return make_float2(ax, ay); }
};
template <int N> struct RunAcc<double, N> {
    double ax[N], ay[N];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int j = 0; j < N; ++j) { ax[j] = 0; ay[j] = 0; }
    }
    __device__ __forceinline__ void fma(double cr, double ci, const double (&w)[N]) {
#pragma unroll
        for (int j = 0; j < N; ++j) { ax[j] = ::fma(cr, w[j], ax[j]); ay[j] = ::fma(ci, w[j], ay[j]); }
    }
    __device__ __forceinline__ double2 get(int it) const { return make_double2(ax[it], ay[it]); }
};

// (The Horner coefficients used to be staged in shared memory here; kernel_vector now reads them from constant
// memory.  The 18*16*sizeof(T) bytes at the head of every kernel's dynamic shared memory are unused.)
template <typename T, int NS>
__device__ __forceinline__ const T *stage_horner(const SIArgs<T> &, T *s_hc) { __syncthreads(); return s_hc; }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <typename T>
__device__ __forceinline__ PtRec<T> null_rec() { PtRec<T> r; r.x = 0; r.y = 0; r.z = 0; r.idx = 0; return r; }

// Work item s (one subproblem: <= maxsub consecutive sorted points of one internal bin) ->
// its point range and the origin of its tile (bin origin minus the halo).
template <typename T, int DIM>
__device__ __forceinline__ void decode_subproblem(const SIArgs<T> &a, int s, int &pstart, int &n, int &ox, int &oy, int &oz)
{
    const int bin = a.s2b[s];
    const int k = s - a.substart[bin];
    const int p0 = a.keyoff[(size_t)bin * a.cpb], p1 = a.keyoff[(size_t)(bin + 1) * a.cpb];
    pstart = p0 + k * a.maxsub;
    n = min(a.maxsub, p1 - pstart);
    const int rb = bin / a.spbt, sub = bin - rb * a.spbt;
    const int b1 = rb % a.nb1, b23 = rb / a.nb1;
    const int s1 = sub % a.spb1, s23 = sub / a.spb1;
    ox = b1 * a.rbs1 + s1 * a.bs1 - a.pad;
    oy = 0; oz = 0;
    if (DIM > 1) { const int b2 = b23 % a.nb2, s2 = s23 % a.spb2; oy = b2 * a.rbs2 + s2 * a.bs2 - a.pad; }
    if (DIM > 2) { const int b3 = b23 / a.nb2, s3 = s23 / a.spb2; oz = b3 * a.rbs3 + s3 * a.bs3 - a.pad; }
}

// =============================================================================
// SM spread: warp-private tile, run accumulation in registers, lane-per-cell flushes.
// dynamic smem: [hcoef 18*16 T][per warp: tile C[tile_cells] | scratch]
// =============================================================================
template <typename T, int DIM, int NS, bool HORNER>
__global__ void __launch_bounds__(32 * Geo<T, DIM, NS>::SM_MAXW)
spread_sm_kernel(const SIArgs<T> a_in)
{
    SIArgs<T> a = a_in;
    a.horner = HORNER ? 1 : 0;          // compile-time constant: the other evaluator's code is not emitted
    using C = typename cplx_of<T>::type;
    using G = Geo<T, DIM, NS>;
    extern __shared__ __align__(16) unsigned char smem[];
    T *s_hc = reinterpret_cast<T *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = (size_t)a.tile_cells * sizeof(C) + warp_scratch_bytes<T, DIM, NS, false>();
    unsigned char *wbase = smem + 18 * 16 * sizeof(T) + warp * per_warp;
    C *tile = reinterpret_cast<C *>(wbase);
    Scratch<T, DIM, NS> sc(wbase + (size_t)a.tile_cells * sizeof(C));
    int *s_off = sc.x0;
    stage_horner<T, NS>(a, s_hc);

    const bool active = lane < G::LANES;
    const int r = active ? lane / NS : 0, ix = active ? lane - r * NS : 0;
    const int nsub = *a.nsub;
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

    // the tile is clean whenever a work item starts: cleared here once, then by every flush
    for (int i = lane; i < a.tile_cells; i += 32) { tile[i].x = 0; tile[i].y = 0; }

    for (;;) {
        long long w = 0;
        if (lane == 0) w = atomicAdd(a.counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total) break;
        const int t = (int)(w / nsub), s = (int)(w - (long long)t * nsub);
        int pstart, n, ox, oy, oz;
        decode_subproblem<T, DIM>(a, s, pstart, n, ox, oy, oz);
        const C *cin = a.c + (size_t)t * a.M;
        C *fwt = a.fw + (size_t)t * a.fwstride;
        const PtRec<T> *recs = a.recs + pstart;

        RunAcc<T, G::MERGE ? G::NACC : 1> acc;
        acc.zero();
        int cur = -1;                                   // tile offset of the open run
        int ylo = 1 << 30, yhi = -1, zlo = 1 << 30, zhi = -1;   // touched stencil origins (tile rows), per lane

        // add the open run's accumulators into the tile (lanes touch distinct cells)
        // (the cells of one flush are distinct: all loads are issued before the first store so that
        // their latencies overlap instead of forming a load -> add -> store chain per pass)
        auto flush_run = [&]() {
            if constexpr (G::MERGE) {
                __syncwarp();
                C *cell0 = tile + cur + ix;
#pragma unroll
                for (int c0 = 0; c0 < G::ITERS; c0 += 8) {       // eight passes' loads in flight at a time
                    C v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c0 + j < G::ITERS && active && (c0 + j) * G::R + r < G::ROWS) v[j] = cell0[tile_off(c0 + j)];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c0 + j < G::ITERS && active && (c0 + j) * G::R + r < G::ROWS) {
                            const C d = acc.get(c0 + j);
                            v[j].x += d.x; v[j].y += d.y;
                        }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c0 + j < G::ITERS && active && (c0 + j) * G::R + r < G::ROWS) cell0[tile_off(c0 + j)] = v[j];
                }
                acc.zero();
            }
        };
        // one point applied to the tile at once (no run bookkeeping): loads of all passes first.
        // (Fetching the operands one point ahead of the tile update was measured and did not pay:
        // the kernel is issue-bound at this point, profiles/r01p: 62 % of the issue slots.)
        auto fetch_point = [&](int q, T (&wq)[G::NACC], T &cr, T &ci, int &off) {
            if constexpr (G::MERGE) {
                const T *kq = sc.ker + q * G::KP;
                load_row_weights<T, DIM, NS>(kq, r, wq);
                const T k1 = active ? kq[ix] : (T)0;
                const C cv = sc.c[q];
                cr = cv.x * k1; ci = cv.y * k1;
                off = s_off[q];
            }
        };
        auto apply_point = [&](const T (&wq)[G::NACC], T cr, T ci, int off) {
            if constexpr (G::MERGE) {
                C *cell0 = tile + off + ix;
#pragma unroll
                for (int c0 = 0; c0 < G::ITERS; c0 += 8) {
                    C v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c0 + j < G::ITERS && active && (c0 + j) * G::R + r < G::ROWS) v[j] = cell0[tile_off(c0 + j)];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c0 + j < G::ITERS && active && (c0 + j) * G::R + r < G::ROWS) {
                            v[j].x = fma(cr, wq[c0 + j], v[j].x); v[j].y = fma(ci, wq[c0 + j], v[j].y);
                        }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c0 + j < G::ITERS && active && (c0 + j) * G::R + r < G::ROWS) cell0[tile_off(c0 + j)] = v[j];
                }
                __syncwarp();      // the next point's loads (other lanes, possibly the same cells) come after these stores
            }
        };

        // software pipeline: records two batches ahead, strengths one batch ahead
        PtRec<T> rec_cur = lane < n ? load_rec(recs + lane) : null_rec<T>();
        PtRec<T> rec_nxt = 32 + lane < n ? load_rec(recs + 32 + lane) : null_rec<T>();
        C c_cur = lane < n ? cin[rec_index(rec_cur)] : C{0, 0};

        for (int base = 0; base < n; base += 32) {
            const int cnt = min(32, n - base);
            C c_nxt = base + 32 + lane < n ? cin[rec_index(rec_nxt)] : C{0, 0};
            PtRec<T> rec_nn = base + 64 + lane < n ? load_rec(recs + base + 64 + lane) : null_rec<T>();
            __syncwarp();
            int myoff = -2;
            if (lane < cnt) {
                int xs0, ys0 = 0, zs0 = 0;
                point_weights<T, DIM, NS>(a, rec_cur, sc.ker + lane * G::KP, s_hc, xs0, ys0, zs0);
                sc.c[lane] = c_cur;
                myoff = clampi(xs0 - ox, 0, a.ex - NS);
                if (DIM > 1) { const int yo = clampi(ys0 - oy, 0, a.ey - NS); myoff += yo * a.sy; ylo = min(ylo, yo); yhi = max(yhi, yo); }
                if (DIM > 2) { const int zo = clampi(zs0 - oz, 0, a.ez - NS); myoff += zo * a.sz; zlo = min(zlo, zo); zhi = max(zhi, zo); }
                s_off[lane] = myoff;
            }
            __syncwarp();
            if constexpr (G::MERGE) {
                int prev = __shfl_up_sync(0xffffffffu, myoff, 1);
                if (lane == 0) prev = cur;
                const unsigned starts = __ballot_sync(0xffffffffu, lane < cnt && myoff != prev);
                // Short runs (sparse regions): the run bookkeeping + flush costs more than it saves, the
                // points of this batch go to the tile one by one.  Break-even mean run length, measured
                // (tools/gpu_thr.sh, profiles/r01w): 3 points for multi-pass stencils, 8 for single-pass ones.
                constexpr int THR_NUM = G::ITERS > 1 ? 3 : 8, THR_DEN = 1;
                const int thr_num = a.thr_num > 0 ? a.thr_num : THR_NUM, thr_den = a.thr_num > 0 ? a.thr_den : THR_DEN;
                if (__popc(starts) * thr_num > cnt * thr_den) {
                    if (cur >= 0) { flush_run(); cur = -1; }
                    __syncwarp();
#pragma unroll 2
                    for (int q = 0; q < cnt; ++q) {
                        T wq[G::NACC], cr, ci;
                        int off;
                        fetch_point(q, wq, cr, ci, off);
                        apply_point(wq, cr, ci, off);
                    }
                    __syncwarp();
                } else {
                // run by run: the accumulators live in registers across the tight inner loop (no
                // branch, no register renaming inside it); a run ends at the next start bit or with the batch
                int q = 0;
                while (q < cnt) {
                    if ((starts >> q) & 1u) {
                        if (cur >= 0) flush_run();
                        cur = s_off[q];
                    }
                    const unsigned rest = q < 31 ? (starts >> (q + 1)) : 0u;
                    const int qe = rest ? q + __ffs(rest) : cnt;
#pragma unroll 2
                    for (; q < qe; ++q) {
                        const T *kq = sc.ker + q * G::KP;
                        T wq[G::NACC];
                        load_row_weights<T, DIM, NS>(kq, r, wq);
                        const T k1 = active ? kq[ix] : (T)0;
                        const C cv = sc.c[q];
                        acc.fma(cv.x * k1, cv.y * k1, wq);
                    }
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
                    // four passes at a time: the loads of a chunk are issued before its first store
                    for (int c0 = 0; c0 < G::ITERS; c0 += 4) {
                        C v[4];
                        T wg[4];
                        C *cp[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int it = c0 + j, row = it * G::R + r;
                            const bool on = it < G::ITERS && active && row < G::ROWS;
                            const int iz = row / NS, iy = row - iz * NS;
                            wg[j] = !on ? (T)0 : (DIM == 1 ? (T)1 : (DIM == 2 ? kq[G::NSX + r * G::WS + it] : kq[NS + iy] * kq[2 * NS + iz]));
                            cp[j] = cell0 + (on ? tile_off(it) : 0);
                            v[j] = on ? *cp[j] : C{0, 0};
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int it = c0 + j;
                            if (it < G::ITERS && active && it * G::R + r < G::ROWS) {
                                v[j].x = fma(cr, wg[j], v[j].x); v[j].y = fma(ci, wg[j], v[j].y);
                                *cp[j] = v[j];
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            rec_cur = rec_nxt; rec_nxt = rec_nn; c_cur = c_nxt;
        }
        if (cur >= 0) flush_run();
        __syncwarp();

        // tile -> fine grid: vector RED over the touched rows only, clearing the tile on the way;
        // single periodic wrap (reference guard ix < nf+pad, src/2d/spreadinterp2d.cu:222-224, is
        // implied: cells past it are never touched and stay zero)
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o)); yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
                zlo = min(zlo, __shfl_xor_sync(0xffffffffu, zlo, o)); zhi = max(zhi, __shfl_xor_sync(0xffffffffu, zhi, o));
            }
            const int y0 = DIM > 1 ? ylo : 0, ny = DIM > 1 ? yhi - ylo + NS : 1;
            const int z0 = DIM > 2 ? zlo : 0, nz = DIM > 2 ? zhi - zlo + NS : 1;
            // flat index over the touched box [nz][ny][ex]: all lanes busy whatever ex is, four
            // independent tile loads in flight; (row, x) and (z, y) by exact float reciprocals
            const int ncell = yhi < 0 && DIM > 1 ? 0 : a.ex * ny * nz;
            const float inv_ex = 1.0f / (float)a.ex, inv_ny = 1.0f / (float)ny;
            for (int base = 0; base < ncell; base += 128) {
                C v[4];
                C *tp[4];
                int lx[4], ly[4], lz[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = base + j * 32 + lane;
                    int row = (int)(((float)idx + 0.5f) * inv_ex);
                    lx[j] = idx - row * a.ex;
                    lz[j] = DIM > 2 ? (int)(((float)row + 0.5f) * inv_ny) : 0;
                    ly[j] = row - lz[j] * ny;
                    tp[j] = tile + (z0 + lz[j]) * a.sz + (y0 + ly[j]) * a.sy + lx[j];
                    v[j] = idx < ncell ? *tp[j] : C{0, 0};
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (v[j].x != 0 || v[j].y != 0) {
                        size_t g = (size_t)wrap_index(ox + lx[j], a.nf1);
                        if (DIM > 1) g += (size_t)wrap_index(oy + y0 + ly[j], a.nf2) * a.nf1;
                        if (DIM > 2) g += (size_t)wrap_index(oz + z0 + lz[j], a.nf3) * a.nf1 * a.nf2;
                        red_add(fwt + g, v[j].x, v[j].y);
                        *tp[j] = C{0, 0};
                    }
                }
            }
        }
        __syncwarp();
    }
}

// =============================================================================
// GM / GM-sort spread: same lane-per-cell mapping and run accumulation, runs go straight
// into the fine grid with vector RED (no tile).  Work unit = 256 consecutive points.
// =============================================================================
template <typename T, int DIM, int NS, bool HORNER>
__global__ void __launch_bounds__(256)
spread_gm_kernel(const SIArgs<T> a_in)
{
    SIArgs<T> a = a_in;
    a.horner = HORNER ? 1 : 0;          // compile-time constant: the other evaluator's code is not emitted
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
        const PtRec<T> *recs = a.recs + p0;

        RunAcc<T, G::MERGE ? G::NACC : 1> acc;
        acc.zero();
        int cx = 0, cy = 0, cz = 0;
        bool open = false;

        auto flush_run = [&]() {
            if constexpr (G::MERGE) {
                const int gx = wrap_index(cx + ix, a.nf1);
#pragma unroll
                for (int it = 0; it < G::ITERS; ++it) {
                    const int row = it * G::R + r;
                    const C d = acc.get(it);
                    if (active && row < G::ROWS && (d.x != 0 || d.y != 0)) {
                        size_t o = gx;
                        if (DIM == 2) o += (size_t)wrap_index(cy + row, a.nf2) * a.nf1;
                        if (DIM == 3) { const int iz = row / NS, iy = row - iz * NS;
                                        o += (size_t)wrap_index(cy + iy, a.nf2) * a.nf1 + (size_t)wrap_index(cz + iz, a.nf3) * plane; }
                        red_add(fwt + o, d.x, d.y);
                    }
                }
                acc.zero();
            }
        };

        PtRec<T> rec_cur = lane < n ? load_rec(recs + lane) : null_rec<T>();
        PtRec<T> rec_nxt = 32 + lane < n ? load_rec(recs + 32 + lane) : null_rec<T>();
        C c_cur = lane < n ? cin[rec_index(rec_cur)] : C{0, 0};

        for (int base = 0; base < n; base += 32) {
            const int cnt = min(32, n - base);
            C c_nxt = base + 32 + lane < n ? cin[rec_index(rec_nxt)] : C{0, 0};
            PtRec<T> rec_nn = base + 64 + lane < n ? load_rec(recs + base + 64 + lane) : null_rec<T>();
            __syncwarp();
            int mx = 0, my = 0, mz = 0;
            if (lane < cnt) {
                point_weights<T, DIM, NS>(a, rec_cur, sc.ker + lane * G::KP, s_hc, mx, my, mz);
                sc.c[lane] = c_cur;
                mx = clampi(mx, -a.nf1, a.nf1); my = clampi(my, -a.nf2, a.nf2); mz = clampi(mz, -a.nf3, a.nf3);
                sc.x0[lane] = mx; sc.y0[lane] = my; sc.z0[lane] = mz;
            }
            __syncwarp();
            int px = __shfl_up_sync(0xffffffffu, mx, 1), py = __shfl_up_sync(0xffffffffu, my, 1), pz = __shfl_up_sync(0xffffffffu, mz, 1);
            if (lane == 0) { px = cx; py = cy; pz = cz; }
            const bool differs = mx != px || my != py || mz != pz || (lane == 0 && !open);
            const unsigned starts = G::MERGE ? __ballot_sync(0xffffffffu, lane < cnt && differs) : 0xffffffffu;
            if constexpr (G::MERGE) {
                int q = 0;
                while (q < cnt) {
                    if ((starts >> q) & 1u) {
                        if (open) flush_run();
                        cx = sc.x0[q]; cy = sc.y0[q]; cz = sc.z0[q];
                        open = true;
                    }
                    const unsigned rest = q < 31 ? (starts >> (q + 1)) : 0u;
                    const int qe = rest ? q + __ffs(rest) : cnt;
#pragma unroll 2
                    for (; q < qe; ++q) {
                        const T *kq = sc.ker + q * G::KP;
                        T wq[G::NACC];
                        load_row_weights<T, DIM, NS>(kq, r, wq);
                        const T k1 = active ? kq[ix] : (T)0;
                        const C cv = sc.c[q];
                        acc.fma(cv.x * k1, cv.y * k1, wq);
                    }
                }
            } else {
                for (int q = 0; q < cnt; ++q) {
                    const T *kq = sc.ker + q * G::KP;
                    const T k1 = active ? kq[ix] : (T)0;
                    const C cv = sc.c[q];
                    const T cr = cv.x * k1, ci = cv.y * k1;
                    {
                    const int gx = wrap_index(sc.x0[q] + ix, a.nf1);
                    const int y0 = sc.y0[q], z0 = sc.z0[q];
#pragma unroll 4
                    for (int it = 0; it < G::ITERS; ++it) {
                        const int row = it * G::R + r;
                        if (active && row < G::ROWS) {
                            const int iz = row / NS, iy = row - iz * NS;
                            const T wgt = DIM == 1 ? (T)1 : (DIM == 2 ? kq[G::NSX + r * G::WS + it] : kq[NS + iy] * kq[2 * NS + iz]);
                            size_t o = gx;
                            if (DIM == 2) o += (size_t)wrap_index(y0 + row, a.nf2) * a.nf1;
                            if (DIM == 3) o += (size_t)wrap_index(y0 + iy, a.nf2) * a.nf1 + (size_t)wrap_index(z0 + iz, a.nf3) * plane;
                            red_add(fwt + o, cr * wgt, ci * wgt);
                        }
                    }
                    }
                }
            }
            rec_cur = rec_nxt; rec_nxt = rec_nn; c_cur = c_nxt;
        }
        if (open) flush_run();
        __syncwarp();
    }
}

// =============================================================================
// Interpolation: lanes over the stencil (lane = (row r, column ix)); the grid values of a run's
// stencil are loaded once (coalesced row segments; points are sorted so neighbouring runs reuse
// L1/L2 lines) and kept in registers for all points of the run.  Each lane accumulates its
// column over the passes with the row weights (2 FMA per cell) and scales by its x-weight once;
// the per-lane partial sums of 8 points are parked in shared memory [point][lane] and summed by
// lane = (point, quarter) with two shuffle steps.  Result scattered to c[index].
// =============================================================================
template <typename T, int DIM, int NS, bool HORNER>
__global__ void __launch_bounds__(256)
interp_kernel(const SIArgs<T> a_in)
{
    SIArgs<T> a = a_in;
    a.horner = HORNER ? 1 : 0;          // compile-time constant: the other evaluator's code is not emitted
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
    constexpr int RS = 34;                             // row stride of the reduction buffer (cells)
    const int rp = lane >> 2, rq = lane & 3;           // reduction role: point rp of the group, quarter rq

    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp; w < total; w += wstride) {
        const int t = (int)(w / nchunk);
        const long long p0 = (w - (long long)t * nchunk) * (32 * CH);
        const int n = (int)min((long long)(32 * CH), a.M - p0);
        C *cout = a.c + (size_t)t * a.M;
        const C *fwt = a.fw + (size_t)t * a.fwstride;
        const PtRec<T> *recs = a.recs + p0;

        C v[G::MERGE ? G::NACC : 1];
        int cx = 0, cy = 0, cz = 0;
        bool open = false;

        auto load_run = [&]() {
            if constexpr (G::MERGE) {
                const C *col = fwt + wrap_index(cx + ix, a.nf1);
#pragma unroll
                for (int it = 0; it < G::NACC; ++it) {
                    const int row = it * G::R + r;
                    v[it].x = 0; v[it].y = 0;
                    if (it < G::ITERS && active && row < G::ROWS) {
                        size_t o = 0;
                        if (DIM == 2) o = (size_t)wrap_index(cy + row, a.nf2) * a.nf1;
                        if (DIM == 3) { const int iz = row / NS, iy = row - iz * NS;
                                        o = (size_t)wrap_index(cy + iy, a.nf2) * a.nf1 + (size_t)wrap_index(cz + iz, a.nf3) * plane; }
                        v[it] = col[o];
                    }
                }
            }
        };

        PtRec<T> rec_cur = lane < n ? load_rec(recs + lane) : null_rec<T>();
        PtRec<T> rec_nxt = 32 + lane < n ? load_rec(recs + 32 + lane) : null_rec<T>();

        for (int base = 0; base < n; base += 32) {
            const int cnt = min(32, n - base);
            PtRec<T> rec_nn = base + 64 + lane < n ? load_rec(recs + base + 64 + lane) : null_rec<T>();
            __syncwarp();
            int mx = 0, my = 0, mz = 0;
            if (lane < cnt) {
                point_weights<T, DIM, NS>(a, rec_cur, sc.ker + lane * G::KP, s_hc, mx, my, mz);
                s_idx[lane] = rec_index(rec_cur);
                mx = clampi(mx, -a.nf1, a.nf1); my = clampi(my, -a.nf2, a.nf2); mz = clampi(mz, -a.nf3, a.nf3);
                sc.x0[lane] = mx; sc.y0[lane] = my; sc.z0[lane] = mz;
            }
            __syncwarp();
            int px = __shfl_up_sync(0xffffffffu, mx, 1), py = __shfl_up_sync(0xffffffffu, my, 1), pz = __shfl_up_sync(0xffffffffu, mz, 1);
            if (lane == 0) { px = cx; py = cy; pz = cz; }
            const bool differs = mx != px || my != py || mz != pz || (lane == 0 && !open);
            const unsigned starts = G::MERGE ? __ballot_sync(0xffffffffu, lane < cnt && differs) : 0xffffffffu;

            for (int q0 = 0; q0 < cnt; q0 += 8) {
                const int gcnt = min(8, cnt - q0);
#pragma unroll 1
                for (int j = 0; j < gcnt; ++j) {
                    const int q = q0 + j;
                    const T *kq = sc.ker + q * G::KP;
                    T sr = 0, si = 0;
                    if constexpr (G::MERGE) {
                        if ((starts >> q) & 1u) {
                            cx = sc.x0[q]; cy = sc.y0[q]; cz = sc.z0[q];
                            open = true;
                            load_run();
                        }
                        T wq[G::NACC];
                        load_row_weights<T, DIM, NS>(kq, r, wq);
#pragma unroll
                        for (int it = 0; it < G::NACC; ++it) {
                            sr = fma(v[it].x, wq[it], sr);
                            si = fma(v[it].y, wq[it], si);
                        }
                    } else {
                        const C *col = fwt + wrap_index(sc.x0[q] + ix, a.nf1);
                        const int y0 = sc.y0[q], z0 = sc.z0[q];
#pragma unroll 4
                        for (int it = 0; it < G::ITERS; ++it) {
                            const int row = it * G::R + r;
                            if (active && row < G::ROWS) {
                                const int iz = row / NS, iy = row - iz * NS;
                                const T wgt = DIM == 1 ? (T)1 : (DIM == 2 ? kq[G::NSX + r * G::WS + it] : kq[NS + iy] * kq[2 * NS + iz]);
                                size_t o = 0;
                                if (DIM == 2) o = (size_t)wrap_index(y0 + row, a.nf2) * a.nf1;
                                if (DIM == 3) o = (size_t)wrap_index(y0 + iy, a.nf2) * a.nf1 + (size_t)wrap_index(z0 + iz, a.nf3) * plane;
                                const C g = col[o];
                                sr = fma(g.x, wgt, sr);
                                si = fma(g.y, wgt, si);
                            }
                        }
                    }
                    const T k1 = active ? kq[ix] : (T)0;
                    sc.red[j * RS + lane] = C{sr * k1, si * k1};
                }
                __syncwarp();
                // lane (rp, rq) sums lanes [8 rq, 8 rq + 8) of point rp, then two butterfly steps
                T tr = 0, ti = 0;
                if (rp < gcnt) {
                    const C *src = sc.red + rp * RS + rq * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const C g = src[i]; tr += g.x; ti += g.y; }
                }
                tr += __shfl_xor_sync(0xffffffffu, tr, 1); ti += __shfl_xor_sync(0xffffffffu, ti, 1);
                tr += __shfl_xor_sync(0xffffffffu, tr, 2); ti += __shfl_xor_sync(0xffffffffu, ti, 2);
                if (rq == 0 && rp < gcnt) cout[s_idx[q0 + rp]] = C{tr, ti};
                __syncwarp();
            }
            rec_cur = rec_nxt; rec_nxt = rec_nn;
        }
        __syncwarp();
    }
}

// =============================================================================
// Tile interpolation (replaces Interp_{2,3}d_Subprob[_Horner], src/2d/spreadinterp2d.cu:597-745,
// src/3d/spreadinterp3d.cu:760-948, and serves the sorted NUptsdriven requests as well).
// Work item = one subproblem (bin, <= maxsubprobsize sorted points) of one transform, pulled from
// a global counter by persistent BLOCKS.  The bin's fine-grid tile with its ceil(ns/2) halo is
// copied global -> shared with cp.async (LDGSTS, no register staging, periodic wrap resolved per
// row) while every thread already evaluates the kernel weights of its first point.  Then
// THREAD-PER-POINT: the d*ns weights stay in registers, the ns^d stencil is read from the tile
// with one LDS per complex cell (points are sorted by stencil origin inside the bin, so the lanes
// of a warp read the same or neighbouring cells: broadcast / conflict-free) and accumulated
// separably (2 FMA per cell + 2 per row + 2 per plane).  No atomics, no shuffles, no scratch.
// Tile layout [ez][ey][ex] complex, x fastest, unpadded.
// =============================================================================
__device__ __forceinline__ void cp_async_cell(float2 *dst, const float2 *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_cell(double2 *dst, const double2 *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename T, int DIM, int NS> struct TileInterp {
    static constexpr bool ROLL_Z = DIM == 3 && NS >= 8;      // plane loop rolled (kz indexed dynamically)
};

// 1-D / 2-D: 256 threads, registers capped for three blocks per SM (the fp64 kernels are bound by the
// FP64 pipe and the shared-memory read path: more resident warps keep both busier); 3-D: up to 512.
template <typename T, int DIM, int NS, bool HORNER>
__global__ void __launch_bounds__(DIM == 3 ? 512 : 256, DIM == 3 ? 1 : 3)
interp_tile_kernel(const SIArgs<T> a_in)
{
    SIArgs<T> a = a_in;
    a.horner = HORNER ? 1 : 0;          // compile-time constant: the other evaluator's code is not emitted
    using C = typename cplx_of<T>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    T *s_hc = reinterpret_cast<T *>(smem);
    C *tile = reinterpret_cast<C *>(smem + 18 * 16 * sizeof(T));
    __shared__ long long s_work;
    __shared__ int s_cs[4][16], s_ce[4][16];              // bank-class order: the range of every class in every member bin of the item
    stage_horner<T, NS>(a, s_hc);

    const int nsub = *a.nsub;
    // 3-D, wide fp64 stencils (one 130 KB tile per SM): a tile that serves ZM z-adjacent bins (bins are 2 planes
    // thick, the halo 10) is loaded once for ZM x as many points -- config 5: 26 x 26 x 14 cells for two bins instead
    // of 2 x (26 x 26 x 12), -16 % kernel time with the bin size doubled by hand (profiles/r03a).  The work item is
    // then the bin GROUP; its member bins and their subproblems are walked inside the block.
    const bool merged = DIM == 3 && a.zm > 1;
    const long long total = merged ? (long long)a.npairs * a.nt : (long long)nsub * a.nt;
    const int ex = a.ex, ey = a.ey, ez = a.ez;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const size_t plane = (size_t)a.nf1 * a.nf2;
    // Bank-class order (setpts.cu; type-2 plans): the points of a bin are grouped by the shared-memory bank
    // class of their first stencil cell, NC = 128 / sizeof(C) classes.  Thread t takes slot (class t % NC,
    // position t / NC): the lanes that share a shared-memory wavefront (8 for 16-byte cells, 16 for 8-byte
    // cells) then read NC different banks in EVERY stencil load -- no bank conflicts, where sorted-by-cell
    // neighbours collide whenever their cells are a multiple of NC apart (32 % of the wavefronts of the
    // config-5 kernel, profiles/r02f).  Classes are not equally full: slots beyond a class's last point are
    // handed the surplus points of fuller classes, so an item still takes ceil(n / blockDim) rounds.
    // (A merged tile shifts every cell of a member bin by the same multiple of ex*ey: the classes stay distinct.)
    const int NC = a.bankc;
    const int cls = NC > 0 ? (int)(threadIdx.x % NC) : 0, cap = NC > 0 ? (int)(blockDim.x / NC) : 0;

    for (;;) {
        __syncthreads();                                  // everybody is done with the previous tile
        if (threadIdx.x == 0) s_work = atomicAdd(a.counter, 1);
        __syncthreads();
        const long long w = s_work;
        if (w >= total) break;
        int t, ox, oy, oz, ezl = ez;                      // transform, tile origin, z extent of the loaded tile
        int first_bin = 0, nmember = 1, s_single = 0;
        if (!merged) {
            t = (int)(w / nsub);
            s_single = (int)(w - (long long)t * nsub);
            int pstart_, n_;
            decode_subproblem<T, DIM>(a, s_single, pstart_, n_, ox, oy, oz);
        } else {
            t = (int)(w / a.npairs);
            const int pp = (int)(w - (long long)t * a.npairs);
            const int b1 = pp % a.nb1, b23 = pp / a.nb1, b2 = b23 % a.nb2, bzp = b23 / a.nb2;
            nmember = min(a.zm, a.nb3 - bzp * a.zm);
            first_bin = b1 + a.nb1 * (b2 + a.nb2 * (bzp * a.zm));
            ox = b1 * a.rbs1 - a.pad; oy = b2 * a.rbs2 - a.pad; oz = bzp * a.zm * a.rbs3 - a.pad;
            ezl = nmember * a.rbs3 + 2 * a.pad;
            int npts = 0;                                 // nothing to do for a group without points
            for (int m = 0; m < nmember; ++m) {
                const size_t b = (size_t)first_bin + (size_t)m * a.nb1 * a.nb2;
                npts += a.keyoff[(b + 1) * a.cpb] - a.keyoff[b * a.cpb];
            }
            if (npts == 0) continue;
        }
        C *cout = a.c + (size_t)t * a.M;
        const C *fwt = a.fw + (size_t)t * a.fwstride;
        const int rows = ey * ezl;

        // ---- tile <- fine grid (async); rows are distributed over the warps
        for (int row = warp; row < rows; row += nwarps) {
            const int lz = DIM > 2 ? row / ey : 0, ly = row - lz * ey;
            int gy = 0, gz = 0;
            bool ok = true;
            if (DIM > 1) { gy = wrap_index(oy + ly, a.nf2); ok = ok && gy >= 0 && gy < a.nf2; }
            if (DIM > 2) { gz = wrap_index(oz + lz, a.nf3); ok = ok && gz >= 0 && gz < a.nf3; }
            if (!ok) continue;                            // beyond nf + pad on a grid smaller than the bin: never read
            const C *grow = fwt + (size_t)gz * plane + (size_t)gy * a.nf1;
            C *trow = tile + (size_t)row * ex;
            for (int lx = lane; lx < ex; lx += 32) {
                const int gx = wrap_index(ox + lx, a.nf1);
                if (gx >= 0 && gx < a.nf1) cp_async_cell(trow + lx, grow + gx);
            }
        }
        bool first = true;                                // the copy is waited for after the first point's weights
        if (NC > 0) {
            // the class segments of every member bin at once: after this barrier the block walks the bins and their
            // rounds without another one (the tile and these tables are read-only), so the warps drift freely and
            // the shared-memory pipe does not drain at every bin boundary
            for (int i = threadIdx.x; i < nmember * NC; i += blockDim.x) {
                const int m = i / NC, c = i - m * NC;
                const size_t bin = merged ? (size_t)first_bin + (size_t)m * a.nb1 * a.nb2 : (size_t)a.s2b[s_single];
                s_cs[m][c] = a.keyoff[bin * a.cpb + c];
                s_ce[m][c] = a.keyoff[bin * a.cpb + c + 1];
            }
            __syncthreads();
        }

        // Merged tile with bank classes: the member bins' points are dealt to the rounds JOINTLY (class c's pool is the
        // concatenation of its segments in the member bins).  A class that is fuller than its share of slots spills
        // into other classes' free slots, and every such point poisons its quarter-warp with a 2-way conflict for all
        // 1000 stencil loads: 2 % of the points of ONE 477-point bin are misplaced (class sizes 60 +- 7 against 64
        // slots), 0.3 % of a four-bin pool (239 +- 14 against 256) -- profiles/r03t: 4.49 -> wavefronts per LDS.128.
        const bool joint = merged && NC > 0;
        for (int m = 0; m < (joint ? 1 : nmember); ++m) {
        const int bin_m = merged ? first_bin + m * a.nb1 * a.nb2 : 0;
        const int s_lo = merged ? a.substart[bin_m] : s_single;
        const int s_hi = joint ? s_lo + 1 : (merged ? a.substart[bin_m + 1] : s_single + 1);
        for (int s = s_lo; s < s_hi; ++s) {
        int pstart = 0, n = 0, oxb = ox, oyb = oy, ozb = oz;
        if (!joint) decode_subproblem<T, DIM>(a, s, pstart, n, oxb, oyb, ozb);
        int zadd = merged ? m * a.rbs3 * ex * ey : 0;         // the member bin's own tile starts this far into the loaded one
        const int *cs = s_cs[m], *ce = s_ce[m];
        // size of class c: in this bin, or (joint) in the whole group
        auto csize = [&](int c) -> int {
            if constexpr (DIM == 3) {
                if (joint) {
                    int t_ = 0;
                    for (int mm = 0; mm < nmember; ++mm) t_ += s_ce[mm][c] - s_cs[mm][c];
                    return t_;
                }
            }
            return ce[c] - cs[c];
        };
        // sorted point index of position j of class c (joint: walks the member bins; also says which member)
        auto cpoint = [&](int c, int j, int &mm) -> int {
            mm = m;
            if constexpr (DIM == 3) {
                if (joint) {
                    for (mm = 0; mm < nmember - 1; ++mm) {
                        const int k_ = s_ce[mm][c] - s_cs[mm][c];
                        if (j < k_) break;
                        j -= k_;
                    }
                    return s_cs[mm][c] + j;
                }
            }
            return cs[c] + j;
        };

        // ---- thread-per-point
        // The bin's points fill R = ceil(n_bin / blockDim) rounds of blockDim slots; slot (class, position) of
        // round r is position r * cap + t / NC of class t % NC.  The bin's K items (one per maxsub points, the
        // subproblem list is shared with the other engines) split the ROUNDS among themselves, so every item
        // sees all classes and no round is partly empty except the bin's last.
        int rnd = 0, rnd_end = 0, ccap = 0;
        if (NC > 0) {
            int nbin = 0;
            for (int c = 0; c < NC; ++c) nbin += csize(c);
            const int R = (nbin + (int)blockDim.x - 1) / (int)blockDim.x;
            if (joint) { rnd = 0; rnd_end = R; }
            else {
                const int bin = a.s2b[s];
                const int k = s - a.substart[bin], K = a.substart[bin + 1] - a.substart[bin];
                rnd = (int)((long long)R * k / K); rnd_end = (int)((long long)R * (k + 1) / K);
            }
            ccap = R * cap;
        }
        // sorted point index (absolute) of this thread in round `rnd`, or -1; mm = the member bin it belongs to
        auto class_point = [&](int rnd, int &mm) -> int {
            const int j = rnd * cap + (int)(threadIdx.x / NC);
            const int mine = csize(cls);
            if (j < mine) return cpoint(cls, j, mm);
            // a free slot: its rank among the free slots (ordered by class, then position) ...
            int k = j - mine;
            for (int c = 0; c < cls; ++c) k += max(0, ccap - csize(c));
            // ... takes the surplus point (position >= ccap in a fuller class) of the same rank
            for (int c = 0; c < NC; ++c) {
                const int over = csize(c) - ccap;
                if (over > 0) {
                    if (k < over) return cpoint(c, ccap + k, mm);
                    k -= over;
                }
            }
            mm = m;
            return -1;
        };
        for (int i = threadIdx.x; (NC > 0 ? rnd < rnd_end : i < n) || first; i += blockDim.x, ++rnd) {
            int mm = m;
            const int pidx = NC > 0 ? (rnd < rnd_end ? class_point(rnd, mm) : -1) : (i < n ? pstart + i : -1);
            const bool valid = pidx >= 0;
            if (joint) { ozb = oz + mm * a.rbs3; zadd = mm * a.rbs3 * ex * ey; }
            T kx[NS], ky[DIM > 1 ? NS : 1], kz[DIM > 2 ? NS : 1];
            int off = 0, idx = 0;
            if (valid) {
                const PtRec<T> rec = load_rec(a.recs + pidx);
                idx = rec_index(rec);
                const int xs = stencil_start(rec.x, NS);
                kernel_vector<T, NS, true>(kx, (T)xs - rec.x, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
                off = clampi(xs - oxb, 0, ex - NS);
                if (DIM > 1) {
                    const int ys = stencil_start(rec.y, NS);
                    kernel_vector<T, NS, true>(ky, (T)ys - rec.y, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
                    off += clampi(ys - oyb, 0, ey - NS) * ex;
                }
                if (DIM > 2) {
                    const int zs = stencil_start(rec.z, NS);
                    kernel_vector<T, NS, true>(kz, (T)zs - rec.z, a.es_c, a.es_beta, a.horner, s_hc, a.ncoef);
                    off += clampi(zs - a.zshift - ozb, 0, ez - NS) * ex * ey + zadd;
                }
            }
            if (first) {                                  // the tile has landed (weights of the first point overlapped the copy)
                cp_async_wait_all();
                __syncthreads();
                first = false;
            }
            if (!valid) continue;
            const C *base = tile + off;
            T ar = 0, ai = 0;
            if constexpr (DIM == 1) {
#pragma unroll
                for (int ix = 0; ix < NS; ++ix) { const C g = base[ix]; ar = fma(kx[ix], g.x, ar); ai = fma(kx[ix], g.y, ai); }
            } else if constexpr (DIM == 2) {
#pragma unroll
                for (int iy = 0; iy < NS; ++iy) {
                    const C *rowp = base + iy * ex;
                    T rr = 0, ri = 0;
#pragma unroll
                    for (int ix = 0; ix < NS; ++ix) { const C g = rowp[ix]; rr = fma(kx[ix], g.x, rr); ri = fma(kx[ix], g.y, ri); }
                    ar = fma(ky[iy], rr, ar); ai = fma(ky[iy], ri, ai);
                }
            } else {
                const int sz = ex * ey;
                auto plane_sum = [&](const C *pl, T &pr, T &pi) {
                    pr = 0; pi = 0;
#pragma unroll
                    for (int iy = 0; iy < NS; ++iy) {
                        const C *rowp = pl + iy * ex;
                        T rr = 0, ri = 0;
#pragma unroll
                        for (int ix = 0; ix < NS; ++ix) { const C g = rowp[ix]; rr = fma(kx[ix], g.x, rr); ri = fma(kx[ix], g.y, ri); }
                        pr = fma(ky[iy], rr, pr); pi = fma(ky[iy], ri, pi);
                    }
                };
                if constexpr (TileInterp<T, DIM, NS>::ROLL_Z) {
#pragma unroll 1
                    for (int iz = 0; iz < NS; ++iz) {
                        T pr, pi;
                        plane_sum(base + iz * sz, pr, pi);
                        ar = fma(kz[iz], pr, ar); ai = fma(kz[iz], pi, ai);
                    }
                } else {
#pragma unroll
                    for (int iz = 0; iz < NS; ++iz) {
                        T pr, pi;
                        plane_sum(base + iz * sz, pr, pi);
                        ar = fma(kz[iz], pr, ar); ai = fma(kz[iz], pi, ai);
                    }
                }
            }
            cout[idx] = C{ar, ai};
        }
        }   // subproblems of the member bin
        }   // member bins
        if (first) cp_async_wait_all();                   // (cannot happen: a group with points has an item; keeps the copy accounted for)
    }
}

// ---- host-side launch helpers (one instantiation per (T, DIM) translation unit) ----
template <typename T, int DIM> int launch_spread(Plan<T> &p, const typename Plan<T>::C *c, typename Plan<T>::C *fw, int nt);
template <typename T, int DIM> int launch_interp(Plan<T> &p, typename Plan<T>::C *c, const typename Plan<T>::C *fw, int nt);
template <typename T, int DIM> size_t sm_spread_smem_per_warp(int ns, int tile_cells);

}  // namespace cfb
