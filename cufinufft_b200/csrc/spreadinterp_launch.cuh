// spreadinterp_launch.cuh -- host launch code for the spread / interp kernels; included
// by the six si_<prec>_d<dim>.cu translation units (one per precision x dimension so the
// ~140 kernel instantiations compile in parallel).
//
// Replaces the reference's method dispatch wrappers CUSPREADnD / CUINTERPnD and their
// per-transform launch loops (src/2d/spread2d_wrapper.cu:99-164,335-384,615-694,
// src/2d/interp2d_wrapper.cu:88-274 and the 1-D / 3-D twins): here ONE launch covers
// all transforms of the batch.
#pragma once
#include <algorithm>
#include <cstdlib>
#include "spreadinterp.cuh"
#include "spread_sm2.cuh"
#include "spread_plane.cuh"

namespace cfb {

template <typename T>
static SIArgs<T> make_args(Plan<T> &p, int nt)
{
    SIArgs<T> a;
    a.recs = p.recs.template as<PtRec<T>>();
    a.c = nullptr; a.fw = nullptr;
    const bool split = p.ilist;                         // the engines' own work list (finer bins and/or larger items, setpts.cu)
    a.keyoff = p.key_offsets(); a.cpb = p.sortgeo.cpb;
    a.s2b = split ? p.is2b.template as<int>() : p.subprob_to_bin.template as<int>();
    a.substart = split ? p.isubstart.template as<int>() : p.subprobstartpts.template as<int>();
    a.nsub = split ? p.isubstart.template as<int>() + p.nibins : p.scalars.template as<int>();
    a.counter = p.scalars.template as<int>() + 1;
    a.hcoef = p.hcoef.template as<T>();
    a.M = p.M; a.nt = nt; a.maxsub = p.ilist ? p.imaxsub : p.opts.gpu_maxsubprobsize;
    a.nf1 = p.nf1; a.nf2 = p.nf2; a.nf3 = p.nf3;
    a.bs1 = p.ibs[0]; a.bs2 = p.ibs[1]; a.bs3 = p.ibs[2]; a.nb1 = p.nbin[0]; a.nb2 = p.nbin[1];
    a.rbs1 = p.bs[0]; a.rbs2 = p.bs[1]; a.rbs3 = p.bs[2];
    a.spb1 = p.spb[0]; a.spb2 = p.spb[1]; a.spbt = p.spb[0] * p.spb[1] * p.spb[2];
    a.pad = p.tile_pad;
    a.ex = p.ibs[0] + 2 * p.tile_pad; a.ey = p.dim > 1 ? p.ibs[1] + 2 * p.tile_pad : 1;
    a.ez = p.dim > 2 ? p.ibs[2] + 2 * p.tile_pad : 1;
    a.sy = p.tile_sy; a.sz = p.tile_sz; a.tile_cells = p.tile_cells;
    a.horner = p.opts.gpu_kerevalmeth == 1; a.ncoef = p.horner_ncoef;
    a.es_c = p.es_c; a.es_beta = p.es_beta;
    a.zshift = p.slab ? p.zshift : 0;
    a.bankc = p.bank_classes;
    a.zm = 1; a.nb3 = p.nbin[2]; a.npairs = 0;
    // experiments only (read once per process): CFB_DIRECT_THR="num/den" = run length below which a batch goes point by point
    static const struct Thr { int n = 0, d = 1; Thr() { if (const char *e = getenv("CFB_DIRECT_THR")) { int a_ = 0, b_ = 1;
                          if (sscanf(e, "%d/%d", &a_, &b_) >= 1 && a_ > 0 && b_ > 0) { n = a_; d = b_; } } } } thr;
    a.thr_num = thr.n; a.thr_den = thr.d;
    a.fwstride = (long long)p.grid_cells();
    return a;
}

template <typename T, int DIM, int NS, bool HORNER>
static int do_spread_h(Plan<T> &p, SIArgs<T> &a)
{
    using C = typename Plan<T>::C;
    const size_t head = 18 * 16 * sizeof(T);
    if constexpr (sizeof(T) == 4 && DIM >= 2 && sm2_applies(DIM, NS)) {
        // single precision, 2-D / 3-D up to ns = 7: the second-generation SM engine (spread_sm2.cuh);
        // CFB_SM_GEN=1 keeps the first one (A/B measurements)
        static const bool gen1 = [] { const char *e = getenv("CFB_SM_GEN"); return e && e[0] == '1'; }();
        if (p.method == 2 && p.sm_warps > 0 && !gen1) {
            const size_t per_warp = (size_t)p.tile_cells * sizeof(C) + Geo2<DIM, NS>::SCRATCH;
            int warps = std::min(p.sm_warps, Geo2<DIM, NS>::MAXW);
            while (warps > 1 && head + (size_t)warps * per_warp + 1024 > (size_t)p.max_smem_optin) --warps;
            const size_t smem = head + (size_t)warps * per_warp;
            int blocks_per_sm = (int)((size_t)p.max_smem_optin / (smem + 1024));
            if (blocks_per_sm < 1) blocks_per_sm = 1;
            if (blocks_per_sm * warps > 32) blocks_per_sm = 32 / warps > 0 ? 32 / warps : 1;
            CFB_CUDA_OK(cudaMemsetAsync(a.counter, 0, sizeof(int), p.stream));
            auto go = [&](auto kern) -> int {
                CFB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<p.num_sms * blocks_per_sm, 32 * warps, smem, p.stream>>>(a);
                return 0;
            };
            int e = 0;
            if constexpr (DIM == 3 && NS == 6) {            // row-order class picked with the tile strides (spread.cu: plan_tile_geometry)
                if (p.sm2_rmc == 1) e = go(spread_sm2_kernel<DIM, NS, HORNER, 1>);
                else if (p.sm2_rmc == 2) e = go(spread_sm2_kernel<DIM, NS, HORNER, 2>);
                else e = go(spread_sm2_kernel<DIM, NS, HORNER, 0>);
            } else e = go(spread_sm2_kernel<DIM, NS, HORNER, 0>);
            if (e) return e;
            p.launches_exec++;
            CFB_CUDA_OK(cudaGetLastError());
            return 0;
        }
    }
    if constexpr (sizeof(T) == 8 && DIM == 3 && plane_engine_ns(NS)) {
        // double precision, 3-D, wide stencils: the block owns the tile, the warps its planes (spread_plane.cuh)
        if (p.method == 2 && p.plane_engine) {
            const int warps = std::min(p.sm_warps, GeoP<NS>::MAXW);
            const size_t smem = head + (size_t)p.tile_cells * sizeof(C) + GeoP<NS>::SCRATCH;
            CFB_CUDA_OK(cudaFuncSetAttribute(spread_plane_kernel<NS, HORNER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CFB_CUDA_OK(cudaMemsetAsync(a.counter, 0, sizeof(int), p.stream));
            spread_plane_kernel<NS, HORNER><<<p.num_sms, 32 * warps, smem, p.stream>>>(a);
            p.launches_exec++;
            CFB_CUDA_OK(cudaGetLastError());
            return 0;
        }
    }
    if (p.method == 2 && p.sm_warps > 0) {
        size_t per_warp = (size_t)p.tile_cells * sizeof(C) + warp_scratch_bytes<T, DIM, NS, false>();
        size_t smem = head + (size_t)p.sm_warps * per_warp;
        CFB_CUDA_OK(cudaFuncSetAttribute(spread_sm_kernel<T, DIM, NS, HORNER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int blocks_per_sm = (int)((size_t)p.max_smem_optin / (smem + 1024));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        if (blocks_per_sm * p.sm_warps > 32) blocks_per_sm = 32 / p.sm_warps > 0 ? 32 / p.sm_warps : 1;
        CFB_CUDA_OK(cudaMemsetAsync(a.counter, 0, sizeof(int), p.stream));
        spread_sm_kernel<T, DIM, NS, HORNER><<<p.num_sms * blocks_per_sm, 32 * p.sm_warps, smem, p.stream>>>(a);
    } else {
        const int warps = 8;
        size_t smem = head + warps * warp_scratch_bytes<T, DIM, NS>();
        CFB_CUDA_OK(cudaFuncSetAttribute(spread_gm_kernel<T, DIM, NS, HORNER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long nb = (((long long)p.M + 255) / 256 * a.nt + warps - 1) / warps;
        long long cap = (long long)p.num_sms * 8;
        int blocks = (int)(nb < 1 ? 1 : (nb > cap ? cap : nb));
        spread_gm_kernel<T, DIM, NS, HORNER><<<blocks, 32 * warps, smem, p.stream>>>(a);
    }
    p.launches_exec++;
    CFB_CUDA_OK(cudaGetLastError());
    return 0;
}

// Tile engine (interp_tile_kernel): used whenever the points are bin-sorted and the bin tile
// with its halo fits in shared memory; the gather engine (interp_kernel) serves the rest
// (gpu_sort = 0, very wide 3-D fp64 stencils).  p.interp_engine: 0 auto, 1 gather, 2 tile.
template <typename T, int DIM, int NS, bool HORNER>
static int do_interp_tile(Plan<T> &p, SIArgs<T> &a, bool &done)
{
    using C = typename Plan<T>::C;
    done = false;
    // Sparse inputs: a tile of `cells` grid values is staged for every subproblem, which only pays
    // when enough points share it.  Measured crossover (tools/ab_interp.py, profiles/r01j_ab_lowdensity):
    // about one point per 64 tile cells (3-D fp64 ns=10: 127 points per bin, 2-D fp32 ns=4: 20).
    // The gather engine lives on L2 reuse between neighbouring bins, though: in 3-D it needs the ns
    // planes a stencil spans to stay resident (config 5's 1024^2 x 10 planes = 168 MB do not: measured
    // 6.1 ns/point gather vs 3.4 tile at 60 points per bin, profiles/r01zc), else the tile engine stays.
    // (the rule itself: interp_tile_applies, spread.cu -- setpts uses it too)
    if (!interp_tile_applies(p)) return 0;
    const size_t head = 18 * 16 * sizeof(T);
    const size_t cells = (size_t)a.ex * a.ey * a.ez;
    size_t smem = head + cells * sizeof(C);
    const int threads = (DIM == 3 && smem > 96 * 1024) ? 512 : 256;
    // one block per SM anyway (wide fp64 stencils in 3-D): let the tile serve as many z-adjacent bins as fit
    // (interp_tile_kernel: merged); CFB_INTERP_ZM=1 keeps one bin per tile (A/B measurements)
    if (DIM == 3 && threads == 512 && !p.ilist && a.spbt == 1) {
        static const int zm_env = [] { const char *e = getenv("CFB_INTERP_ZM"); return e ? atoi(e) : 0; }();
        for (int zm = zm_env > 0 && zm_env < 4 ? zm_env : 4; zm >= 2; --zm) {     // <= 4: interp_tile_kernel's class tables
            const size_t sm = head + (size_t)a.ex * a.ey * (a.rbs3 * zm + 2 * a.pad) * sizeof(C);
            if (zm <= a.nb3 && sm + 2048 <= (size_t)p.max_smem_optin) {
                a.zm = zm; a.npairs = a.nb1 * a.nb2 * ((a.nb3 + zm - 1) / zm); smem = sm;
                break;
            }
        }
    }
    CFB_CUDA_OK(cudaFuncSetAttribute(interp_tile_kernel<T, DIM, NS, HORNER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CFB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, interp_tile_kernel<T, DIM, NS, HORNER>, threads, smem));
    if (occ < 1) return 0;
    CFB_CUDA_OK(cudaMemsetAsync(a.counter, 0, sizeof(int), p.stream));
    interp_tile_kernel<T, DIM, NS, HORNER><<<p.num_sms * occ, threads, smem, p.stream>>>(a);
    p.launches_exec++;
    CFB_CUDA_OK(cudaGetLastError());
    done = true;
    return 0;
}

template <typename T, int DIM, int NS, bool HORNER>
static int do_interp_h(Plan<T> &p, SIArgs<T> &a)
{
    bool done = false;
    if (int e = do_interp_tile<T, DIM, NS, HORNER>(p, a, done)) return e;
    if (done) return 0;
    const size_t head = 18 * 16 * sizeof(T);
    const int warps = 8;
    size_t smem = head + warps * warp_scratch_bytes<T, DIM, NS>();
    CFB_CUDA_OK(cudaFuncSetAttribute(interp_kernel<T, DIM, NS, HORNER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long nb = (((long long)p.M + 255) / 256 * a.nt + warps - 1) / warps;
    long long cap = (long long)p.num_sms * 8;
    int blocks = (int)(nb < 1 ? 1 : (nb > cap ? cap : nb));
    interp_kernel<T, DIM, NS, HORNER><<<blocks, 32 * warps, smem, p.stream>>>(a);
    p.launches_exec++;
    CFB_CUDA_OK(cudaGetLastError());
    return 0;
}

// one instantiation per kernel evaluator (gpu_kerevalmeth 0: exp/sqrt, 1: Horner)
template <typename T, int DIM, int NS>
static int do_spread(Plan<T> &p, SIArgs<T> &a)
{
    return a.horner ? do_spread_h<T, DIM, NS, true>(p, a) : do_spread_h<T, DIM, NS, false>(p, a);
}
template <typename T, int DIM, int NS>
static int do_interp(Plan<T> &p, SIArgs<T> &a)
{
    return a.horner ? do_interp_h<T, DIM, NS, true>(p, a) : do_interp_h<T, DIM, NS, false>(p, a);
}

template <typename T> struct max_ns;
template <> struct max_ns<float>  { static constexpr int v = 16; };  // upsampfac 2 clamps at ns = 9, other upsampling factors do not (sigma 1.25, tol 1e-6: ns = 10)
template <> struct max_ns<double> { static constexpr int v = 16; };

// development builds (make EXTRA=-DCFB_DEV_NS) instantiate only the widths of the five
// BASELINE.json configs and the smoke test; the other widths then return error 10
#ifdef CFB_DEV_NS
constexpr bool ns_enabled(int ns) { return ns == 4 || ns == 5 || ns == 6 || ns == 7 || ns == 10 || ns == 11 || ns == 13; }
#else
constexpr bool ns_enabled(int) { return true; }
#endif

template <typename T, int DIM, int NS>
struct NsDispatch {
    static int spread(Plan<T> &p, SIArgs<T> &a) {
        if constexpr (ns_enabled(NS)) { if (p.ns == NS) return do_spread<T, DIM, NS>(p, a); }
        return NsDispatch<T, DIM, NS - 1>::spread(p, a);
    }
    static int interp(Plan<T> &p, SIArgs<T> &a) {
        if constexpr (ns_enabled(NS)) { if (p.ns == NS) return do_interp<T, DIM, NS>(p, a); }
        return NsDispatch<T, DIM, NS - 1>::interp(p, a);
    }
    static size_t scratch(int ns) {
        if (ns == NS) {
            if constexpr (sizeof(T) == 4 && DIM >= 2 && sm2_applies(DIM, NS))
                return std::max(warp_scratch_bytes<T, DIM, NS, false>(), (size_t)Geo2<DIM == 1 ? 2 : DIM, NS>::SCRATCH);
            return warp_scratch_bytes<T, DIM, NS, false>();
        }
        return NsDispatch<T, DIM, NS - 1>::scratch(ns);
    }
};
template <typename T, int DIM>
struct NsDispatch<T, DIM, 1> {
    static int spread(Plan<T> &, SIArgs<T> &) { return 10; }
    static int interp(Plan<T> &, SIArgs<T> &) { return 10; }
    static size_t scratch(int) { return 0; }
};

template <typename T, int DIM>
int launch_spread(Plan<T> &p, const typename Plan<T>::C *c, typename Plan<T>::C *fw, int nt)
{
    SIArgs<T> a = make_args(p, nt);
    a.c = const_cast<typename Plan<T>::C *>(c);
    a.fw = fw;
    return NsDispatch<T, DIM, max_ns<T>::v>::spread(p, a);
}

template <typename T, int DIM>
int launch_interp(Plan<T> &p, typename Plan<T>::C *c, const typename Plan<T>::C *fw, int nt)
{
    SIArgs<T> a = make_args(p, nt);
    a.c = c;
    a.fw = const_cast<typename Plan<T>::C *>(fw);
    return NsDispatch<T, DIM, max_ns<T>::v>::interp(p, a);
}

template <typename T, int DIM>
size_t sm_spread_smem_per_warp(int ns, int tile_cells)
{
    return (size_t)tile_cells * sizeof(typename Plan<T>::C) + NsDispatch<T, DIM, max_ns<T>::v>::scratch(ns);
}

#define CFB_INSTANTIATE_SI(T, DIM)                                                                           \
    template int launch_spread<T, DIM>(Plan<T> &, const typename Plan<T>::C *, typename Plan<T>::C *, int);  \
    template int launch_interp<T, DIM>(Plan<T> &, typename Plan<T>::C *, const typename Plan<T>::C *, int);  \
    template size_t sm_spread_smem_per_warp<T, DIM>(int, int);

}  // namespace cfb
